"""CPU tests: the C restatement (oracle/cgmres_oracle.c) is pinned bit for bit to
(a) the committed golden fixtures generated from the unmodified reference, and
(b) the compiled reference itself (oracle/_ref) whenever it is present."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = (po.MSD, po.ARM, po.SEMIACTIVE)


def gold(model):
    return np.load(os.path.join(GOLD, f"{po.MODEL_NAMES[model]}.npz"))


@pytest.mark.parametrize("model", MODELS)
def test_dims_and_params_match_reference_headers(oracle_port, model):
    # <example>/model.hpp:7-34
    want = {po.MSD: (4, 6, 2, 50, 5, 2), po.ARM: (4, 3, 2, 25, 5, 1), po.SEMIACTIVE: (2, 3, 0, 50, 5, 1)}[model]
    d = oracle_port.dims(model)
    assert (d.dim_x, d.dim_u, d.dim_p, d.dv, d.k_max, d.n_ctrl) == want
    p = oracle_port.params(model)
    assert p["dt"] == 0.001 and p["h"] == 0.002 and p["zeta"] == 1000.0 and p["alpha"] == 0.5 and p["tol"] == 1e-6
    assert p["Tf"] == (0.5 if model == po.ARM else 1.0)


@pytest.mark.parametrize("model", MODELS)
def test_port_matches_golden_shipped_run(oracle_port, model):
    g, s = gold(model), po.SHIPPED[model]
    p = [s["p"]] if s["p"] else None
    a = oracle_port.run_closed_loop(model, [s["x0"]], p, s["u0"], 2000, rec_stride=100)
    assert np.array_equal(a["x_traj"][:, 0], g["shipped_x_traj"])
    assert np.array_equal(a["u_traj"][:, 0], g["shipped_u_traj"])
    a = oracle_port.run_closed_loop(model, [s["x0"]], p, s["u0"], int(g["shipped_steps"]), want_U=True)
    for k in ("x_fin", "u_fin", "U_fin", "dUdt_fin"):
        assert np.array_equal(a[k][0], g["shipped_" + k]), k


def test_shipped_final_states_match_reference_text_output(oracle_port):
    # last lines of the reference's own <example>_x.txt ("%f")
    want = {po.MSD: [0.897925, -0.934620, -0.000013, -0.000001], po.ARM: [0.786209, 0.000060, -0.003976, 0.000335],
            po.SEMIACTIVE: [0.086840, -0.004049]}
    for model, w in want.items():
        g = gold(model)
        assert [float("%f" % v) for v in g["shipped_x_fin"]] == w


@pytest.mark.parametrize("model", MODELS)
def test_port_matches_golden_batch(oracle_port, model):
    g = gold(model)
    a = oracle_port.run_closed_loop(model, g["batch_x0"], g["batch_p"], g["batch_u0"], 1000, rec_stride=100,
                                    want_U=True, n_threads=4)
    assert np.array_equal(a["x_traj"], g["batch_x_traj"])
    assert np.array_equal(a["u_traj"], g["batch_u_traj"])
    assert np.array_equal(a["U_fin"], g["batch_U_fin"])
    assert np.array_equal(a["dUdt_fin"], g["batch_dUdt_fin"])
    # the synthetic batch generator is part of the fixture contract
    x0, p, u0 = po.synthetic_batch(model, 8)
    assert np.array_equal(x0, g["batch_x0"]) and np.array_equal(p, g["batch_p"]) and np.array_equal(u0, g["batch_u0"])


@pytest.mark.parametrize("model", MODELS)
def test_port_teacher_forced_matches_golden(oracle_port, model):
    g, s = gold(model), po.SHIPPED[model]
    c = oracle_port.controller(model)
    if c.dims.dim_p:
        c.set_ptau_repeat(s["p"])
    u = c.init_u0_newton(s["u0"], s["x0"], s["p"] if c.dims.dim_p else [0.0], 10)
    assert np.array_equal(u, g["newton_u0"])
    for i in range(len(g["tf_steps"])):
        c.set_state(float(g["tf_t"][i]), g["tf_U"][i], g["tf_dUdt"][i])
        u = c.control(g["tf_x"][i])
        t, U, dUdt = c.get_state()
        assert np.array_equal(U, g["tf_U_after"][i])
        assert np.array_equal(dUdt, g["tf_dUdt_after"][i])
        assert np.array_equal(u, g["tf_u_after"][i])


def test_exit_paths_of_the_shipped_msd_run(oracle_port):
    # 12 942 full, 7 038 early convergences, 21 rho0<tol returns, no breakdown
    s = po.SHIPPED[po.MSD]
    a = oracle_port.run_closed_loop(po.MSD, [s["x0"]], [s["p"]], s["u0"], s["steps"])
    assert a["exit_hist"][0].tolist() == [12942, 7038, 21, 0]


def test_time_varying_reference_and_threads_are_deterministic(oracle_port):
    dm = oracle_port.dims(po.MSD)
    x0, p, u0 = po.synthetic_batch(po.MSD, 6)
    ramp = np.linspace(0.0, 0.2, dm.dv + 1)[None, :, None]
    pfull = (p[:, None, :] + ramp).reshape(6, -1)
    a = oracle_port.run_closed_loop(po.MSD, x0, pfull, u0, 50, p_full=True, n_threads=1, want_U=True)
    b = oracle_port.run_closed_loop(po.MSD, x0, pfull, u0, 50, p_full=True, n_threads=3, want_U=True)
    assert np.array_equal(a["U_fin"], b["U_fin"]) and np.array_equal(a["x_fin"], b["x_fin"])
    c = oracle_port.run_closed_loop(po.MSD, x0, p, u0, 50, want_U=True)
    assert not np.array_equal(a["U_fin"], c["U_fin"])


def test_empty_batch(oracle_port):
    a = oracle_port.run_closed_loop(po.MSD, np.zeros((0, 4)), np.zeros((0, 2)), po.SHIPPED[po.MSD]["u0"], 10)
    assert a["x_fin"].shape == (0, 4)


@pytest.mark.skipif(not po.available("reference"), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("model", MODELS)
def test_port_is_bit_identical_to_compiled_reference(oracle_port, model):
    ref = po.load("reference")
    assert ref.dims(model) == oracle_port.dims(model) and ref.params(model) == oracle_port.params(model)
    x0, p, u0 = po.synthetic_batch(model, 24, seed=777)
    a = oracle_port.run_closed_loop(model, x0, p, u0, 1200, rec_stride=50, want_U=True, n_threads=8)
    b = ref.run_closed_loop(model, x0, p, u0, 1200, rec_stride=50, want_U=True, n_threads=8)
    for k in ("x_traj", "u_traj", "x_fin", "u_fin", "U_fin", "dUdt_fin"):
        assert np.array_equal(a[k], b[k]), k
