"""GPU parity tests of the on-chip control-update kernels (cgmres_cpp_b200/csrc/fast_update.cuh, pipe_update.cuh):

  MODE_ONCHIP_EXACT  the on-chip kernel with the reference's sequential sums and no FMA -> bit-identical to the
                     oracle (mass_spring_damper, semiactive_damper), which verifies its data movement and control flow;
  MODE_FAST          the persistent warp-specialised kernel with FMA contraction and shuffle reductions -> the
                     north-star tolerances: |dU|_inf/|U|_inf <= 1e-9 per (teacher-forced) update, max|dx| <= 1e-6
                     over 1000 closed-loop steps.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_parity import BIT_EXACT, MODELS, TOL_U_REL, TOL_X_ABS, gold, make, rel_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg(built):
    import cgmres_cpp_b200 as m

    if m.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return m


# ---------------------------------------------------------------- on-chip kernels, exact arithmetic
# MODE_ONCHIP_EXACT = the first-generation on-chip kernel; MODE_PIPELINED_EXACT = the FAST mode's persistent pipelined
# kernel built with sequential sums and no FMA: everything below must be bit-identical for both
EXACT_ONCHIP = ["MODE_ONCHIP_EXACT", "MODE_PIPELINED_EXACT"]


@pytest.mark.parametrize("mode_name", EXACT_ONCHIP)
@pytest.mark.parametrize("model", MODELS)
def test_onchip_exact_golden_batch_1000_steps(cg, model, mode_name):
    g = gold(model)
    c, un = make(cg, model, g["batch_x0"], g["batch_p"], g["batch_u0"], mode=getattr(cg, mode_name))
    for r in range(10):
        c.step_closed_loop(100)
        x, u = c.get_x(), c.get_u()
        if BIT_EXACT[model]:
            assert np.array_equal(x, g["batch_x_traj"][r]), (r, np.abs(x - g["batch_x_traj"][r]).max())
            assert np.array_equal(u, g["batch_u_traj"][r])
        else:
            assert np.abs(x - g["batch_x_traj"][r]).max() <= TOL_X_ABS
    _, U, dUdt = c.get_state()
    if BIT_EXACT[model]:
        assert np.array_equal(U, g["batch_U_fin"]) and np.array_equal(dUdt, g["batch_dUdt_fin"])
    c.close()


@pytest.mark.parametrize("mode_name", EXACT_ONCHIP)
@pytest.mark.parametrize("model", [po.MSD, po.SEMIACTIVE])
@pytest.mark.parametrize("n", [1, 7, 45])
def test_onchip_exact_ragged_batches_match_oracle(cg, oracle_best, model, n, mode_name):
    """n not a multiple of the instances-per-CTA group (partial last CTA / partly filled groups)."""
    x0, p, u0 = po.synthetic_batch(model, n, seed=31 + n)
    want = oracle_best.run_closed_loop(model, x0, p, u0, 120, want_U=True)
    c, _ = make(cg, model, x0, p, u0, mode=getattr(cg, mode_name))
    c.step_closed_loop(120)
    assert np.array_equal(c.get_x(), want["x_fin"])
    t, U, dUdt = c.get_state()
    assert np.array_equal(U, want["U_fin"]) and np.array_equal(dUdt, want["dUdt_fin"])
    c.close()


@pytest.mark.parametrize("mode_name", EXACT_ONCHIP)
def test_onchip_exact_exit_paths_on_long_msd_run(cg, oracle_port, mode_name):
    """Early convergence (one column dropped) and rho0<tol stale-dUdt returns (0-4): the shipped msd
    run reaches them after step 11 875 / 13 286.  The oracle runs the first 11 800 steps, the GPU takes over its
    checkpoint {t,U,dUdt,x} and both continue for 2 400 steps; exit path, columns used and state must agree."""
    model, s = po.MSD, po.SHIPPED[po.MSD]
    ctl = oracle_port.controller(model)
    ctl.set_ptau_repeat(s["p"])
    x = np.array(s["x0"])
    ctl.init_u0_newton(s["u0"], x, s["p"], 10)
    for _ in range(11800):
        oracle_port.plant_step(model, x, ctl.control(x))
    c = cg.BatchedCgmres(model, 1, mode=getattr(cg, mode_name))
    c.set_ptau_repeat([s["p"]])
    t, U, dUdt = ctl.get_state()
    c.set_state(t, U[None], dUdt[None])
    c.set_x(x[None])
    seen = set()
    for step in range(2400):
        u = ctl.control(x)
        oracle_port.plant_step(model, x, u)
        c.step_closed_loop(1)
        code, ncol = c.get_status()
        assert (int(code[0]), int(ncol[0])) == ctl.last_status(), step
        seen.add((int(code[0]), int(ncol[0])))
    assert np.array_equal(c.get_x()[0], x)
    assert np.array_equal(c.get_state()[1][0], ctl.get_state()[1])
    assert np.array_equal(c.get_state()[2][0], ctl.get_state()[2])
    print("exit paths seen:", sorted(seen))
    assert {(0, 5), (1, 0), (1, 1), (1, 2), (1, 3), (1, 4), (2, 0)} <= seen
    c.close()



@pytest.mark.parametrize("mode_name", EXACT_ONCHIP)
def test_onchip_exact_time_varying_reference_and_host_api(cg, oracle_best, mode_name):
    model, n = po.MSD, 13
    dm = oracle_best.dims(model)
    x0, p, u0 = po.synthetic_batch(model, n, seed=5)
    ramp = np.linspace(0.0, 0.3, dm.dv + 1)[None, :, None]
    pfull = np.ascontiguousarray((p[:, None, :] + ramp).reshape(n, -1))
    want = oracle_best.run_closed_loop(model, x0, pfull, u0, 60, p_full=True, want_U=True)
    c, _ = make(cg, model, x0, pfull, u0, mode=getattr(cg, mode_name), ptau_full=True)
    x = x0.copy()
    for _ in range(60):  # host-buffer API like the reference's main()
        u = c.control(x)
        for i in range(n):
            oracle_best.plant_step(model, x[i], u[i])
    assert np.array_equal(x, want["x_fin"])
    assert np.array_equal(c.get_state()[1], want["U_fin"])
    c.close()


# ---------------------------------------------------------------- fast mode, tolerance bars
@pytest.mark.parametrize("model", MODELS)
def test_fast_teacher_forced_update(cg, model):
    g, s = gold(model), po.SHIPPED[model]
    c = cg.BatchedCgmres(model, 1, mode=cg.MODE_FAST)
    if c.dim_p:
        c.set_ptau_repeat([s["p"]])
    worst = 0.0
    for i in range(len(g["tf_steps"])):
        c.set_state(float(g["tf_t"][i]), g["tf_U"][i][None], g["tf_dUdt"][i][None])
        u = c.control(g["tf_x"][i][None])
        _, U, _ = c.get_state(want_dUdt=False)
        worst = max(worst, rel_inf(U, g["tf_U_after"][i][None]), rel_inf(u, g["tf_u_after"][i][None]))
    print(f"fast {po.MODEL_NAMES[model]}: worst teacher-forced rel dU = {worst:.3e}")
    assert worst <= TOL_U_REL
    c.close()


@pytest.mark.parametrize("model", MODELS)
def test_fast_closed_loop_1000_steps(cg, oracle_best, model):
    g = gold(model)
    n = 264  # 8 golden instances + 256 seeded ones
    x0s, ps, u0 = po.synthetic_batch(model, n - 8, seed=4242)
    x0 = np.concatenate([g["batch_x0"], x0s])
    p = np.concatenate([g["batch_p"], ps]) if ps.shape[1] else np.zeros((n, 0))
    want = oracle_best.run_closed_loop(model, x0, p, u0, 1000, rec_stride=100, n_threads=os.cpu_count() or 4)
    assert np.array_equal(want["x_traj"][:, :8], g["batch_x_traj"])  # the oracle reproduces the golden rows
    c, _ = make(cg, model, x0, p, u0, mode=cg.MODE_FAST)
    worst = 0.0
    for r in range(10):
        c.step_closed_loop(100)
        worst = max(worst, float(np.abs(c.get_x() - want["x_traj"][r]).max()))
    print(f"fast {po.MODEL_NAMES[model]}: closed-loop max|dx| over 1000 steps, {n} instances = {worst:.3e}")
    assert worst <= TOL_X_ABS
    code, _ = c.get_status()
    assert ((code >= 0) & (code <= 3)).all()
    c.close()


@pytest.mark.parametrize("model", [po.MSD, po.SEMIACTIVE, po.ARM])
@pytest.mark.parametrize("n", [5, 17, 32, 9481])
def test_fast_group_and_round_edges_track_the_bit_exact_kernel(cg, model, n):
    """The fast kernel is persistent and warp-specialised (csrc/pipe_update.cuh): one CTA holds two groups of 16
    instances and loops over rounds.  Batch sizes that leave a group partly filled (5), the second group partly
    filled (17), exactly one round (32), and several rounds per CTA with a ragged tail (9481 > 2 * 148 * 32) must
    all give what the bit-exact on-chip kernel gives, to the closed-loop tolerance, with the same exit paths."""
    steps = 60
    x0, p, u0 = po.synthetic_batch(model, n, seed=900 + n)
    got = {}
    for mode in (cg.MODE_FAST, cg.MODE_ONCHIP_EXACT):
        c, _ = make(cg, model, x0, p, u0, mode=mode)
        c.step_closed_loop(steps)
        got[mode] = (c.get_x(), c.get_u(), c.get_status(), c.get_state()[1])
        c.close()
    xf, uf, (cf, kf), Uf = got[cg.MODE_FAST]
    xe, ue, (ce, ke), Ue = got[cg.MODE_ONCHIP_EXACT]
    assert np.isfinite(xf).all() and np.isfinite(Uf).all()
    assert np.abs(xf - xe).max() <= TOL_X_ABS
    assert rel_inf(Uf, Ue) <= 1e-6 and rel_inf(uf, ue) <= 1e-6
    assert np.array_equal(cf, ce) and np.array_equal(kf, ke)


def test_pipelined_exact_equals_first_generation_kernel_on_multi_round_batch(cg):
    """9481 instances = several rounds per persistent CTA with a ragged tail: the pipelined kernel's exact build
    must equal the first-generation on-chip kernel bit for bit (x, u, U, dUdt, exit status)."""
    model, n, steps = po.MSD, 9481, 40
    x0, p, u0 = po.synthetic_batch(model, n, seed=77)
    got = {}
    for mode in (cg.MODE_PIPELINED_EXACT, cg.MODE_ONCHIP_EXACT):
        c, _ = make(cg, model, x0, p, u0, mode=mode)
        c.step_closed_loop(steps)
        got[mode] = (c.get_x(), c.get_u(), c.get_state()[1], c.get_state()[2], *c.get_status())
        c.close()
    for a, b in zip(got[cg.MODE_PIPELINED_EXACT], got[cg.MODE_ONCHIP_EXACT]):
        assert np.array_equal(a, b)


def test_fast_exit_paths_on_long_msd_run(cg, oracle_port):
    """The early-exit paths (converged with 0..4 columns used) through the fast kernel's deferred final
    update: same checkpointed segment of the shipped msd run as the bit-exact test above; the fast kernel must
    visit the early exits and stay within the closed-loop bar of the oracle over the 2 400 steps."""
    model, s = po.MSD, po.SHIPPED[po.MSD]
    ctl = oracle_port.controller(model)
    ctl.set_ptau_repeat(s["p"])
    x = np.array(s["x0"])
    ctl.init_u0_newton(s["u0"], x, s["p"], 10)
    for _ in range(11800):
        oracle_port.plant_step(model, x, ctl.control(x))
    c = cg.BatchedCgmres(model, 1, mode=cg.MODE_FAST)
    c.set_ptau_repeat([s["p"]])
    t, U, dUdt = ctl.get_state()
    c.set_state(t, U[None], dUdt[None])
    c.set_x(x[None])
    seen, worst = set(), 0.0
    for step in range(2400):
        oracle_port.plant_step(model, x, ctl.control(x))
        c.step_closed_loop(1)
        code, ncol = c.get_status()
        seen.add((int(code[0]), int(ncol[0])))
        if step % 100 == 99:
            worst = max(worst, float(np.abs(c.get_x()[0] - x).max()))
    print("fast exit paths seen:", sorted(seen), "max|dx| =", worst)
    assert worst <= TOL_X_ABS
    # (the rho0 < tol return needs a residual below 1e-6; with the fast mode's rounding it may or may not occur)
    assert (0, 5) in seen and {(1, 0), (1, 1), (1, 2), (1, 3), (1, 4)} <= seen
    c.close()


def test_fast_time_varying_reference_and_host_api(cg, oracle_best):
    """set_ptau with a full per-stage reference trajectory (cgmres.hpp:36-39) through the fast kernel: the serial
    warps and the stage-parallel dHdu read p per stage from global memory; host-buffer control() like main.cpp."""
    model, n = po.MSD, 37
    dm = oracle_best.dims(model)
    x0, p, u0 = po.synthetic_batch(model, n, seed=6)
    ramp = np.linspace(0.0, 0.3, dm.dv + 1)[None, :, None]
    pfull = np.ascontiguousarray((p[:, None, :] + ramp).reshape(n, -1))
    want = oracle_best.run_closed_loop(model, x0, pfull, u0, 200, p_full=True, want_U=True)
    c, _ = make(cg, model, x0, pfull, u0, mode=cg.MODE_FAST, ptau_full=True)
    x = x0.copy()
    for _ in range(200):
        u = c.control(x)
        cg.plant_step_host(model, x, u)
    assert np.abs(x - want["x_fin"]).max() <= TOL_X_ABS
    # (the bars are on x in closed loop and on U per teacher-forced update; closed-loop U is only a sanity check)
    assert rel_inf(c.get_state()[1], want["U_fin"]) <= 1e-3
    c.close()


def test_fast_full_size_batch_shard_invariance(cg):
    """65,536 instances: every instance of the big batch equals the same instance run in a small batch, bit for
    bit (no cross-instance arithmetic exists), and stays finite."""
    model, n, steps = po.MSD, 65536, 10
    x0, p, u0 = po.synthetic_batch(model, n, seed=2024)
    c, _ = make(cg, model, x0, p, u0, mode=cg.MODE_FAST)
    c.step_closed_loop(steps)
    x_big = c.get_x()
    c.close()
    assert np.isfinite(x_big).all()
    lo, hi = 40000, 40123
    c2, _ = make(cg, model, x0[lo:hi], p[lo:hi], u0, mode=cg.MODE_FAST)
    c2.step_closed_loop(steps)
    assert np.array_equal(c2.get_x(), x_big[lo:hi])
    c2.close()


@pytest.mark.parametrize("mode_name", ["MODE_ONCHIP_EXACT", "MODE_FAST"])
def test_sliced_host_control_equals_device_closed_loop(cg, mode_name):
    """Large batches take the pipelined path of cgmres_b200_control (4 slices on side streams, copies overlapping
    kernels).  Instances are independent, so it must give exactly what the device-resident closed loop gives."""
    mode = getattr(cg, mode_name)
    model, n, steps = po.MSD, 8200 + 37, 6
    x0, p, u0 = po.synthetic_batch(model, n, seed=77)
    a, _ = make(cg, model, x0, p, u0, mode=mode)
    a.step_closed_loop(steps)
    b, _ = make(cg, model, x0, p, u0, mode=mode)
    x = x0.copy()
    for _ in range(steps):
        u = b.control(x)
        cg.plant_step_host(model, x, u)
    ta, Ua, dUa = a.get_state()
    tb, Ub, dUb = b.get_state()
    assert ta == tb
    if mode == cg.MODE_ONCHIP_EXACT:
        assert np.array_equal(a.get_x(), x)
        assert np.array_equal(Ua, Ub) and np.array_equal(dUa, dUb) and np.array_equal(a.get_u(), b.get_u())
    else:  # the in-kernel plant step of the fast build contracts x + f*dt into an FMA, the host plant does not
        assert np.abs(a.get_x() - x).max() <= 1e-12
        assert rel_inf(Ua, Ub) <= 1e-12
    a.close()
    b.close()


@pytest.mark.parametrize("mode_name", ["MODE_EXACT", "MODE_ONCHIP_EXACT", "MODE_FAST"])
def test_rk4_plant_option(cg, mode_name):
    """The optional on-device RK4 plant (north star item 5; the reference only has Euler, so there is no reference
    oracle): one closed-loop step must equal control() + an independent numpy RK4 of the same plant equations."""
    mode = getattr(cg, mode_name)
    model, n = po.SEMIACTIVE, 50
    x0, p, u0 = po.synthetic_batch(model, n, seed=9)
    a, _ = make(cg, model, x0, p, u0, mode=mode)
    a.set_plant_integrator("rk4")
    b, _ = make(cg, model, x0, p, u0, mode=mode)
    x = x0.copy()
    dt = 0.001

    def f(x, u):  # semiactive_damper/simulator.hpp:13-16
        return np.stack([x[:, 1], -1.0 * x[:, 0] + -1.0 * u[:, 0] * x[:, 1]], axis=1)

    for _ in range(20):
        a.step_closed_loop(1)
        u = b.control(x)
        k1 = f(x, u)
        k2 = f(x + 0.5 * dt * k1, u)
        k3 = f(x + 0.5 * dt * k2, u)
        k4 = f(x + dt * k3, u)
        x = x + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        assert np.abs(a.get_x() - x).max() <= 1e-13
    a.set_plant_integrator("euler")
    a.close()
    b.close()


@pytest.mark.parametrize("mode_name", ["MODE_EXACT", "MODE_ONCHIP_EXACT", "MODE_FAST"])
def test_per_instance_start_times(cg, oracle_port, mode_name):
    """Controllers started at different times: each instance carries its own clock t_i, so its
    horizon step dtau(t_i) differs.  The oracle runs every instance as its own object with its own t; the device
    evaluates exp() itself (<= 1 ulp from glibc), hence the tolerance bars instead of bit equality."""
    mode = getattr(cg, mode_name)
    model, n, steps = po.MSD, 21, 150
    x0, p, u0 = po.synthetic_batch(model, n, seed=17)
    # staggered starts within 10 ms: the continuation method needs U consistent with the horizon length, so a
    # controller initialised by init_u0_newton cannot be dropped at a large t (the reference diverges there too)
    t0 = np.linspace(0.0, 0.01, n)
    c, un = make(cg, model, x0, p, u0, mode=mode)
    c.set_t(t0)
    c.step_closed_loop(steps)
    x = c.get_x()
    _, U, _ = c.get_state(want_dUdt=False)
    assert np.allclose(c.get_t(), t0 + steps * c.dt, rtol=0, atol=1e-12)
    for i in range(n):
        k = oracle_port.controller(model)
        k.set_ptau_repeat(p[i])
        k.init_u0_newton(u0, x0[i], p[i], 10)
        k.set_state(t=float(t0[i]))
        xi = x0[i].copy()
        for _ in range(steps):
            oracle_port.plant_step(model, xi, k.control(xi))
        assert np.abs(x[i] - xi).max() <= TOL_X_ABS, (i, np.abs(x[i] - xi).max())
        assert rel_inf(U[i][None], k.get_state()[1][None]) <= 1e-6
    c.set_t(None)  # back to the uniform clock
    c.step_closed_loop(1)
    c.close()


def test_concurrent_onchip_handles_share_tensor_memory(cg):
    """Two on-chip handles of different models stepping concurrently on their own streams: their CTAs can be
    co-resident on an SM and both allocate tensor memory (tcgen05.alloc blocks until columns are free).  Results
    must equal the stand-alone runs; the test would hang (pytest-timeout / driver limit) on an allocation deadlock."""
    n1, n2, steps = 3000, 5000, 40
    x1, p1, u1 = po.synthetic_batch(po.MSD, n1, seed=21)
    x2, p2, u2 = po.synthetic_batch(po.SEMIACTIVE, n2, seed=22)
    a1, _ = make(cg, po.MSD, x1, p1, u1, mode=cg.MODE_ONCHIP_EXACT)
    a2, _ = make(cg, po.SEMIACTIVE, x2, p2, u2, mode=cg.MODE_ONCHIP_EXACT)
    for _ in range(steps):
        a1.step_closed_loop(1)
        a2.step_closed_loop(1)
    xa, xb = a1.get_x(), a2.get_x()
    b1, _ = make(cg, po.MSD, x1, p1, u1, mode=cg.MODE_ONCHIP_EXACT)
    b1.step_closed_loop(steps)
    b2, _ = make(cg, po.SEMIACTIVE, x2, p2, u2, mode=cg.MODE_ONCHIP_EXACT)
    b2.step_closed_loop(steps)
    assert np.array_equal(xa, b1.get_x()) and np.array_equal(xb, b2.get_x())
    for c in (a1, a2, b1, b2):
        c.close()


@pytest.mark.parametrize("mode_name", ["MODE_EXACT", "MODE_ONCHIP_EXACT"])
def test_device_pointer_control_on_a_caller_stream(cg, oracle_best, mode_name):
    """cgmres_b200_control_dev: x and u stay in device memory (instance-major), the handle runs on the caller's
    stream (cgmres_b200_set_stream) -- here torch tensors on a torch stream; results bit-identical to the oracle."""
    import torch

    mode = getattr(cg, mode_name)
    model, n, steps = po.MSD, 70, 25
    x0, p, u0 = po.synthetic_batch(model, n, seed=41)
    want = oracle_best.run_closed_loop(model, x0, p, u0, steps, want_U=True)
    c, _ = make(cg, model, x0, p, u0, mode=mode)
    stream = torch.cuda.Stream()
    c.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        xd = torch.from_numpy(x0).cuda()
        ud = torch.empty((n, c.dim_u), dtype=torch.float64, device="cuda")
        dt = 0.001
        for _ in range(steps):
            c.control_dev(ud.data_ptr(), xd.data_ptr())
            # plant on the device in the reference's operation order (mass_spring_damper/simulator.hpp:13-18)
            f2 = (((-1.0 * xd[:, 0] + 1.0 * xd[:, 1]) - 2.0 * xd[:, 2]) + 1.0 * xd[:, 3]) + ud[:, 0]
            f3 = (((1.0 * xd[:, 0] - 1.0 * xd[:, 1]) + 1.0 * xd[:, 2]) - 1.0 * xd[:, 3]) + ud[:, 1]
            xn = torch.stack([xd[:, 0] + xd[:, 2] * dt, xd[:, 1] + xd[:, 3] * dt, xd[:, 2] + f2 * dt,
                              xd[:, 3] + f3 * dt], dim=1)
            xd = xn.contiguous()
        stream.synchronize()
    assert np.array_equal(xd.cpu().numpy(), want["x_fin"])
    assert np.array_equal(c.get_state()[1], want["U_fin"])
    c.set_stream(None)
    c.close()
