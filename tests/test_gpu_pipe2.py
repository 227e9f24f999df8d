"""GPU tests of what the third-generation persistent kernel (cgmres_cpp_b200/csrc/pipe2_update.cuh) adds on top of
the parity tests in test_gpu_onchip.py (which already run it: MODE_PIPELINED_EXACT, and MODE_FAST for large batches):

  * multi-step launches: step_closed_loop(k) = ONE launch advancing the same resident instances k steps must equal k
    single-step launches bit for bit (both build modes; ragged, multi-round and small batches; per-instance clocks);
  * the device-side trajectory log (cgmres_b200_step_closed_loop_log) in every mode against per-step get_x / get_u,
    and against the oracle's recorded trajectory;
  * the bit-exact build against the CPU oracle over a multi-step launch that crosses every GMRES exit path.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from test_gpu_parity import make

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg(built):
    import cgmres_cpp_b200 as m

    if m.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return m


FUSED_MODES = ["MODE_PIPELINED_EXACT", "MODE_FAST"]


@pytest.mark.parametrize("mode_name", FUSED_MODES)
@pytest.mark.parametrize("model,n", [(po.MSD, 1), (po.MSD, 45), (po.MSD, 148 * 32 * 2 + 77), (po.SEMIACTIVE, 333),
                                     (po.ARM, 97)])
def test_multi_step_launch_equals_single_step_launches(cg, model, n, mode_name):
    mode = getattr(cg, mode_name)
    x0, p, u0 = po.synthetic_batch(model, n, seed=5 + n)
    a, _ = make(cg, model, x0, p, u0, mode=mode)
    b, _ = make(cg, model, x0, p, u0, mode=mode)
    steps = 23
    a.step_closed_loop(steps)  # one launch
    for _ in range(steps):
        b.step_closed_loop(1)
    if mode == cg.MODE_FAST and n <= 16 * 148:
        # single-step launches of small batches take the first-generation kernel (shorter chain), the multi-step
        # launch the persistent one: same algorithm, FMA contraction may differ in the last bit
        assert np.abs(a.get_x() - b.get_x()).max() <= 1e-9
    else:
        ta, Ua, dUa = a.get_state()
        tb, Ub, dUb = b.get_state()
        assert ta == tb
        assert np.array_equal(a.get_x(), b.get_x()) and np.array_equal(a.get_u(), b.get_u())
        assert np.array_equal(Ua, Ub) and np.array_equal(dUa, dUb)
        assert np.array_equal(a.get_status()[0], b.get_status()[0])
    a.close()
    b.close()


def test_multi_step_exact_launch_matches_oracle_through_exit_paths(cg, oracle_port):
    """The shipped msd run from step 11,800: early convergences (k = 0..4) and rho0 < tol returns all occur in the next
    2,400 steps.  One fused launch per 256 steps, bit-identical U, dUdt, x to the C oracle."""
    s = po.SHIPPED[po.MSD]
    x0, pp, u0 = np.array([s["x0"]]), np.array([s["p"]]), np.array(s["u0"])
    c = oracle_port.controller(po.MSD)
    c.set_ptau_repeat(pp[0])
    u = u0.copy()
    c.init_u0(u)
    u = c.init_u0_newton(u, x0[0], pp[0], 10)
    x = x0[0].copy()
    for _ in range(11800):
        uu = c.control(x)
        oracle_port.plant_step(po.MSD, x, uu)
    t, U, dUdt = c.get_state()
    g = cg.BatchedCgmres(po.MSD, 1, mode=cg.MODE_PIPELINED_EXACT)
    g.set_ptau_repeat(pp)
    g.set_state(t, U[None], dUdt[None])
    g.set_x(x[None])
    seen = set()
    for _ in range(6):
        g.step_closed_loop(400)  # 256 + 144 steps: two fused launches
        for _ in range(400):
            uu = c.control(x)
            oracle_port.plant_step(po.MSD, x, uu)
            seen.add(c.last_status()[0])
        tg, Ug, dUg = g.get_state()
        to, Uo, dUo = c.get_state()
        assert np.array_equal(g.get_x()[0], x)
        assert np.array_equal(Ug[0], Uo) and np.array_equal(dUg[0], dUo)
        assert tg == to
    assert {0, 1, 2} <= seen  # full, early convergence, rho0 < tol all happened inside fused launches
    g.close()


@pytest.mark.parametrize("mode_name", ["MODE_EXACT", "MODE_ONCHIP_EXACT", "MODE_PIPELINED_EXACT", "MODE_FAST"])
def test_device_trajectory_log_matches_per_step_readback(cg, mode_name):
    mode = getattr(cg, mode_name)
    model, n, steps = po.MSD, 300, 31
    x0, p, u0 = po.synthetic_batch(model, n, seed=11)
    a, _ = make(cg, model, x0, p, u0, mode=mode)
    b, _ = make(cg, model, x0, p, u0, mode=mode)
    xl, ul = a.step_closed_loop_log(steps)
    assert xl.shape == (steps, n, a.dim_x) and ul.shape == (steps, n, a.dim_u)
    for s in range(steps):
        b.step_closed_loop(1)
        xs, us = b.get_x(), b.get_u()
        if mode == cg.MODE_FAST:  # logged run = persistent kernel, per-step small batch = first-generation kernel
            assert np.abs(xl[s] - xs).max() <= 1e-9 and np.abs(ul[s] - us).max() <= 1e-7
        else:
            assert np.array_equal(xl[s], xs), s
            assert np.array_equal(ul[s], us), s
    assert np.array_equal(a.get_x(), xl[-1])
    a.close()
    b.close()


def test_device_trajectory_log_reproduces_the_reference_trajectory(cg, oracle_best):
    """The logged rows are what the reference's main.cpp prints: semiactive shipped run, first 700 steps, bit for bit."""
    s = po.SHIPPED[po.SEMIACTIVE]
    x0, pp, u0 = np.array([s["x0"]]), np.zeros((1, 0)), np.array(s["u0"])
    steps = 700
    want = oracle_best.run_closed_loop(po.SEMIACTIVE, x0, pp, u0, steps, rec_stride=1)
    c, _ = make(cg, po.SEMIACTIVE, x0, pp, u0, mode=cg.MODE_PIPELINED_EXACT)
    xl, ul = c.step_closed_loop_log(steps)
    assert np.array_equal(xl[:, 0, :], want["x_traj"][:, 0, :])
    assert np.array_equal(ul[:, 0, :], want["u_traj"][:, 0, :])
    c.close()


def test_multi_step_launch_with_per_instance_clocks(cg):
    """Per-instance controller clocks (device-evaluated horizon ramp) inside a fused launch equal per-step launches."""
    model, n = po.MSD, 70
    x0, p, u0 = po.synthetic_batch(model, n, seed=3)
    t0 = np.linspace(0.0, 0.02, n)
    a, _ = make(cg, model, x0, p, u0, mode=cg.MODE_PIPELINED_EXACT)
    b, _ = make(cg, model, x0, p, u0, mode=cg.MODE_PIPELINED_EXACT)
    a.set_t(t0)
    b.set_t(t0)
    a.step_closed_loop(12)
    for _ in range(12):
        b.step_closed_loop(1)
    assert np.array_equal(a.get_x(), b.get_x())
    assert np.array_equal(a.get_t(), b.get_t())
    a.close()
    b.close()
