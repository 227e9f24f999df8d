"""Closed-loop parity at BASELINE.json's FULL batch sizes, 1000 steps, every instance held to the bar.

The small-batch tests compare with the CPU oracle directly; a 65,536 x 1000-step batch is ~30 core-minutes on the
CPU, so here the chain is:  candidate mode  --(every instance, max|dx| <= 1e-6)-->  bit-exact GPU mode
--(bit for bit on a strided sample of the same batch)-->  compiled reference.
`bench.py` runs the same comparison live and only lets a mode with ZERO instances above the bar carry the headline
(bench.choose_headline); this test asserts that policy's outcome per model.
"""
import numpy as np
import pytest

import bench
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

STEPS = 1000
FULL = {"msd": 65536, "semiactive": 131072, "arm": 262144}  # BASELINE configs 2, 4 (per GPU), 3
SAMPLE = 192  # instances of the full batch re-run on the CPU reference (strided over the whole index range)


@pytest.fixture(scope="module")
def cg(built):
    import cgmres_cpp_b200 as m

    if m.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return m


def make_controller(cg, model, mode, x0, p, u0):
    c = cg.BatchedCgmres(bench.MODELS[model], x0.shape[0], device=0, mode=bench.MODE_IDS[mode])
    c.set_ptau_repeat(p)
    c.init_u0(u0)
    c.init_u0_newton(u0, x0, p, 10)
    c.set_x(x0)
    return c


def run_mode(cg, model, mode, x0, p, u0, steps):
    with make_controller(cg, model, mode, x0, p, u0) as c:
        c.step_closed_loop(steps)
        x = c.get_x()
        code, _ = c.get_status()
    return x, code


@pytest.mark.parametrize("model", ["msd", "semiactive", "arm"])
def test_default_mode_full_batch_1000_steps(cg, oracle_best, model):
    from cgmres_cpp_b200 import workloads

    mid, n = bench.MODELS[model], FULL[model]
    x0, p, u0 = workloads.synthetic_batch(mid, n, seed=12345)  # the benchmark's batch
    cand = bench.CANDIDATES[model]
    anchor = next(m for m in cand if m in bench.BIT_EXACT_MODES)
    # every candidate against the anchor at EVERY one of the 1000 steps (device-side trajectory log, chunked)
    drift, ends = bench.trajectory_drift(lambda m: make_controller(cg, model, m, x0, p, u0), cand, anchor, STEPS, n,
                                         bench.MODEL_ROW_DOUBLES[model])
    assert all(np.isfinite(v).all() for v in ends.values())

    # (1) the anchor is the reference: a strided sample of the SAME batch through the compiled reference on the CPU
    idx = np.linspace(0, n - 1, SAMPLE).astype(np.int64)
    want = oracle_best.run_closed_loop(mid, x0[idx], p[idx] if p.shape[1] else p[idx], u0, STEPS, n_threads=8)
    got = ends[anchor][idx]
    if model == "arm":  # libm sin/cos in the reference, portable sin/cos on the device: tolerance bar
        assert np.abs(got - want["x_fin"]).max() <= bench.CLOSED_LOOP_BAR
    else:
        assert np.array_equal(got, want["x_fin"]), float(np.abs(got - want["x_fin"]).max())

    # (2) every bit-exact mode agrees with the anchor bit for bit on all n instances
    for m in cand:
        if m in bench.BIT_EXACT_MODES:
            assert np.array_equal(ends[m], ends[anchor]), m

    # (3) the mode the benchmark would pick has zero instances above the bar -- on EVERY instance of the full batch, at
    #     every step of the closed loop
    stats = {m: bench.drift_stats(drift[m]) for m in cand}
    head = bench.choose_headline({m: (0.0 if m == "fast" else 1.0, stats[m]["n_above_bar"]) for m in cand}, anchor)
    assert stats[head]["n_above_bar"] == 0
    assert stats[head]["max_abs_dx"] <= bench.CLOSED_LOOP_BAR
    # a mode that misses the bar on any instance must never be chosen, however fast it is
    for m in cand:
        if stats[m]["n_above_bar"] > 0:
            assert head != m
    print(model, {m: (stats[m]["n_above_bar"], stats[m]["max_abs_dx"]) for m in cand}, "headline:", head)


def test_fast_mode_full_batch_known_miss_is_measured(cg):
    """The FMA + shuffle-sum mode on the msd benchmark batch: the closed loop amplifies its 1e-16-level reordering
    noise past 1e-6 on a handful of the 65,536 instances.  This records the count; it is a
    KNOWN MISS of that mode (xfail), which is why bench.py never reports it as the headline while the count is > 0."""
    from cgmres_cpp_b200 import workloads

    n = FULL["msd"]
    x0, p, u0 = workloads.synthetic_batch(po.MSD, n, seed=12345)
    drift, _ = bench.trajectory_drift(lambda m: make_controller(cg, "msd", m, x0, p, u0), ["fast"], "onchip_exact",
                                      STEPS, n, bench.MODEL_ROW_DOUBLES["msd"])
    st = bench.drift_stats(drift["fast"])
    print("fast vs bit-exact, msd 65536 x 1000, max over every step:", st)
    assert st["p99_abs_dx"] <= bench.CLOSED_LOOP_BAR  # the bulk is far inside the bar
    if st["n_above_bar"] > 0:
        pytest.xfail(f"fast mode: {st['n_above_bar']} of {n} instances above 1e-6 (max {st['max_abs_dx']:.2e})")
