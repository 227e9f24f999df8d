#!/usr/bin/env python
"""Full-size closed-loop drift of one build mode against a bit-exact mode: max over EVERY step and every state component
of |x_mode - x_exact|, per instance, on the BASELINE batch (the measurement bench.py runs live; see
bench.trajectory_drift)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200 import workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="msd", choices=tuple(bench.MODELS))
ap.add_argument("--mode", default="fast", choices=tuple(bench.MODE_IDS))
ap.add_argument("--anchor", default="", help="bit-exact mode to compare with (default: the model's first bit-exact candidate)")
ap.add_argument("--instances", type=int, default=0)
ap.add_argument("--steps", type=int, default=1000)
a = ap.parse_args()
mid = bench.MODELS[a.model]
n = a.instances or bench.INSTANCES_PER_GPU[a.model]
anchor = a.anchor or next(m for m in bench.CANDIDATES[a.model] if m in bench.BIT_EXACT_MODES)
x0, p, u0 = workloads.synthetic_batch(mid, n)


def make(mode):
    c = cg.BatchedCgmres(mid, n, mode=bench.MODE_IDS[mode])
    c.set_ptau_repeat(p)
    c.init_u0(u0)
    c.init_u0_newton(u0, x0, p, 10)
    c.set_x(x0)
    return c


drift, _ = bench.trajectory_drift(make, [a.mode], anchor, a.steps, n, bench.MODEL_ROW_DOUBLES[a.model])
out = bench.drift_stats(drift[a.mode])
out.update({"model": a.model, "mode": a.mode, "anchor": anchor, "steps": a.steps})
print(json.dumps(out))
