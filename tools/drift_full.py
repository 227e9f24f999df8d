#!/usr/bin/env python
"""Full-size closed-loop drift of the fast mode against the exact mode (which is bit-identical to the reference):
max |x_fast - x_exact| over all instances and steps of the BASELINE batch."""
import argparse, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgmres_cpp_b200 as cg
from cgmres_cpp_b200 import workloads as po

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="msd")
ap.add_argument("--instances", type=int, default=65536)
ap.add_argument("--steps", type=int, default=1000)
a = ap.parse_args()
mid = {"msd": 0, "arm": 1, "semiactive": 2}[a.model]
x0, p, u0 = po.synthetic_batch(mid, a.instances)
ctl = []
for mode in (cg.MODE_EXACT, cg.MODE_FAST):
    c = cg.BatchedCgmres(mid, a.instances, mode=mode)
    c.set_ptau_repeat(p); c.init_u0(u0); c.init_u0_newton(u0, x0, p, 10); c.set_x(x0)
    ctl.append(c)
worst = np.zeros(a.instances)
for r in range(a.steps // 50):
    for c in ctl:
        c.step_closed_loop(50)
    d = np.abs(ctl[0].get_x() - ctl[1].get_x()).max(axis=1)
    worst = np.maximum(worst, d)
print(json.dumps({"model": a.model, "instances": a.instances, "steps": a.steps, "max_abs_dx": float(worst.max()),
                  "p99_abs_dx": float(np.quantile(worst, 0.99)), "median_abs_dx": float(np.median(worst)),
                  "n_above_1e-6": int((worst > 1e-6).sum()), "finite": bool(np.isfinite(worst).all())}))
