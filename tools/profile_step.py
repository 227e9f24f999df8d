#!/usr/bin/env python
"""Minimal driver for ncu: builds one batch and advances it a few closed-loop steps (no torch, no timing)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200 import workloads as po  # noqa: E402  (seeded synthetic inputs)

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="msd", choices=("msd", "arm", "semiactive"))
ap.add_argument("--mode", default="fast", choices=("exact", "fast", "onchip_exact", "pipelined_exact"))
ap.add_argument("--instances", type=int, default=0)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--per-step", action="store_true", help="one launch per step instead of one multi-step launch")
a = ap.parse_args()
mid = {"msd": 0, "arm": 1, "semiactive": 2}[a.model]
n = a.instances or {"msd": 65536, "arm": 262144, "semiactive": 131072}[a.model]
x0, p, u0 = po.synthetic_batch(mid, n)
c = cg.BatchedCgmres(mid, n, mode={"exact": cg.MODE_EXACT, "fast": cg.MODE_FAST, "onchip_exact": cg.MODE_ONCHIP_EXACT,
                                       "pipelined_exact": cg.MODE_PIPELINED_EXACT}[a.mode])
c.set_ptau_repeat(p)
c.init_u0(u0)
c.init_u0_newton(u0, x0, p, 10)
c.set_x(x0)
if a.per_step:
    for _ in range(a.steps):
        c.step_closed_loop(1)
else:
    c.step_closed_loop(a.steps)
c.synchronize()
print("ok", a.model, a.mode, n, a.steps, float(abs(c.get_x()).max()))
