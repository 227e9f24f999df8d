#!/usr/bin/env python
"""Phase breakdown of one warp of CTA 0 of the first-generation on-chip kernel (fast_update.cuh; library built with
EXTRA=-DCG_FAST_TIMING; it now serves MODE_ONCHIP_EXACT).  For the pipelined fast-mode kernel see
tools/pipe_wait_times.py (FAST_EXTRA=-DCG_PIPE_TIMING)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200._lib import check, lib  # noqa: E402
from cgmres_cpp_b200 import workloads as po  # noqa: E402

model = {"msd": 0, "arm": 1, "semiactive": 2}[sys.argv[1] if len(sys.argv) > 1 else "msd"]
n = 65536
x0, p, u0 = po.synthetic_batch(model, n)
c = cg.BatchedCgmres(model, n, mode=cg.MODE_ONCHIP_EXACT)
c.set_ptau_repeat(p); c.init_u0(u0); c.init_u0_newton(u0, x0, p, 10); c.set_x(x0)
c.step_closed_loop(20)
c.synchronize()
t = np.zeros(64, dtype=np.int64)
check(lib().cgmres_b200_debug_phase_times(c._h, C.c_void_p(t.ctypes.data)))
names = {6: "  (state in: global loads -> smem issued)", 7: "  (state in: first barrier)",
         41: "  (epilogue: back substitution)", 42: "  (epilogue: -)", 43: "  (epilogue: V*y from TMEM + U/dUdt read-modify-write)",
         1: "alloc+start", 2: "state in, x+dx*h", 3: "first evaluation (3 trajectories, serial lanes)",
         5: "b, r0, ||r0||, v0", 39: "(end of Arnoldi)", 40: "back-subst, V*y, U update, state out"}
for k in range(5):
    names[10 + 5 * k + 1] = f"k={k}: X = U + h*v"
    names[10 + 5 * k + 2] = f"k={k}: serial rollout+costates (waiting at barriers)"
    names[10 + 5 * k + 3] = f"k={k}: stage-parallel dHdu, w"
    names[10 + 5 * k + 4] = f"k={k}: Gram-Schmidt + norm + v store"
    names[10 + 5 * k + 5] = f"k={k}: Householder / residual scalars"
for k in range(5):
    if t[50 + 2 * k] and t[51 + 2 * k]:
        print(f"serial warp, k={k}: rollout + costate recursion of 16 lanes = {t[51 + 2 * k] - t[50 + 2 * k]} cycles")
if t[60] and t[61]:
    print(f"warp 0: tcgen05.alloc + relinquish = {t[61] - t[60]} cycles; CTA start to first mark of warp 5 = {t[1] - t[60]} cycles")
marks = sorted([i for i in range(50) if t[i] != 0], key=lambda i: t[i])
total = t[marks[-1]] - t[marks[0]]
prev = marks[0]
agg = {}
for i in marks[1:]:
    d = t[i] - t[prev]
    label = names.get(i, f"mark {i}")
    print(f"{d:8d} cycles {100.0 * d / total:5.1f}%  {label}")
    key = label.split(": ")[-1] if label.startswith("k=") else label
    agg[key] = agg.get(key, 0) + d
    prev = i
print(f"{total:8d} cycles total (one 16-instance round)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"   {100.0 * v / total:5.1f}%  {k}")
