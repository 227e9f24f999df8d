#!/usr/bin/env python
"""Blocked-vs-total cycles per warp of CTA 0 of the pipelined fast kernel (library built with
FAST_EXTRA=-DCG_PIPE_TIMING): warp 0 is the serial warp, the others are vector warps."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200._lib import check, lib  # noqa: E402
from cgmres_cpp_b200 import workloads as po  # noqa: E402

model = {"msd": 0, "arm": 1, "semiactive": 2}[sys.argv[1] if len(sys.argv) > 1 else "msd"]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
mode = {"fast": cg.MODE_FAST, "pipelined_exact": cg.MODE_PIPELINED_EXACT}[sys.argv[3] if len(sys.argv) > 3 else "fast"]
x0, p, u0 = po.synthetic_batch(model, n)
c = cg.BatchedCgmres(model, n, mode=mode)
if p is not None:
    c.set_ptau_repeat(p)
c.init_u0(u0); c.init_u0_newton(u0, x0, p, 10); c.set_x(x0)
c.step_closed_loop(20)
c.synchronize()
t = np.zeros(64, dtype=np.int64)
check(lib().cgmres_b200_debug_phase_times(c._h, C.c_void_p(t.ctypes.data)))
per = cg.instances_per_cta(model) if hasattr(cg, "instances_per_cta") else 0
for w in range(24):
    if t[2 * w + 1]:
        role = "serial" if w == 0 else "vector"
        print(f"warp {w:2d} ({role}): total {t[2*w+1]:8d} cycles, blocked at barriers {t[2*w]:8d} = {100.0*t[2*w]/t[2*w+1]:.1f} %")
if os.environ.get("CGMRES_B200_PIPE_GEN") != "2":
    print(f"serial warp 0: inside sweeps {t[48]} cycles, inside sequential sums {t[49]} (sum over its rounds)")
    print(f"first vector warp: stage-parallel dHdu {t[51]}, final updates {t[52]}, state in {t[53]}")
    print(f"first vector warp, phases incl. their dHdu: after sweep 1 {t[54]}, after sweep 2 {t[55]}, after sweep 3 {t[56]}, "
          f"Arnoldi phases {t[57]} (reflectors / v store / next input {t[58]}; bit-exact build: the phases between "
          f"the sequential sums {t[59]})")
    sys.exit(0)
if t[48]:
    print(f"serial warp 0: first pass {t[48]} cycles, second pass {t[49]}, Arnoldi sweeps {t[50]} (sum over its rounds)")
if t[51] or t[54]:
    names = ["final update (deferred)", "after pass 1: F1 -> TMEM, U + h*dUdt", "after pass 2: dHdu, b, r0, v0, U + h*v0",
             "Arnoldi: stage-parallel dHdu", "Arnoldi: (F - F1)/h", "Arnoldi: Gram-Schmidt + norm",
             "Arnoldi: reflectors, v store, flags", "Arnoldi: U + h*v (next sweep input)", "state in (next round)"]
    tot = sum(int(t[51 + i]) for i in range(9))
    for i, nm in enumerate(names):
        print(f"vector warp 0: {nm:45s} {t[51+i]:9d} cycles  {100.0*t[51+i]/tot:5.1f} %")
