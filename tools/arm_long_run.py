#!/usr/bin/env python
"""The shipped arm run (10,001 steps) in every bit-exact mode against the C oracle built with the same portable
sin/cos: first step at which x or u differ (none expected)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200 import workloads  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

ic = workloads.SHIPPED[po.ARM]
steps = ic["steps"]
ora = po.load("port_ptrig")
x0, p, u0 = np.array([ic["x0"]]), np.array([ic["p"]]), np.array(ic["u0"])
want = ora.run_closed_loop(po.ARM, x0, p, u0, steps, rec_stride=1)
for name in ("MODE_EXACT", "MODE_ONCHIP_EXACT", "MODE_PIPELINED_EXACT"):
    c = cg.BatchedCgmres(po.ARM, 1, mode=getattr(cg, name))
    c.set_ptau_repeat(p)
    c.init_u0(u0)
    c.init_u0_newton(u0, x0, p, 10)
    c.set_x(x0)
    xl, ul = c.step_closed_loop_log(steps)
    dx = np.abs(xl[:, 0] - want["x_traj"][:, 0]).max(axis=1)
    du = np.abs(ul[:, 0] - want["u_traj"][:, 0]).max(axis=1)
    bad = np.nonzero((dx > 0) | (du > 0))[0]
    print(name, "first differing step:", (int(bad[0]) if bad.size else None), "max|dx|", dx.max(), "max|du|", du.max())
    if bad.size:
        s = int(bad[0])
        print("   around it: dx", dx[max(0, s - 2):s + 3], "du", du[max(0, s - 2):s + 3])
    c.close()
