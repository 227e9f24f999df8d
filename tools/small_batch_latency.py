#!/usr/bin/env python
"""Per-step latency of small batches (closed loop on the device, one launch per step), all build modes."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200.workloads import synthetic_batch  # noqa: E402

model = {"msd": 0, "arm": 1, "semiactive": 2}[sys.argv[1] if len(sys.argv) > 1 else "msd"]
modes = (("fast", cg.MODE_FAST), ("onchip_exact", cg.MODE_ONCHIP_EXACT), ("pipelined_exact", cg.MODE_PIPELINED_EXACT),
         ("exact", cg.MODE_EXACT))
for n in (1, 64, 1024, 2368, 4736, 9472):
    x0, p, u0 = synthetic_batch(model, n, seed=5)
    row = []
    for name, mode in modes:
        c = cg.BatchedCgmres(model, n=n, device=0, mode=mode)
        if c.dim_p:
            c.set_ptau_repeat(p)
        c.init_u0(u0); c.init_u0_newton(u0, x0, p, 10); c.set_x(x0)
        c.step_closed_loop(20); c.synchronize()
        t = time.perf_counter(); c.step_closed_loop(300); c.synchronize()
        row.append("%s %.1f us" % (name, (time.perf_counter() - t) / 300 * 1e6))
        c.close()
    print("n = %5d: " % n + ", ".join(row), flush=True)
