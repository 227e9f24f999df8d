#!/usr/bin/env python
"""Throughput of every build mode on the benchmark batch of one model: per-step launches and (where the mode
supports it) multi-step launches.  usage: mode_rates.py [msd|arm|semiactive] [n] [steps] [modes,comma,separated]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402  (CUDA events)

import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200 import workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "msd"
model = {"msd": 0, "arm": 1, "semiactive": 2}[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else {"msd": 65536, "arm": 262144, "semiactive": 131072}[name]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["fast", "pipelined_exact", "onchip_exact", "exact"]
ids = {"exact": 0, "fast": 1, "onchip_exact": 2, "pipelined_exact": 3}
x0, p, u0 = workloads.synthetic_batch(model, n)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ref = None
for m in modes:
    for fused in (False, True):
        c = cg.BatchedCgmres(model, n, mode=ids[m])
        c.set_stream(stream.cuda_stream)
        c.set_ptau_repeat(p)
        c.init_u0(u0)
        c.init_u0_newton(u0, x0, p, 10)
        c.set_x(x0)
        c.step_closed_loop(10)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if fused:
            c.step_closed_loop(steps)
        else:
            for _ in range(steps):
                c.step_closed_loop(1)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        x = c.get_x()
        if ref is None:
            ref = x
        print(f"{name} n={n} {m:16s} {'one call' if fused else 'per-step':9s} {ms:8.4f} ms/step {n/ms*1e3:.4e} updates/s "
              f"max|dx| vs first = {np.abs(x-ref).max():.3e} finite={np.isfinite(x).all()}", flush=True)
        c.close()
