#!/usr/bin/env python
"""End-to-end (host-buffer) loop only, under torchrun: u = control(x) through pinned host buffers + the plant step on
the host, per-rank and max-over-ranks ms per step.  OpenMP settings come from the environment of the launch, e.g.
  OMP_NUM_THREADS=4 OMP_WAIT_POLICY=active python -m torch.distributed.run --nproc-per-node 8 tools/e2e_scaling.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cgmres_cpp_b200 as cg  # noqa: E402
from cgmres_cpp_b200 import workloads  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mode = {"fast": cg.MODE_FAST, "pipelined_exact": cg.MODE_PIPELINED_EXACT}[sys.argv[1] if len(sys.argv) > 1 else "fast"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
n = 65536
x0, p, u0 = workloads.synthetic_batch(cg.MSD, n * world)
x0, p = x0[rank * n:(rank + 1) * n], p[rank * n:(rank + 1) * n]
c = cg.BatchedCgmres(cg.MSD, n, device=local, mode=mode)
c.set_ptau_repeat(p); c.init_u0(u0); c.init_u0_newton(u0, x0, p, 10)
xh = torch.from_numpy(x0.copy()).pin_memory()
uh = torch.empty((n, c.dim_u), dtype=torch.float64).pin_memory()
xn, un = xh.numpy(), uh.numpy()
t_ctl = t_plant = 0.0
for it in range(steps + 5):
    if it == 5:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_ctl = t_plant = 0.0
        t0 = time.perf_counter()
    a = time.perf_counter()
    c.control_raw(uh.data_ptr(), xh.data_ptr())
    b = time.perf_counter()
    cg.plant_step_host(cg.MSD, xn, un)
    t_ctl += b - a
    t_plant += time.perf_counter() - b
tot = time.perf_counter() - t0
v = torch.tensor([tot, t_ctl, t_plant], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world={world} OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')} OMP_WAIT_POLICY={os.environ.get('OMP_WAIT_POLICY')} "
          f"HOST_THREADS={os.environ.get('CGMRES_B200_HOST_THREADS')}: "
          f"{v[0].item()/steps*1e3:.3f} ms/step (control {v[1].item()/steps*1e3:.3f}, plant {v[2].item()/steps*1e3:.3f}) "
          f"-> {n*world*steps/v[0].item():.3e} updates/s", flush=True)
if world > 1:
    dist.destroy_process_group()
