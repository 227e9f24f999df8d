"""Worker of tests/test_multi_rank.py: launched by torch.distributed.run with world_size 2, gloo backend."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgmres_cpp_b200.sharding import aggregate_updates_per_second, max_over_ranks, weak_scaling_range  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
per_gpu, steps = 6, 40
lo, hi = weak_scaling_range(per_gpu, world, rank)
ranges = [None] * world
dist.all_gather_object(ranges, (lo, hi))
# the rank's shard of the global seeded batch, run on the CPU oracle (stand-in for the device in this CPU test)
x0, p, u0 = po.synthetic_batch(po.MSD, per_gpu * world, seed=12345)
out = po.load("port").run_closed_loop(po.MSD, x0[lo:hi], p[lo:hi], u0, steps)
gathered = [None] * world
dist.all_gather_object(gathered, out["x_fin"].tolist())
dist.barrier()
ms = max_over_ranks([10.0 + 5.0 * rank, 3.0 - rank], dist)
if rank == 0:
    print(json.dumps({"ranges": ranges, "x_fin": sum(gathered, []), "max_ms": ms,
                      "value": aggregate_updates_per_second(per_gpu, world, steps, ms[0])}))
dist.destroy_process_group()
