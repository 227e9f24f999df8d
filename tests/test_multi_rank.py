"""CPU tests of the N>1 path (world_size 2, gloo): shard assignment, max-over-ranks timing, rank-0-only output.
Instances never interact (multiple_controller/main.cpp:107-118), so shard g of a G-way run must equal the same
rows of a single run bit for bit; there is no data-path collective to test."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def torchrun(script_args, nproc=2, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_shard_ranges():
    from cgmres_cpp_b200.sharding import shard_range, weak_scaling_range

    for total, world in ((65536, 8), (10, 3), (0, 2), (7, 8)):
        r = [shard_range(total, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert weak_scaling_range(65536, 8, 3) == (3 * 65536, 4 * 65536)


def test_two_ranks_gloo_shards_equal_single_run(built):
    from oracle import pyoracle as po

    r = torchrun([os.path.join("tests", "_dist_worker.py")])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["ranges"] == [[0, 6], [6, 12]]
    assert d["max_ms"] == [15.0, 3.0]
    assert d["value"] == 12 * 40 / 15e-3
    x0, p, u0 = po.synthetic_batch(po.MSD, 12, seed=12345)
    want = po.load("port").run_closed_loop(po.MSD, x0, p, u0, 40)
    assert np.array_equal(np.array(d["x_fin"]), want["x_fin"])


def test_reference_arm_prints_one_line_from_rank0_only(built):
    r = torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "20", "--warmup", "2",
                  "--cpu-instances-per-core", "1"], port=29612)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "updates/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
