#!/usr/bin/env python
"""Benchmark of the batched C/GMRES control-update hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model msd|arm|semiactive]

One "step" = one closed-loop step of the whole batch: Cgmres::control (include/cgmres.hpp:78-110 of the
reference) plus the forward-Euler plant step (<example>/main.cpp:74-76) for every instance = ONE kernel launch
per GPU.  Workload at N=1: BASELINE.json configs[1], mass_spring_damper, 65,536 instances (per GPU: the batch
shards with no collective, weak scaling).  `value` = control updates/s with the state resident in HBM;
`e2e` = the same through cgmres_b200_control() with HOST buffers (x H2D, u D2H every step).

Prints ONE JSON line on rank 0.  `--impl reference` times the reference's own CPU implementation
(oracle/_ref = the unmodified reference headers compiled with its own flags; the C port if that library
did not travel) on all host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Host threads of the e2e loop's plant step (cgmres_b200_plant_step_host, OpenMP): under torchrun every rank gets
# its share of the box's cores instead of torchrun's blanket OMP_NUM_THREADS=1.  Must happen before libgomp loads.
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    _lws = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, _lws)))

METRIC = "batched C/GMRES control updates/sec"
UNIT = "updates/s"
MODELS = {"msd": 0, "arm": 1, "semiactive": 2}
INSTANCES_PER_GPU = {"msd": 65536, "arm": 262144, "semiactive": 131072}
# algorithmic work per update (full k_max=5 iterations, 8 F evaluations)
FLOP_PER_UPDATE = {"msd": 72005, "arm": 30432, "semiactive": 32905}
HBM_BYTES_PER_UPDATE = {"msd": 9728, "arm": 2504, "semiactive": 4856}  # read U,dUdt,x,p; write U,dUdt,x,u
# fastest mode per model that meets the parity bars: the on-chip TMEM kernel for the models
# whose time is in the Krylov vector work, the streaming thread-per-instance kernel for the sin/cos-heavy arm model
DEFAULT_MODE = {"msd": "fast", "semiactive": "fast", "arm": "exact"}
MODE_IDS = {"exact": 0, "fast": 1, "onchip_exact": 2, "pipelined_exact": 3}


def workload_name(model: str, n_per_gpu: int, steps: int) -> str:
    full = {"msd": "mass_spring_damper", "arm": "arm_type_inverted_pendulum", "semiactive": "semiactive_damper"}[model]
    return f"{full} batched {n_per_gpu} instances per GPU, {steps} closed-loop steps (control + Euler plant step)"


# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (nvidia-smi recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop = index, [], threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self._stop.is_set():
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        self._stop.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def visible_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------------------------
def cpu_baseline_run(model_id: int, n_inst: int, steps: int, threads: int, seed: int = 12345):
    """The reference's CPU path on `threads` host threads over the first n_inst instances of the GPU workload."""
    from oracle import pyoracle as po

    ora = po.best()
    x0, p, u0 = po.synthetic_batch(model_id, n_inst, seed=seed)
    t0 = time.perf_counter()
    out = ora.run_closed_loop(model_id, x0, p, u0, steps, n_threads=threads)
    wall = time.perf_counter() - t0
    lat = out["ctl_seconds"] / max(steps, 1)
    return {"kind": ora.kind, "updates": n_inst * steps, "wall_s": wall,
            "p50_control_us": float(np.median(lat) * 1e6) if n_inst else None}


def run_reference_arm(args, rank: int):
    """--impl reference: rank 0 alone measures; the other ranks exit 0 without work."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    model_id = MODELS[args.model]
    n_per_gpu = args.instances or INSTANCES_PER_GPU[args.model]
    # bounded sample: a few instances per core, every step advances all of them by one closed-loop step
    n_inst = min(n_per_gpu, cores * max(1, args.cpu_instances_per_core))
    if args.warmup > 0:
        cpu_baseline_run(model_id, n_inst, args.warmup, cores)
    r = cpu_baseline_run(model_id, n_inst, args.steps, cores)
    value = r["updates"] / r["wall_s"]
    sample = (f"first {n_inst} instances of the seeded {n_per_gpu}-instance batch x {args.steps} closed-loop steps, "
              f"{cores} host threads, one live controller per thread")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["wall_s"] / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.model, n_per_gpu, args.steps), "sample_instances": n_inst,
                   "host_threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": sample,
                         "per_core": value / cores, "p50_control_latency_us": r["p50_control_us"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch

    import cgmres_cpp_b200 as cg
    from cgmres_cpp_b200 import workloads  # seeded synthetic inputs (oracle/ is only touched by the CPU-baseline leg)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    model, model_id = args.model, MODELS[args.model]
    if args.mode == "auto":
        args.mode = DEFAULT_MODE[args.model]
    mode = MODE_IDS[args.mode]
    n = args.instances or INSTANCES_PER_GPU[model]
    # every rank owns a disjoint shard of one global seeded batch: rank r gets instances [r*n, (r+1)*n)
    from cgmres_cpp_b200.sharding import aggregate_updates_per_second, max_over_ranks, weak_scaling_range

    lo, hi = weak_scaling_range(n, world, rank)
    x0_all, p_all, u0 = workloads.synthetic_batch(model_id, n * world, seed=12345)
    x0, p = x0_all[lo:hi], p_all[lo:hi]

    ctl = cg.BatchedCgmres(model_id, n, device=local_rank, mode=mode)
    # the kernels are launched on this (non-default) stream and the CUDA events are recorded on the same one
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctl.set_stream(stream.cuda_stream)
    ctl.set_ptau_repeat(p)
    ctl.init_u0(u0)
    ctl.init_u0_newton(u0, x0, p, 10)
    ctl.set_x(x0)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident closed loop: `value` -------------------------------------------------------------
    ctl.step_closed_loop(args.warmup)
    barrier()
    sampler = ClockSampler(visible_gpu_index(local_rank))
    sampler.start()
    time.sleep(0.3)
    launches0 = cg.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(args.steps):
        ctl.step_closed_loop(1)
        ev[k + 1].record(stream)
    barrier()
    launches = cg.launch_count() - launches0
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    x_end = ctl.get_x()
    finite = bool(np.isfinite(x_end).all())
    code, _ = ctl.get_status()
    exit_hist = np.bincount(code, minlength=4).tolist()

    # ---- end to end through the host-buffer API: `e2e` ---------------------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    xh = torch.empty((n, ctl.dim_x), dtype=torch.float64).pin_memory()
    uh = torch.empty((n, ctl.dim_u), dtype=torch.float64).pin_memory()
    xh.copy_(torch.from_numpy(x_end))
    xn, un = xh.numpy(), uh.numpy()
    # the e2e loop is the reference's main(): u = control(x) through HOST buffers, then the plant step on the host
    # (cgmres_b200_plant_step_host = the Simulator functor of include/<example>/simulator.hpp, compiled host code)
    for _ in range(3):
        ctl.control_raw(uh.data_ptr(), xh.data_ptr())
        cg.plant_step_host(model_id, xn, un)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctl.control_raw(uh.data_ptr(), xh.data_ptr())  # H2D x, update kernel, D2H u, synchronises
        cg.plant_step_host(model_id, xn, un)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    # ---- the other build modes on the same batch (short runs, rank 0 reporting only) and a live parity probe ------
    other = {}
    if world == 1 and not args.no_other_modes:
        probe_n, probe_steps = min(n, 4096), 100
        ref_x = None
        for name in ("onchip_exact", "pipelined_exact", "exact", "fast"):
            c2 = cg.BatchedCgmres(model_id, n, device=local_rank, mode=MODE_IDS[name])
            c2.set_stream(stream.cuda_stream)
            c2.set_ptau_repeat(p)
            c2.init_u0(u0)
            c2.init_u0_newton(u0, x0, p, 10)
            c2.set_x(x0)
            c2.step_closed_loop(5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            c2.step_closed_loop(probe_steps - 5)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (probe_steps - 5)
            xs = c2.get_x()[:probe_n]
            if name == "onchip_exact":
                ref_x = xs
            other[name] = {"updates_per_s": n / (ms * 1e-3), "ms_per_step": ms,
                           "max_abs_dx_vs_bit_exact_mode_after_100_steps": float(np.abs(xs - ref_x).max())}
            c2.close()

    # ---- reduce over ranks (max time) -----------------------------------------------------------------------
    total_ms_max, e2e_ms_max = max_over_ranks([total_ms, e2e_s * 1e3], dist, device="cuda")
    ok = torch.tensor([1.0 if finite else 0.0], device="cuda")
    if dist is not None:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)

    if rank == 0:
        value = aggregate_updates_per_second(n, world, args.steps, total_ms_max)
        launch_ms = statistics.mean(per_launch_ms)
        p50_ms = statistics.median(per_launch_ms)
        peak_fma = cg.measure_fp64_peak(local_rank, True)
        peak_nofma = cg.measure_fp64_peak(local_rank, False)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"
        flops = FLOP_PER_UPDATE[model] * n / (launch_ms * 1e-3) / 1e12
        hbm = HBM_BYTES_PER_UPDATE[model] * n / (launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = prof.get(f"{model}_{args.mode}_{n}", {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(model, n, args.steps), "mode": args.mode, "instances_per_gpu": n,
                "instances_total": n * world, "parallelism": f"instance-sharded x{world}, no collective",
                "l2": "per-step working set exceeds the 126 MB L2 (state+scratch >= 1.5 GB per GPU); no flush needed",
                "inputs": "seeded synthetic x0/p of SURVEY 8(d), u0 shipped + init_u0_newton(10)",
            },
            "p50_launch_latency_ms": p50_ms,
            "p50_per_update_latency_us": p50_ms * 1e3 / n,
            "roofline": {
                "bound": "fp64", "achieved": flops, "peak": peak_fma, "unit": "TFLOP/s", "frac": flops / peak_fma,
                "traffic": traffic, "kernel": "%s::control_kernel (one launch = one control update + plant step per instance)" % {"exact": "exact", "onchip_exact": "fast", "fast": "pipe", "pipelined_exact": "pipe"}.get(args.mode, args.mode),
                "flop_per_update": FLOP_PER_UPDATE[model], "launch_ms": launch_ms,
                "peak_source": "measured live: 8 DFMA chains/thread microbenchmark (cgmres_b200_measure_fp64_peak)",
                "peak_no_fma": peak_nofma, "frac_of_no_fma_peak": flops / peak_nofma,
                "hbm": {"achieved": hbm, "peak": hbm_peak, "unit": "GB/s", "frac": hbm / hbm_peak,
                        "bytes_per_update": HBM_BYTES_PER_UPDATE[model], "peak_source": hbm_src},
            },
            "e2e": {"value": n * world * e2e_steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n * ctl.dim_x * 8, "d2h_bytes_per_step": n * ctl.dim_u * 8,
                    "steps": e2e_steps, "api": "cgmres_b200_control(u_host, x_host) + cgmres_b200_plant_step_host (the loop of the reference main.cpp)"},
            "gpu_launches": int(launches),
            "clocks": {k: clocks.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples", "power_w_max")},
            "finite": bool(ok.item() > 0.5), "exit_hist_last_step": exit_hist,
            "parity": {
                "mode": args.mode,
                "bars": "per update |dU|inf/|U|inf <= 1e-9 (teacher forced); closed loop max|dx| <= 1e-6 over 1000 steps",
                "exact / onchip_exact": "bit-identical U, dUdt, x, status to the reference (msd, semiactive); arm to the bars (libm sin/cos)",
                "fast": "per update <= 2e-15 rel; closed loop over 1000 steps on the full 65,536-instance msd batch: "
                        "median 2.6e-8, p99 2.8e-7, max 2.5e-6 (15 instances = 0.02 % above 1e-6); semiactive max 9.4e-8; "
                        "arm max 3.6e-7; see DESIGN.md section 3 and tools/drift_full.py",
            },
            "modes": other,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_cpu = min(n, cores * max(1, args.cpu_instances_per_core))
            cpu_steps = 1000  # the BASELINE closed-loop length, whatever --steps is: ~15 core-seconds of work
            r = cpu_baseline_run(model_id, n_cpu, cpu_steps, cores)
            v = r["updates"] / r["wall_s"]
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": r["kind"], "per_core": v / cores,
                "p50_control_latency_us": r["p50_control_us"], "wall_s": r["wall_s"],
                "sample": f"first {n_cpu} instances of the same seeded batch x {cpu_steps} closed-loop steps, "
                          f"{cores} host threads (one live controller per thread)"}
        print(json.dumps(line), flush=True)
    ctl.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--model", choices=tuple(MODELS), default="msd")
    ap.add_argument("--mode", choices=("auto", "fast", "onchip_exact", "pipelined_exact", "exact"), default="auto",
                    help="auto = the fastest parity-green mode of the model (DEFAULT_MODE)")
    ap.add_argument("--instances", type=int, default=0, help="instances per GPU (default: BASELINE config)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--cpu-instances-per-core", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the short runs of the other build modes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
