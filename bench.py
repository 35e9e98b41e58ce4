#!/usr/bin/env python
"""bench.py -- SEPAIHRD Dopri5 + Poisson likelihood evaluations per second on N B200s.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # CPU arm (oracle on host cores)

Workload (BASELINE.json configs[1]): a batched likelihood sweep of
B = 1,048,576 parameter sets PER GPU (weak scaling), 4 age groups, Spain-2020 window; set b =
clamp(base + sigma * N(0,1)) -- the reference benchmark's own jitter recipe
(sepaihrd_objective_benchmark_main.cpp:452-460), std::mt19937(1); all 1,048,576 sets are distinct draws.
One "step" = one pass of the hot path over the batch = ONE launch of the fused kernel.

value   : whole-job evals/s with the parameter batch resident in HBM; CUDA events on the launching
          stream, barrier + synchronize on both sides, max over ranks.
e2e     : the same metric through the host-buffer C-ABI call (sepaihrd_eval_batch): pinned host
          params H2D + kernel + logL/status D2H inside the timed region, every step.
roofline: FP64 pipe.  achieved = algorithmic FLOP per launch / average launch duration, with the
          algorithmic FLOP of SURVEY.md 8(d): sum_b attempts_b * 3750 + 22,000 * B (unfused count);
          peak = 2 * measured DFMA/s of this GPU (sepaihrd_measure_fp64_peak, run just before the
          timed region -- MEASURED_PEAKS.json has no FP64 entry).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "SEPAIHRD Dopri5+Poisson likelihood evals/sec"
UNIT = "evals/s"
B_PER_GPU = 1 << 20
DISTINCT = 1 << 20              # every set of the batch is its own draw
FLOP_PER_ATTEMPT = 3750.0      # 6 RHS + stage/solution/error combinations + error norm, n = 4
FLOP_PER_SET_FIXED = 22000.0   # 3672 likelihood terms * ~6
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at B = 2^20 (ncu --set full, profiles/r01_v12_ncu_full_summary.txt):
# 527.3 MB + 17.5 MB, i.e. the algorithmic 532.7 MB (62 doubles in, 12 bytes out per set) -- no re-reads.
NCU_DRAM_BYTES_PER_LAUNCH_1M = 527.267328e6 + 17.500160e6
NCU_FP64_PIPE_PCT = 61.52        # sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active, same capture
NCU_ISSUE_ACTIVE_PCT = 51.78     # smsp__issue_active.avg.pct_of_peak_sustained_active


def workload_config(n_gpus: int, batch: int) -> dict:
    return {"workload": "batched likelihood sweep: 1M jittered parameter sets per GPU, 4 age groups, Spain-2020 "
                        "window (BASELINE configs[1])",
            "sets_per_gpu": batch, "global_sets": batch * n_gpus, "params_per_set": 62, "age_groups": 4,
            "output_days": 326, "tolerances": "abs=rel=1e-6", "param_distribution": "clamp(base+sigma*N(0,1)), mt19937(1)",
            "distinct_sets": min(DISTINCT, batch), "math": "fast (FMA)", "parallelism": f"independent shards x{n_gpus}",
            "l2": "inputs (520 MB of parameters per launch) larger than the 126 MB L2; two buffers alternate"}


def make_params(pkg, orc, batch: int):
    prob = pkg.load_default_problem()
    oracle = orc.Oracle(prob)
    distinct = oracle.jitter_params(min(DISTINCT, batch), seed=1)
    reps = (batch + len(distinct) - 1) // len(distinct)
    return prob, oracle, np.tile(distinct, (reps, 1))[:batch]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        rows = []
        with open(self.f.name) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) >= 7:
                    try:
                        rows.append((float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
                    except ValueError:
                        pass
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [r for r in rows if r[2] > 300.0] or rows     # samples under load (power well above idle)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in load for i in range(4) if r[3][i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[0] for r in load])), "sm_max_mhz": rows[0][1],
                "power_w_max": max(r[2] for r in rows), "samples": len(rows), "samples_under_load": len(load),
                "reasons": reasons}


def host_cores() -> int:
    """Host threads the CPU arm may use: every core this process is allowed on.  Passed to the oracle EXPLICITLY because
    torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently shrink the CPU baseline to one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(oracle, params, target_seconds: float = 12.0):
    """The oracle (restated reference path, kind = "port") on all host cores, on a bounded sample."""
    nt = host_cores()
    probe = params[:256]
    t0 = time.perf_counter()
    _, _, _, used = oracle.eval_batch(probe, nthreads=nt)
    rate = len(probe) / (time.perf_counter() - t0)
    m = int(min(len(params), max(1024, rate * target_seconds)))
    t0 = time.perf_counter()
    ll, _, _, used = oracle.eval_batch(params[:m], nthreads=nt)
    dt = time.perf_counter() - t0
    return {"value": m / dt, "unit": UNIT, "cores": int(used), "kind": "port",
            "sample": f"first {m} sets of the same batch, OpenMP schedule(dynamic) over sets, {dt:.1f} s; the oracle is the "
                      "dependency-free restatement of the reference path (the reference needs Boost/Eigen, absent here)"}, ll


def run_reference(args, json_out):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package(); orc = entry.load_oracle()
    orc.build()
    prob, oracle, params = make_params(pkg, orc, B_PER_GPU)
    probe = params[:256]
    nt = host_cores()
    t0 = time.perf_counter(); oracle.eval_batch(probe, nthreads=nt); rate = 256 / (time.perf_counter() - t0)
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    m = int(min(len(params), max(512, rate * min(20.0, budget))))
    for _ in range(args.warmup):
        oracle.eval_batch(params[:m], nthreads=nt)
    t0 = time.perf_counter()
    used = 1
    for s in range(args.steps):
        off = (s * m) % (len(params) - m + 1)
        _, _, _, used = oracle.eval_batch(params[off:off + m], nthreads=nt)
    dt = time.perf_counter() - t0
    value = args.steps * m / dt
    sample = f"{m} sets per step (bounded sample of the 1M-set batch), {used} OpenMP threads"
    json_out.write(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": workload_config(args.gpus, B_PER_GPU),
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(used), "kind": "port", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}) + "\n")
    json_out.flush()


def _claim_stdout():
    """Keep file descriptor 1 for the ONE JSON line: everything else that writes to stdout while the benchmark runs
    (NCCL prints its version banner there, libraries may log) is sent to stderr.  Returns the stream for the JSON line."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="parameter sets per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, json_out)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if rank == 0:
        entry.build_cuda()
    if world > 1:
        dist.barrier()
    pkg = entry.load_package(); orc = entry.load_oracle()
    from sepaihrd_b200.evaluator import BatchEvaluator, measure_fp64_peak

    B = args.batch
    prob, oracle, params = make_params(pkg, orc, B)
    # every rank evaluates its own shard: rank r rotates the batch so shards differ (independent units, no collective)
    params = np.roll(params, shift=rank * 7919, axis=0)
    dev = torch.device("cuda", local_rank)
    h_params = torch.from_numpy(params).pin_memory()
    d_params = [h_params.to(dev, non_blocking=True), torch.from_numpy(params[::-1].copy()).to(dev)]
    d_ll = torch.empty(B, dtype=torch.float64, device=dev)
    d_st = torch.empty(B, dtype=torch.int32, device=dev)
    d_steps = torch.empty((B, 2), dtype=torch.int32, device=dev)
    h_ll = torch.empty(B, dtype=torch.float64).pin_memory()
    h_st = torch.empty(B, dtype=torch.int32).pin_memory()

    ev = BatchEvaluator(prob, device=local_rank)
    stream = torch.cuda.current_stream(dev)
    ev.set_stream(stream.cuda_stream)
    P = prob.n_params

    def step(i: int, with_steps: bool = False):
        ev.eval_into(d_params[i & 1].data_ptr(), B, P, d_ll.data_ptr(), d_st.data_ptr(), d_steps.data_ptr() if with_steps else 0)

    peak_dfma = measure_fp64_peak(local_rank)
    # attempts per set (for the algorithmic FLOP count), measured once outside the timed region
    step(0, True); step(1, True); torch.cuda.synchronize()
    attempts_total = float(d_steps.sum().item())           # buffer 1 is buffer 0 reversed: same total
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0, _ = ev.counters()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    sync_all()
    ms_kernel = e0.elapsed_time(e1)
    launches1, _ = ev.counters()

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------------
    for _ in range(2):
        ev.eval_host_into(h_params.data_ptr(), B, P, h_ll.data_ptr(), h_st.data_ptr())
    f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
    sync_all()
    t_wall0 = time.perf_counter()
    f0.record(stream)
    for i in range(args.steps):
        ev.eval_host_into(h_params.data_ptr(), B, P, h_ll.data_ptr(), h_st.data_ptr())
    f1.record(stream)
    sync_all()
    ms_e2e = max(f0.elapsed_time(f1), (time.perf_counter() - t_wall0) * 1e3)
    launches2, _ = ev.counters()
    clocks = sampler.stop() if sampler else None
    checksum = float(h_ll.sum().item())

    if world > 1:
        t = torch.tensor([ms_kernel, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_kernel, ms_e2e = t[0].item(), t[1].item()

    if rank == 0:
        value = world * args.steps * B / (ms_kernel * 1e-3)
        e2e_value = world * args.steps * B / (ms_e2e * 1e-3)
        flop_per_launch = attempts_total * FLOP_PER_ATTEMPT + FLOP_PER_SET_FIXED * B
        launch_s = ms_kernel * 1e-3 / args.steps
        achieved = flop_per_launch / launch_s / 1e12
        peak = 2.0 * peak_dfma / 1e12
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_kernel / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload_config(world, B),
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * P * 8, "d2h_bytes_per_step": B * 12,
                       "ms_per_step": ms_e2e / args.steps},
               "gpu_launches": int(launches1 - launches0),
               "gpu_launches_e2e": int(launches2 - launches1),
               "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            "traffic": NCU_DRAM_BYTES_PER_LAUNCH_1M if B == (1 << 20) else None,
                            "kernel": "sepaihrd_batch_kernel<4,fast,LL>",
                            "flop_per_launch": flop_per_launch, "attempts_per_set": attempts_total / B,
                            "launch_ms": launch_s * 1e3,
                            "peak_source": "measured live: sepaihrd_measure_fp64_peak (dependent DFMA chains), x2 FLOP per DFMA; "
                                           "MEASURED_PEAKS.json has no FP64 entry",
                            "hbm_bytes_per_launch_algorithmic": B * (P * 8 + 12),
                            # BASELINE.json's "% FP64 peak (ncu fp64-pipe utilisation and issue efficiency)": from the ncu --set full
                            # capture of this kernel at this batch size (profiles/r01_v12_ncu_full_summary.txt), not measured live
                            "ncu_fp64_pipe_pct": NCU_FP64_PIPE_PCT if B == (1 << 20) else None,
                            "ncu_issue_active_pct": NCU_ISSUE_ACTIVE_PCT if B == (1 << 20) else None},
               "clocks": clocks, "logl_checksum": checksum}
        if world == 1:
            # SURVEY.md 8(d) names a second distribution: uniform in bounds (the PSO initialisation recipe), mt19937(2).  Not the
            # headline (warps idle more when their 8 sets need different attempt counts); reported beside it, outside the
            # timed region above.
            try:
                uni = torch.from_numpy(oracle.uniform_params(B, seed=2)).to(dev)
                step_u = lambda: ev.eval_into(uni.data_ptr(), B, P, d_ll.data_ptr(), d_st.data_ptr(), d_steps.data_ptr())
                step_u(); torch.cuda.synchronize()
                att_u = float(d_steps.sum().item())
                u0 = torch.cuda.Event(enable_timing=True); u1 = torch.cuda.Event(enable_timing=True)
                u0.record(stream)
                for _ in range(3):
                    step_u()
                u1.record(stream); torch.cuda.synchronize()
                ms_u = u0.elapsed_time(u1) / 3
                out["second_distribution"] = {"param_distribution": "uniform in bounds, mt19937(2)", "value": B / (ms_u * 1e-3), "unit": UNIT,
                                              "ms_per_step": ms_u, "attempts_per_set": att_u / B,
                                              "roofline_frac": (att_u * FLOP_PER_ATTEMPT + FLOP_PER_SET_FIXED * B) / (ms_u * 1e-3) / 1e12 / peak}
                del uni
            except Exception as exc:           # informational only: never lose the headline line over it
                out["second_distribution"] = {"error": str(exc)}
        if not args.no_cpu_baseline and world == 1:
            base, ll_cpu = cpu_baseline(oracle, params)
            m = len(ll_cpu)
            rel = np.abs(h_ll.numpy()[:m] - ll_cpu) / np.abs(ll_cpu)
            base["max_rel_logl_diff_vs_gpu"] = float(rel.max())
            out["cpu_baseline"] = base
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    ev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
