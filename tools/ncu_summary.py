#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the handful of numbers DESIGN.md / profiles/ cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index] > profiles/rNN_<name>.txt
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --json=profiles/current_ncu_capture.json --sets=1048576 --summary=profiles/rNN_<name>.txt
        also writes the capture record bench.py reports under roofline.ncu_capture (with the hash of the kernel sources)
"""
import csv
import subprocess
import sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
opts = dict(a[2:].split("=", 1) for a in sys.argv[1:] if a.startswith("--") and "=" in a)
rep = args[0]
which = int(args[1]) if len(args) > 1 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + which]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def show(key, label=None):
    if key in m:
        v, u = m[key]
        print(f"{label or key:70s} {v} {u}")


print("kernel:", m.get("Kernel Name", ("?",))[0][:120])
for k in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
          "sm__cycles_elapsed.max", "sm__cycles_active.avg",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
          "smsp__issue_inst0.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
          "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads",
          "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
          "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_ops_dadd_dmul_dfma_pred_on.sum",
          "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shuffle.sum" ]:
    show(k)
print("\n-- warp issue-stall reasons (smsp__average_warp*_issue_stalled_*_per_issue_active / pcsamp) --")
st = [(h, v) for h, (v, u) in m.items() if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
if not st:
    st = [(h, v) for h, (v, u) in m.items() if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
for h, v in sorted(st, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:14]:
    print(f"{h:90s} {v}")
print("\n-- pc sampling totals --")
ps = [(h, v) for h, (v, u) in m.items() if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
for h, v in sorted(ps, key=lambda kv: -float(kv[1].replace(",", "") or 0))[:14]:
    print(f"{h:90s} {v}")

if "json" in opts:
    import hashlib, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "mathematical-modeling-of-infectious-diseases-v1_b200", "csrc")
    h = hashlib.sha256()
    for f in ("sepaihrd_kernels.cuh", "sepaihrd_constraints.cuh"):
        h.update(open(os.path.join(csrc, f), "rb").read())
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}

    def num(key):
        v, u = m[key]
        return float(v.replace(",", "")) * scale.get(u, 1.0)
    rec = {"kernel": m.get("Kernel Name", ("?",))[0][:160], "summary_file": opts.get("summary"), "sets_per_launch": int(opts.get("sets", 0)),
           "kernel_source_hash": h.hexdigest()[:16],
           "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
           "fp64_pipe_pct": num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "duration_ms_under_ncu": num("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(m["gpu__time_duration.sum"][1], 1.0),
           "registers_per_thread": num("launch__registers_per_thread")}
    with open(opts["json"], "w") as f:
        json.dump(rec, f, indent=1)
