#!/usr/bin/env python
"""Per-region view of an ncu --set full --import-source on capture (SASS page).

Splits the kernel's SASS into runs of equal execution count (loop nests) and prints, per run, its share
of the issued instructions and of the warp-state samples, the opcode mix and the dominant stall reasons.

    python tools/ncu_source_regions.py gpurun_out/prof.ncu-rep [--min-share 0.5] [--dump A B]
"""
import argparse, csv, subprocess, sys
from collections import Counter

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--min-share", type=float, default=0.4, help="hide runs below this %% of samples and of instructions")
ap.add_argument("--dump", type=int, nargs=2, default=None, help="dump SASS lines A..B with samples and stalls")
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[h0], rows[h0 + 1:]
ix = {h: i for i, h in enumerate(hdr)}
STALLS = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
smp = [int(r[ix["# Samples"]]) for r in data]
exe = [int(r[ix["Instructions Executed"]]) for r in data]
T, E = sum(smp), sum(exe)


def opcode(r):
    return [o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]


if a.dump:
    for i in range(a.dump[0], a.dump[1] + 1):
        r = data[i]
        tags = [f"{k[6:]}={int(r[ix[k]])}" for k in STALLS if int(r[ix[k]]) > max(100, 0.2 * smp[i])]
        print(f"{i:5d} {exe[i]:10d} {smp[i]:7d}  {r[ix['Source']].strip()[:72]:72s} {' '.join(tags)}")
    sys.exit(0)

print(f"kernel: {rows[0][1][:110] if rows[0] else '?'}")
print(f"SASS lines {len(data)}, warp instructions executed {E:.4e}, samples {T}")
segs, start = [], 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(exe[i] - exe[start]) > 0.02 * max(exe[start], 1):
        segs.append((start, i - 1)); start = i
print(f"{'lines':>11s} {'exec/line':>10s} {'inst%':>6s} {'smp%':>6s} {'cyc/inst':>8s}  opcode mix / stalls")
for s, e in segs:
    n_i = sum(exe[s:e + 1]); n_s = sum(smp[s:e + 1])
    if 100.0 * n_i / E < a.min_share and 100.0 * n_s / T < a.min_share:
        continue
    ops = Counter(opcode(r) for r in data[s:e + 1])
    st = Counter()
    for r in data[s:e + 1]:
        for k in STALLS:
            st[k[6:]] += int(r[ix[k]])
    sel = max(st.get("selected", 0), 1)
    mix = " ".join(f"{k}:{v}" for k, v in ops.most_common(6))
    stl = " ".join(f"{k}={100.0 * v / max(n_s, 1):.0f}%" for k, v in st.most_common(4))
    print(f"{s:5d}-{e:5d} {exe[s]:10d} {100.0 * n_i / E:6.1f} {100.0 * n_s / T:6.1f} {n_s / sel:8.2f}  {mix} | {stl}")
