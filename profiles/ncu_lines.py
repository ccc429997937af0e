#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics + the hottest CUDA source lines.
usage: python profiles/ncu_lines.py REPORT.ncu-rep [kernel-index] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for i, r in enumerate(rows[2:]):
    d = dict(zip(hdr, r))
    print(f"== launch {i}")
    for k in WANT:
        if k in d:
            print(f"  {k} = {d[k]} {rows[1][hdr.index(k)]}")
    st = sorted(((float(d[k]), k) for k in STALL if d[k] not in ("", "n/a")), reverse=True)[:6]
    print("  top stalls (warps per issue):", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, k in st))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = {}
cur_file, k, hdr2 = None, -1, None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        k += 1
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr2 = r
        continue
    if hdr2 is None or len(r) != len(hdr2) or r[0] == "":
        continue
    if k not in (kidx, -1):
        continue
    d = dict(zip(hdr2, r))
    key = (cur_file, int(r[0]))
    e = lines.setdefault(key, {"src": r[1].strip(), "samples": 0, "inst": 0, "bar": 0, "lsb": 0, "wait": 0, "ssb": 0, "mio": 0})
    e["samples"] += int(d["# Samples"] or 0)
    e["inst"] += int(d["Instructions Executed"] or 0)
    for a, b in (("bar", "stall_barrier"), ("lsb", "stall_long_sb"), ("wait", "stall_wait"), ("ssb", "stall_short_sb"), ("mio", "stall_mio")):
        e[a] += int(d.get(b, 0) or 0)
tot = sum(e["samples"] for e in lines.values()) or 1
toti = sum(e["inst"] for e in lines.values()) or 1
print(f"== source lines of kernel {kidx}: {tot} samples, {toti} warp instructions")
for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    print(f"  {100 * e['samples'] / tot:5.1f}% smp {100 * e['inst'] / toti:5.1f}% inst  bar {e['bar']:5d} lsb {e['lsb']:5d} "
          f"wait {e['wait']:5d} ssb {e['ssb']:4d}  {f}:{ln}  {e['src'][:90]}")
