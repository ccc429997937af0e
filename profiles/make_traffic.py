#!/usr/bin/env python
"""DRAM bytes per launch of every kernel of a step from an `ncu --set full` report -> profiles/r2_traffic.json,
stamped with the sha256 of the kernel sources (bench.py reports `roofline.traffic` from it only while that sha
matches).  usage: python profiles/make_traffic.py REPORT.ncu-rep config lanes contexts"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (source_sha)

rep, config, lanes, contexts = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


seen = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0]
    rd = to_bytes(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
    wr = to_bytes(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
    seen.setdefault(name, []).append(rd + wr)
out = {"source": f"profiles/{os.path.basename(rep).replace('.ncu-rep', '')}_kernels.txt (ncu --set full --clock-control none; launches of one "
                 f"step of bench.py at {lanes} lanes in {contexts} contexts, {config}, range-image input); "
                 "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the captured launches of the kernel",
       "config": config, "lanes": lanes, "contexts": contexts, "input": "range", "source_sha": bench.source_sha(),
       "kernels": {k: int(sum(v) / len(v)) for k, v in seen.items()}, "launches_captured": {k: len(v) for k, v in seen.items()}}
with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
