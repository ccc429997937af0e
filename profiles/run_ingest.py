#!/usr/bin/env python
"""The ingest side run of bench.py on its own (k_decode_packets: frames per second, GB/s against the HBM peak),
optionally for several grid shapes: python profiles/run_ingest.py [blocks_per_sm ...]  (0 = one block per packet)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

for v in (sys.argv[1:] or [""]):
    if v == "":
        os.environ.pop("PTK_DECODE_BLOCKS_PER_SM", None)
    else:
        os.environ["PTK_DECODE_BLOCKS_PER_SM"] = v
    r = bench.side_ingest(0)
    print(json.dumps({"blocks_per_sm": v or "default", "scans_per_s": r["value"], "ms_per_launch": r["ms_per_launch"],
                      "GBps": r["roofline"]["achieved"], "frac": r["roofline"]["frac"], "e2e": r["e2e"]["value"]}), flush=True)
