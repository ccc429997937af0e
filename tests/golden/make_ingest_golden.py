#!/usr/bin/env python
"""Golden vectors of the ingest step: for each of the four lidar UDP profiles one small frame (H = 4, W = 32, 16 columns
per packet) as packet bytes, with an invalid column, plus the LidarScan fields the oracle's ScanBatcher rules give for it.
Written by oracle/ingest_oracle.py (the SDK is absent: **parity unpinned**, see its header); the tests hold the oracle, the
host-side batcher and the device decode to these bytes.  usage: python tests/golden/make_ingest_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ingest_oracle as io  # noqa: E402

out = {}
for prof in (io.LEGACY, io.DUAL, io.RNG19, io.RNG15):
    F = io.Format(prof, 4, 16, 32)
    rng = np.random.default_rng(100 + prof)
    rmax = {io.LEGACY: 1 << 20, io.RNG19: 1 << 19, io.DUAL: 1 << 19, io.RNG15: 1 << 18}[prof]
    f = {"RANGE": rng.integers(0, rmax, size=(F.H, F.W), dtype=np.uint32),
         "RANGE2": rng.integers(0, 1 << 19, size=(F.H, F.W), dtype=np.uint32),
         "REFLECTIVITY": rng.integers(0, 256, size=(F.H, F.W), dtype=np.uint32),
         "SIGNAL": rng.integers(0, 65536, size=(F.H, F.W), dtype=np.uint32),
         "NEAR_IR": rng.integers(0, 4096, size=(F.H, F.W), dtype=np.uint32) << 4}
    if prof == io.RNG15:
        f["RANGE"] &= ~np.uint32(7)
    valid = np.ones(F.W, bool)
    valid[19] = False
    ts = (5_000_000_000 + np.arange(F.W) * 48_828).astype(np.uint64)
    pk = io.encode_frame(F, 65535, f, ts, valid=valid)
    out[f"p{prof}_packets"] = np.frombuffer(b"".join(pk), dtype=np.uint8).reshape(F.ppf, F.size)
    for k, v in io.decode_frame(F, pk).items():
        out[f"p{prof}_{k}"] = v
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest_tiny.npz"), **out)
print("wrote ingest_tiny.npz:", sorted(out)[:6], "...")
