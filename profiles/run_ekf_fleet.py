#!/usr/bin/env python
"""BASELINE config 4 at fleet scale: B synthetic lidar+IMU sequences (OS0-128 1024x10 + 100 Hz IMU, ekf-bench
ranges min 1 / max 70), per step and lane: 10 IMU samples into the host-native ESEKF (ptk_ekf_*), the filter's
pose as the ICP initial guess (--use-imu-prediction, cli/ekf_bench.py:533-535), one batched GPU odometry step
from pinned host range images, the pose into the filter (processPose, :554-557).
usage: python profiles/run_ekf_fleet.py [lanes] [scans] [python|native]"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ptudes_lab_b200 import _ffi, odometry, synth
from ptudes_lab_b200.ekf_bench import SynthLidarImuSource
from ptudes_lab_b200.ins import ESEKF, ESEKFNative, IMU, calc_ate

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kind = sys.argv[3] if len(sys.argv) > 3 else "native"
dev = torch.device("cuda", 0)
seqs = [synth.make_sequence("os0_quad", l) for l in range(B)]
gens = [synth.TorchScanGenerator(s, dev) for s in seqs]
scans, gts = [], []
for s in range(T):
    row = []
    for g in gens:
        h = _ffi.pinned_empty((128, 1024), dtype=np.uint32)
        rng, _, gt = g.range_image(s)
        h[...] = rng.cpu().numpy().astype(np.uint32)
        row.append(h)
        if g is gens[0]:
            gts.append(gt)
    scans.append(row)
srcs = [SynthLidarImuSource(s, T, seed=1 + l) for l, s in enumerate(seqs)]
imus = [[[src.imu_at(k * 0.1 + j * 0.01) for j in range(10)] for k in range(T)] for src in srcs]     # [lane][scan][10]
la = np.array([[[i.lacc for i in sc] for sc in lane] for lane in imus]); av = np.array([[[i.avel for i in sc] for sc in lane] for lane in imus])
ts = np.array([[[i.ts for i in sc] for sc in lane] for lane in imus])
cfg = odometry.load_config(None, deskew=True, max_range=70.0); cfg.data.min_range = 1.0
o = odometry.Odometry(cfg, device=0, max_points=131072, map_capacity=65536, batch=B)
o.set_sensor(seqs[0].dirs)
ekfs = [ESEKFNative() if kind == "native" else ESEKF() for _ in range(B)]
poses0 = []
t_ekf = t_gpu = 0.0
torch.cuda.synchronize()
t_start = None
for s in range(T):
    if s == 3:
        torch.cuda.synchronize(); t_start = time.perf_counter(); t_ekf = t_gpu = 0.0
    t0 = time.perf_counter()
    guesses = []
    for l, f in enumerate(ekfs):
        if kind == "native":
            f.processImuBatch(la[l, s], av[l, s], ts[l, s])
            guesses.append(f.pose_mat())
        else:
            for j in range(10):
                f.processImu(IMU(la[l, s, j].copy(), av[l, s, j].copy(), float(ts[l, s, j])))
            guesses.append(f.nav.pose_mat())
    t1 = time.perf_counter()
    if s + 1 < T:
        o.prefetch_scan_batch(scans[s + 1])
    poses, _ = o.register_scan_batch(scans[s], guesses=guesses)
    t2 = time.perf_counter()
    for l, f in enumerate(ekfs):
        f.processPose(poses[l])
    t3 = time.perf_counter()
    t_ekf += (t1 - t0) + (t3 - t2); t_gpu += t2 - t1
    poses0.append(poses[0].copy())
dt = time.perf_counter() - t_start
n = T - 3
g0 = np.linalg.inv(gts[0])
ate = calc_ate(poses0, [g0 @ g for g in gts])
print(json.dumps({"config": "ekf-bench fleet, OS0-128 1024x10 + 100 Hz IMU, ranges 1/70, imu prediction as ICP guess",
                  "filter": kind, "lanes": B, "scans_per_s": B * n / dt, "ms_per_step": 1e3 * dt / n,
                  "host_filter_ms_per_step": 1e3 * t_ekf / n, "gpu_step_ms": 1e3 * t_gpu / n,
                  "ate_lane0_vs_gt": {"rot": ate[0], "trans_m2": ate[1]}}))
o.close()
