#!/usr/bin/env python
"""Time the hash-sharded single-sequence mode under torchrun:
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/run_sharded.py [scans]
Prints one JSON line on rank 0: scans/s of ONE sequence on N GPUs, collectives per scan, and whether the
poses equal the single-GPU step's bit for bit (checked on every rank)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptudes_lab_b200 import odometry, sharded, synth  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
seq = synth.make_sequence("os0_quad", 0)
gen = synth.TorchScanGenerator(seq, torch.device("cuda", local))
ranges = [gen.range_image(k)[0].contiguous() for k in range(n_scans)]
cfg = odometry.load_config(None, deskew=True, max_range=100.0)
o = odometry.Odometry(cfg, device=local, max_points=131072, map_capacity=65536)
single = odometry.Odometry(cfg, device=local, max_points=131072, map_capacity=65536)
for x in (o, single):
    x.set_sensor(seq.dirs)
so = sharded.ShardedOdometry(sharded.PtkShardBackend(o, rank, world))
same, iters = True, 0
t_sh = t_1 = 0.0
for k in range(n_scans):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p1, s1 = single.register_scan(ranges[k])
    t1 = time.perf_counter()
    pose, st = so.register_frame(None, None, range_mm=ranges[k])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    same = same and np.array_equal(pose, p1)
    if k >= 3:
        t_1 += t1 - t0
        t_sh += t2 - t1
        iters += st["iterations"]
flag = torch.tensor([1 if same else 0], device="cuda")
tt = torch.tensor([t_sh], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    n = n_scans - 3
    print(json.dumps({"mode": "hash-sharded map, one sequence", "n_gpus": world, "scans": n,
                      "sharded_scans_per_s": n / float(tt.item()), "single_gpu_scans_per_s": n / t_1,
                      "mean_icp_iterations": iters / n, "collectives_per_scan": so.collectives / n_scans,
                      "poses_bit_identical_to_single_gpu": bool(flag.item())}))
if world > 1:
    dist.destroy_process_group()
