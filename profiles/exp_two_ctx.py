#!/usr/bin/env python
"""Experiment: does driving TWO contexts (half the lanes each) from two host threads on two streams beat
one context with all lanes?  (tail of one half's ICP overlapping the other half's streaming kernels)
usage: python profiles/exp_two_ctx.py [lanes_total] [steps]"""
import os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from ptudes_lab_b200 import odometry, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
W = 5; T = W + K
dev = torch.device("cuda", 0)
gens = [synth.TorchScanGenerator(synth.make_sequence("os0_quad", l), dev) for l in range(B)]
ranges = [[g.range_image(s)[0].contiguous() for g in gens] for s in range(T)]
cfg = odometry.load_config(None, deskew=True, max_range=100.0)

def run(parts):
    ctxs = []
    for lanes in parts:
        o = odometry.Odometry(cfg, device=0, max_points=131072, map_capacity=32768, batch=len(lanes))
        o.set_sensor(gens[0].seq.dirs)
        ctxs.append((o, lanes, torch.cuda.Stream(device=dev)))
    def work(o, lanes, st, lo, hi):
        for s in range(lo, hi):
            o.register_scan_batch([ranges[s][l] for l in lanes], stream=st.cuda_stream)
    for o, lanes, st in ctxs:
        work(o, lanes, st, 0, W)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(o, lanes, st, W, T)) for o, lanes, st in ctxs]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for o, _, _ in ctxs: o.close()
    return B * K / dt
print("one context :", round(run([list(range(B))])))
print("two contexts:", round(run([list(range(B // 2)), list(range(B // 2, B))])))
print("three       :", round(run([list(range(0, B, 3)), list(range(1, B, 3)), list(range(2, B, 3))])))
