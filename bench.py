#!/usr/bin/env python
"""bench.py - ICP odometry scans/sec on synthetic OS0-128 1024x10 sequences (BASELINE.json).

A "step" advances every resident sequence ("lane") of this rank by one scan through the full
odometry step (deskew + range filter + 2x voxel downsample + ICP registration + local-map
insert + prune): one call of ptk_register_frame_batch.  Workload = BASELINE.json configs[1]
(100-scan OS0-128 1024x10 sequence on the quad scene), `--lanes` independent copies with
different seeds per GPU (configs[4]'s fleet replay: no data-path collective, weak scaling).

  python bench.py --gpus N --steps K --warmup W           own arm (CUDA path through the C ABI)
  python bench.py --impl reference ...                   reference arm: the CPU restatement of
                                                         the kiss-icp path (oracle/) on host cores

One JSON line on stdout (rank 0).  Keys: see the task contract; `value` = scans/s with inputs
resident in HBM, `e2e` = the same through the C ABI with HOST (pinned) buffers, H2D inside
the timed region and the poses read back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RESULT = sys.stdout
METRIC = "icp_odometry_scans_per_sec"
UNIT = "scans/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ptk", choices=["ptk", "reference"])
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("PTK_BENCH_LANES", "64")),
                    help="independent sequences per GPU")
    ap.add_argument("--contexts", type=int, default=int(os.environ.get("PTK_BENCH_CONTEXTS", "8")),
                    help="contexts the lanes of a GPU are dealt to (each advances its lanes in lock step, on its own thread)")
    ap.add_argument("--icp-blocks", type=int, default=6, help="ICP blocks per lane when several contexts share the GPU")
    ap.add_argument("--config", default="os0_quad", choices=["os0_quad", "os2_street", "os0_hall"])
    ap.add_argument("--input", default="range", choices=["range"],
                    help="range: RANGE image as KissICPWrapper.register_frame gets it")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side-runs", action="store_true", help="skip the small runs of configs[2..4] (side_runs)")
    ap.add_argument("--no-prefetch", action="store_true", help="e2e leg: do not overlap the next scans' H2D copy")
    ap.add_argument("--profile-pass", action="store_true", default=True)
    return ap.parse_args()


CONFIGS = {
    # name: (min_range, max_range, max_points, map_capacity)
    "os0_quad": (5.0, 100.0, 131072, 32768),
    "os0_hall": (5.0, 100.0, 131072, 131072),
    "os2_street": (5.0, 200.0, 262144, 65536),
}


# --------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled DURING the timed region: NVML polled from a thread
    every few milliseconds (the timed region of a 40-step run is ~0.1 s, too short for `nvidia-smi -lms`)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu, self.period = gpu_index, period_s
        self.samples = []            # (sm_mhz, reasons bitmask, power_w)
        self.h = None
        self._stop = threading.Event()
        self.t = None
        self.sm_max = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.h = None
            self.t = threading.Thread(target=self._poll_smi, daemon=True)     # fallback: nvidia-smi one-shots
            self.t.start()

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.hw_thermal_slowdown")
        names = ["hw_slowdown", "sw_power_cap", "sw_thermal_slowdown", "hw_thermal_slowdown"]
        while not self._stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                bits = 0
                for n_, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        bits |= self.REASONS[n_]
                self.sm_max = float(r[1])
                self.samples.append((float(r[0]), bits, float(r[2])))
            except Exception:
                self._stop.wait(0.05)

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.samples.append((sm, rs, pw))
            except Exception:
                pass
            self._stop.wait(self.period)

    def mark(self):
        return len(self.samples)

    def stop(self):
        self._stop.set()
        if self.t is not None:
            self.t.join(timeout=1)

    def summary(self, lo=0, hi=None):
        window = "timed region"
        rows = self.samples[lo:hi]
        if not rows:
            rows, window = self.samples, "whole run (no sample fell into the timed region)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        bits = 0
        for _, r, _ in rows:
            bits |= r
        pw = [p for _, _, p in rows if p is not None]
        return {"sm_mhz": float(np.median([r[0] for r in rows])), "sm_max_mhz": self.sm_max,
                "reasons": sorted(k for k, m in self.REASONS.items() if bits & m), "samples": len(rows),
                "power_w_max": max(pw) if pw else None, "window": window}


def source_sha():
    """sha256 (16 hex digits) of the kernel sources with comments and white space removed: ties a committed ncu
    capture to the CODE that produced it (editing a comment does not orphan a capture, editing a kernel does)."""
    import hashlib
    import re
    h = hashlib.sha256()
    for f in ("ptk_device.cuh", "ptk_canon.cuh"):
        with open(os.path.join(ROOT, "ptudes_lab_b200", "csrc", f), "r") as fh:
            src = fh.read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        h.update(re.sub(r"\s+", "", src).encode())
    return h.hexdigest()[:16]


def recorded_traffic(kernel, config, lanes, inp, contexts=1):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/r2_traffic.json),
    but only if it was taken on this very configuration AND on these very kernel sources (source_sha);
    None otherwise - never a guess, never a stale number."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            t = json.load(f)
        if (t["config"] == config and t["lanes"] == lanes and t["input"] == inp and t.get("contexts", 1) == contexts
                and t.get("source_sha") == source_sha()):
            return t["kernels"].get(kernel)
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# algorithmic (compulsory) HBM bytes of every kernel of the step, per lane-scan, from the
# counters the library returns (DESIGN.md "Algorithmic bytes").  b = 8 B per
# coordinate, S = 16 B map slot.
def algorithmic_bytes(st, n_pixels=0):
    b, S = 8, 16
    N, Nd, Ns, V, M = st["n_in"], st["n_ds"], st["n_src"], st["n_voxels"], st["map_points"]
    Vt = min(27 * Ns, V)
    Mt = min(20 * Vt, M)
    return {
        # range-image input: 4 B per pixel; the direction LUT is a per-sensor constant (3 MB) shared by
        # every lane and scan, not per-scan traffic
        "k_scan_insert": n_pixels * 4 if n_pixels else N * (3 * b + 8),
        "k_compact1": Nd * 3 * b,
        "k_compact2": Nd * 3 * b + Ns * 3 * b,
        "k_icp": Ns * 3 * b + Vt * S + Mt * 3 * b,
        "k_map_insert": Nd * 3 * b + Nd * S,
        "k_map_commit": Nd * 3 * b,
        "k_map_prune": V * (S + 3 * b),
        "k_finish": 344,
    }


# --------------------------------------------------------------------------------------
# Side runs carried in the same JSON line: the other BASELINE.json configs, so that every round's driver-run
# BENCH/SCALE record holds evidence for them (they are parity-test cases first; tests/test_gpu_parity.py).
def side_fleet(config, lanes, warmup, steps, local, seed0, sync, contexts=1, icp_blocks=6):
    """`lanes` sequences of `config` advanced `warmup + steps` scans, dealt to `contexts` free-running contexts
    (ptk_fleet_replay); device-resident range images, CUDA events around the timed scans.  Returns (ms of the timed
    scans, last-scan counters of lane 0)."""
    import torch
    from ptudes_lab_b200 import odometry, synth
    dev = torch.device("cuda", local)
    min_r, max_r, max_pts, map_cap = CONFIGS[config]
    T = warmup + steps
    G = max(1, min(contexts, lanes))
    gens = [synth.TorchScanGenerator(synth.make_sequence(config, seed0 + l), dev) for l in range(lanes)]
    ranges = [[g.range_image(s)[0].contiguous() for g in gens] for s in range(T)]
    cfg = odometry.load_config(None, deskew=True, max_range=max_r)
    cfg.data.min_range = min_r
    parts = [list(range(g * lanes // G, (g + 1) * lanes // G)) for g in range(G)]
    odos = [odometry.Odometry(cfg, device=local, max_points=max_pts, map_capacity=map_cap, batch=len(p)) for p in parts]
    streams = [torch.cuda.Stream(device=dev) for _ in odos]
    sh = [st.cuda_stream for st in streams]
    try:
        for o in odos:
            o.set_sensor(gens[0].seq.dirs)
            if G > 1:
                o.set_icp_blocks_per_lane(icp_blocks)

        def replay(lo, hi, want_stats=False):
            rg = [[[ranges[s][l] for l in parts[g]] for s in range(lo, hi)] for g in range(G)]
            return odometry.fleet_replay(odos, rg, sh, want_stats=want_stats)
        replay(0, warmup)
        sync()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in odos]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in odos]
        for ev, st in zip(e0, streams):
            ev.record(st)
        _, stats = replay(warmup, T, want_stats=True)
        for ev, st in zip(e1, streams):
            ev.record(st)
        sync()
        last = stats[0][steps - 1][0]
        return (max(e0[0].elapsed_time(ev) for ev in e1),
                {k: last[k] for k in ("n_in", "n_ds", "n_src", "n_voxels", "map_points", "iterations")})
    finally:
        for o in odos:
            o.close()


def side_ekf_fleet(lanes, warmup, steps, local):
    """configs[3]: the ekf-bench loop (cli/ekf_bench.py:493-563) at fleet scale - per step and lane 10 IMU samples
    into the host-native ESEKF, the filter's pose as the ICP initial guess (--use-imu-prediction, :533-535), one
    batched odometry step from pinned HOST range images, the pose into the filter (:554-557).  Host wall clock."""
    import torch
    from ptudes_lab_b200 import _ffi, odometry, synth
    from ptudes_lab_b200.ekf_bench import SynthLidarImuSource
    from ptudes_lab_b200.ins import ESEKFNative
    dev = torch.device("cuda", local)
    T = warmup + steps
    seqs = [synth.make_sequence("os0_quad", l) for l in range(lanes)]
    gens = [synth.TorchScanGenerator(s, dev) for s in seqs]
    scans = []
    for s in range(T):
        row = []
        for g in gens:
            h = _ffi.pinned_empty((seqs[0].sensor.H, seqs[0].sensor.W), dtype=np.uint32)
            h[...] = g.range_image(s)[0].cpu().numpy().astype(np.uint32)
            row.append(h)
        scans.append(row)
    srcs = [SynthLidarImuSource(s, T, seed=1 + l) for l, s in enumerate(seqs)]
    imus = [[[src.imu_at(k * 0.1 + j * 0.01) for j in range(10)] for k in range(T)] for src in srcs]
    la = np.array([[[i.lacc for i in sc] for sc in lane] for lane in imus])
    av = np.array([[[i.avel for i in sc] for sc in lane] for lane in imus])
    ts = np.array([[[i.ts for i in sc] for sc in lane] for lane in imus])
    cfg = odometry.load_config(None, deskew=True, max_range=70.0)          # ekf-bench defaults (cli/ekf_bench.py:356-363)
    cfg.data.min_range = 1.0
    o = odometry.Odometry(cfg, device=local, max_points=131072, map_capacity=65536, batch=lanes)
    o.set_sensor(seqs[0].dirs)
    ekfs = [ESEKFNative() for _ in range(lanes)]
    t_start, t_ekf = None, 0.0
    try:
        for s in range(T):
            if s == warmup:
                torch.cuda.synchronize()
                t_start, t_ekf = time.perf_counter(), 0.0
            t0 = time.perf_counter()
            guesses = []
            for l, f in enumerate(ekfs):
                f.processImuBatch(la[l, s], av[l, s], ts[l, s])
                guesses.append(f.pose_mat())
            t1 = time.perf_counter()
            if s + 1 < T:
                o.prefetch_scan_batch(scans[s + 1])
            poses, st = o.register_scan_batch(scans[s], guesses=guesses)
            t2 = time.perf_counter()
            for l, f in enumerate(ekfs):
                f.processPose(poses[l])
            t_ekf += (t1 - t0) + (time.perf_counter() - t2)
        dt = time.perf_counter() - t_start
    finally:
        o.close()
    return {"value": lanes * steps / dt, "unit": UNIT, "lanes": lanes, "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "host_filter_ms_per_step": 1e3 * t_ekf / steps, "timing": "host wall clock, pinned host range images",
            "workload": "configs[3]: ekf-bench loop, OS0-128 1024x10 + 100 Hz IMU, ranges 1/70 (voxel 0.7), ESEKF pose as "
                        "ICP guess, host-native ESEKF"}


def side_ingest(local, frames=384, reps=20):
    """SURVEY 8f-4, the step before the path: LEGACY lidar packets of `frames` OS0-128 1024x10 scans -> RANGE images.
    (a) packets resident in HBM, ONE launch of k_decode_packets per repetition (384 scans = 0.6 GB of packets per launch, so
    that the launch, not the host's enqueue rate, is what the CUDA events see); algorithmic bytes = every packet byte
    read once + every pixel of the image and every column header written once.  (b) the same through the
    ScanBatcher object with HOST packets pushed one by one (pinned ring -> H2D -> decode), host wall clock."""
    import ctypes as C
    import torch
    from ptudes_lab_b200 import _ffi, ingest, synth
    dev = torch.device("cuda", local)
    seq = synth.make_sequence("os0_quad", 0)
    H, W = seq.sensor.H, seq.sensor.W
    pf = ingest.PacketFormat(ingest.PROFILE_LIDAR_LEGACY, H, 16, W)
    gen = synth.TorchScanGenerator(seq, dev)
    distinct = []
    for k in range(8):
        r = gen.range_image(k)[0].cpu().numpy().astype(np.uint32)
        ts = (np.arange(W) * (100_000_000 // W) + k * 100_000_000).astype(np.uint64)
        distinct.append((r, ingest.encode_scan_packets(pf, k, r, ts)))
    host = np.concatenate([distinct[f % 8][1] for f in range(frames)])            # (frames * ppf, packet size)
    d_packets = torch.from_numpy(host).to(dev)
    out = ingest.decode_frames(pf, d_packets, frames, device=local)                # warm-up + check
    torch.cuda.synchronize()
    for f in (0, frames - 1):
        assert np.array_equal(out["RANGE"][f].cpu().numpy().view(np.uint32), distinct[f % 8][0]), "ingest: decoded image differs"
    for _ in range(3):
        ingest.decode_frames(pf, d_packets, frames, device=local, out=out)
    torch.cuda.synchronize()
    blocks = []
    for _ in range(5):                      # like the peak it is compared with (MEASURED_PEAKS.json: best of 10): best block
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ingest.decode_frames(pf, d_packets, frames, device=local, out=out)
        e1.record()
        torch.cuda.synchronize()
        blocks.append(e0.elapsed_time(e1) / reps)
    ms = min(blocks)
    algo = frames * (pf.packets_per_frame * pf.lidar_packet_size + H * W * 4 + W * (8 + 4 + 2))
    peak, peak_src = measured_peaks()
    # (b) packet by packet through the batcher
    lib = _ffi.load()
    h = C.c_void_p()
    assert lib.ptk_batcher_create(C.byref(h), local, C.byref(pf.c), 4) == 0
    ls = ingest.DeviceLidarScan(H, W, ("RANGE",), local)
    fs = ls._fields_struct()
    ready, n_done = C.c_int(), 0
    base = host.ctypes.data
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(min(host.shape[0], 96 * pf.packets_per_frame)):
        lib.ptk_batcher_push(h, base + i * pf.lidar_packet_size, C.byref(ready))
        if ready.value:
            lib.ptk_batcher_decode(h, C.byref(fs), None, None, None)
            n_done += 1
    lib.ptk_batcher_flush(h, C.byref(ready))
    if ready.value:
        lib.ptk_batcher_decode(h, C.byref(fs), None, None, None)
        n_done += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    lib.ptk_batcher_destroy(h)
    return {"value": frames / (ms * 1e-3), "unit": "scans/s", "frames_per_launch": frames, "ms_per_launch": ms,
            "timing": f"CUDA events around {reps} launches, best of 5 such blocks (ms per launch of each block: "
                      f"{[round(b, 4) for b in blocks]})",
            "roofline": {"bound": "hbm", "kernel": "k_decode_packets", "achieved": algo / (ms * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": algo / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo},
            "e2e": {"value": n_done / dt, "unit": "scans/s", "h2d_bytes_per_scan": pf.packets_per_frame * pf.lidar_packet_size,
                    "api": "ptk_batcher_push per packet (host bytes) + ptk_batcher_decode per frame, host wall clock"},
            "workload": "OS0-128 1024x10 LEGACY lidar packets (64 x 24896 B per scan) -> staggered RANGE image + column headers"}


def side_sharded(world, rank, local, scans_n=12):
    """configs[4], second half: ONE sequence whose voxel map is sharded by hash key over the ranks.  Every rank
    steps the same scans; poses must equal those of an unsharded single-GPU run (rank 0 checks)."""
    import torch
    import torch.distributed as dist
    from ptudes_lab_b200 import odometry, sharded, synth
    dev = torch.device("cuda", local)
    seq = synth.make_sequence("os0_quad", 0)
    gen = synth.TorchScanGenerator(seq, dev)
    ranges = [gen.range_image(s)[0].contiguous() for s in range(scans_n)]
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    warm = 3
    so = sharded.make_sharded(cfg, local, rank, world, max_points=131072, map_capacity=32768, dirs=seq.dirs)
    try:
        poses = []
        t0 = None
        for s in range(scans_n):
            if s == warm:
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            p, _ = so.register_frame(None, None, range_mm=ranges[s])
            poses.append(p)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        info = so.describe()
    finally:
        so.close()
    out = {"value": (scans_n - warm) / float(dt.item()), "unit": UNIT, "ranks": world, "scans": scans_n - warm,
           "timing": "host wall clock, max over ranks", **info}
    if rank == 0:
        o1 = odometry.Odometry(cfg, device=local, max_points=131072, map_capacity=32768, batch=1)
        o1.set_sensor(seq.dirs)
        ref = [o1.register_scan(r)[0] for r in ranges]
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        o1.reset()
        for r in ranges:
            o1.register_scan(r)
        torch.cuda.synchronize()
        out["single_gpu_value"] = scans_n / (time.perf_counter() - t1)
        o1.close()
        out["bit_identical_to_single_gpu"] = bool(all(np.array_equal(a, b) for a, b in zip(poses, ref)))
    return out


# --------------------------------------------------------------------------------------
def run_ptk(args):
    import torch
    import torch.distributed as dist
    from ptudes_lab_b200 import _ffi, odometry, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the ptk path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # one rank = one GPU = its own slice of host cores: a step has ~0.1 ms of host work between the pose
        # read-back and the next launches, and ranks migrating over each other's cores show up in the max over ranks
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local * per:(local + 1) * per] or cores
            os.sched_setaffinity(0, mine)
        except Exception:
            pass

    B, K, W = args.lanes, args.steps, args.warmup
    G = max(1, min(args.contexts, B))
    T = W + K
    min_r, max_r, max_pts, map_cap = CONFIGS[args.config]

    # ---- synthetic scans, generated on the device (data plumbing) -------------------------
    # the scan as KissICPWrapper.register_frame receives it (kiss.py:54-61): the RANGE image (H, W) uint32 mm;
    # projection + mask + column timestamps run inside the step.
    use_range = True
    gens = [synth.TorchScanGenerator(synth.make_sequence(args.config, rank * B + l), dev) for l in range(B)]
    ranges = [[None] * B for _ in range(T)]
    for l, g in enumerate(gens):
        for s in range(T):
            ranges[s][l] = g.range_image(s)[0].contiguous()
    torch.cuda.synchronize()
    scan_bytes = sum(r.numel() * 4 for r in ranges[W])
    total_bytes = sum(r.numel() * 4 for s in range(T) for r in ranges[s])
    n_points = int((ranges[W][0] != 0).sum().item())

    cfg = odometry.load_config(None, deskew=True, max_range=max_r)
    cfg.data.min_range = min_r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The B sequences of this rank are dealt to G contexts.  The lanes of one context advance in lock step (one
    # batched call = one step of all of them); the contexts are NOT synchronised with each other: ptk_fleet_replay
    # runs each on its own host thread and stream, so the latency-bound ICP loop of one context overlaps the
    # streaming kernels of the others.  G = 1 is the plain lock-step batch.
    parts = [list(range(g * B // G, (g + 1) * B // G)) for g in range(G)]
    odos = [odometry.Odometry(cfg, device=local, max_points=max_pts, map_capacity=map_cap, batch=len(p)) for p in parts]
    for o in odos:
        o.set_sensor(gens[0].seq.dirs)
        if G > 1:
            o.set_icp_blocks_per_lane(args.icp_blocks)
    streams = [torch.cuda.Stream(device=dev) for _ in odos]
    sh = [st.cuda_stream for st in streams]
    stream = streams[0]

    def replay(images, lo, hi, want_stats=False):
        rg = [[[images[s][l] for l in parts[g]] for s in range(lo, hi)] for g in range(G)]
        return odometry.fleet_replay(odos, rg, sh, want_stats=want_stats)

    def merge(per_ctx):          # [g] (n, b_g, 4, 4) -> (n, B, 4, 4)
        out = np.empty((per_ctx[0].shape[0], B, 4, 4))
        for g, pz in enumerate(per_ctx):
            out[:, parts[g]] = pz
        return out

    icp_phase_cycles = None

    def timed_run(images, profiling, clocks=None):
        """W untimed scans, then K timed ones: one ptk_fleet_replay call each.  Device time = from an event recorded
        on every context's stream before the call to the last of the events recorded on them after it."""
        for o in odos:
            o.reset()
        poses_w = merge(replay(images, 0, W))
        if profiling:
            for o in odos:
                o.set_profiling(True)
        l0 = sum(o.launch_count() for o in odos)
        barrier()
        c0 = clocks.mark() if clocks else 0
        e0 = [torch.cuda.Event(enable_timing=True) for _ in odos]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in odos]
        for ev, st in zip(e0, streams):
            ev.record(st)
        res = replay(images, W, T, want_stats=profiling)
        for ev, st in zip(e1, streams):
            ev.record(st)
        barrier()
        c1 = clocks.mark() if clocks else 0
        ms = max(e0[0].elapsed_time(ev) for ev in e1)
        launches = sum(o.launch_count() for o in odos) - l0
        prof, stats_acc = None, None
        if profiling:
            nonlocal icp_phase_cycles
            poses_t, stats_ctx = res
            icp_phase_cycles = odos[0].icp_phases(0)
            prof = {}
            for o in odos:
                for k, (ms_k, n_k) in o.get_profile().items():
                    a = prof.get(k, (0.0, 0))
                    prof[k] = (a[0] + ms_k, a[1] + n_k)
                o.set_profiling(False)
            stats_acc = [[stats_ctx[g][s][i] for g in range(G) for i in range(len(parts[g]))] for s in range(K)]
        else:
            poses_t = res
        return ms, launches, prof, stats_acc, np.concatenate([poses_w, merge(poses_t)]), (c0, c1)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)

    # ---- leg 1: inputs resident in HBM ----------------------------------------------------
    ms_dev, launches, _, _, poses_dev, cspan = timed_run(ranges, False, clocks)
    # per-kernel device times: same scans again with every launch bracketed by CUDA events on
    # the launching stream (separate pass so the events do not sit inside the headline number)
    ms_prof, _, prof, stats_acc, poses_prof, _ = timed_run(ranges, True)

    # ---- leg 2: end to end from pinned host buffers ----------------------------------------
    # every timed scan is copied host -> device inside the timed call (the first one synchronously, the others
    # prefetched one scan ahead on a side stream while the previous scan computes): exactly K copies per lane
    e2e = None
    h_frames = None
    if not args.no_e2e:
        h_frames = [[None] * B for _ in range(T)]
        for s in range(T):
            for l in range(B):
                hr = _ffi.pinned_empty(tuple(ranges[s][l].shape), dtype=np.uint32)
                hr[...] = ranges[s][l].cpu().numpy().astype(np.uint32)
                h_frames[s][l] = hr
        ms_e2e, _, _, _, poses_e2e, _ = timed_run(h_frames, False)
        if not np.array_equal(poses_e2e, poses_dev):
            raise SystemExit("bench.py: host-buffer and device-buffer runs disagree")
        import ctypes
        cb_in, cb_out = ctypes.c_int(0), ctypes.c_int(0)
        _ffi.load().ptk_control_bytes(ctypes.byref(cb_in), ctypes.byref(cb_out))
        h2d = scan_bytes + B * cb_in.value          # scans + the per-lane parameter records
        d2h = B * cb_out.value                      # per-lane result records (pose + counters)
    if not np.array_equal(poses_prof, poses_dev):
        raise SystemExit("bench.py: run-to-run poses differ (non-deterministic step)")

    # ---- the same scans as ONE sequence (configs[1] literally): latency-bound, reported beside the batch
    single = None
    if B > 1:
        odo1 = odometry.Odometry(cfg, device=local, max_points=max_pts, map_capacity=map_cap, batch=1)
        odo1.set_sensor(gens[0].seq.dirs)

        def run1(inp):
            odo1.reset()
            for s_ in range(W):
                odo1.register_scan(inp[s_][0], stream=sh[0])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for s_ in range(W, T):
                odo1.register_scan(inp[s_][0], stream=sh[0])
            torch.cuda.synchronize()
            return K / (time.perf_counter() - t0)
        single = {"value": run1(ranges), "unit": UNIT,
                  "note": "one sequence, one scan per step (each scan waits for the previous pose): host wall clock"}
        if not args.no_e2e:
            single["e2e"] = run1(h_frames)
        odo1.close()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the other BASELINE.json configs, small runs (skipped with --no-side-runs) -----------
    side = {}
    for o in odos:
        o.close()
    odos = []
    if not args.no_side_runs:
        # configs[4], first half, LITERALLY: 64 sequences in total, dealt to the ranks (strong scaling: 64/N per GPU)
        if 64 % world == 0:
            per = 64 // world
            g64 = max(1, min(G, per // 4))
            ms64, _ = side_fleet("os0_quad", per, 3, 10, local, rank * per, barrier, contexts=g64, icp_blocks=args.icp_blocks)
            ms64 = max_over_ranks(ms64)
            side["fleet64_strong"] = {"value": 64 * 10 / (ms64 * 1e-3), "unit": UNIT, "sequences_total": 64,
                                      "lanes_per_gpu": per, "contexts_per_gpu": g64, "steps": 10, "ms_per_step": ms64 / 10,
                                      "scaling": "strong",
                                      "workload": "configs[4]: 64 independent os0_quad sequences over the ranks"}
        if world > 1:
            try:
                side["sharded_single_sequence"] = side_sharded(world, rank, local)
            except Exception as e:      # never lose the headline line to a side run
                side["sharded_single_sequence"] = {"error": f"{type(e).__name__}: {e}"}
            dist.barrier()
        if rank == 0:
            ms2, c2 = side_fleet("os2_street", 16, 3, 10, local, 0, torch.cuda.synchronize, contexts=4, icp_blocks=args.icp_blocks)
            side["os2_street"] = {"value": 16 * 10 / (ms2 * 1e-3), "unit": UNIT, "lanes": 16, "contexts": 4, "steps": 10,
                                  "ms_per_step": ms2 / 10, "counters": c2,
                                  "workload": "configs[2]: OS2-128 2048x10, 262144 pixels/scan, max_range 200 m (voxel 2 m)"}
            side["ekf_bench"] = side_ekf_fleet(16, 3, 10, local)
            try:
                side["ingest"] = side_ingest(local)
            except Exception as e:      # never lose the headline line to a side run
                side["ingest"] = {"error": f"{type(e).__name__}: {e}"}
        if world > 1:
            dist.barrier()

    ms_dev_max = max_over_ranks(ms_dev)
    value = world * B * K / (ms_dev_max * 1e-3)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"configs[1] 100-scan OS0-128 1024x10 sequence shape ({args.config}), full odometry step, "
                        f"{B} independent sequences (lanes) per GPU in {G} context(s) of {[len(p) for p in parts]} lanes; a "
                        f"step = every lane advanced by one scan (one batched call per context, contexts free-running "
                        f"on their own host threads: ptk_fleet_replay); scans {W}..{W + K - 1} of each sequence timed",
            "lanes_per_gpu": B, "contexts_per_gpu": G, "icp_blocks_per_lane": (args.icp_blocks if G > 1 else "all"),
            "points_per_scan": n_points, "input": args.input, "max_range": max_r,
            "min_range": min_r, "voxel_size": cfg.mapping.voxel_size,
            "l2": f"every step reads scans never touched before ({total_bytes / 1e9:.2f} GB of scans per GPU, "
                  f"{scan_bytes / 1e6:.1f} MB per step); the local maps are persistent state and stay wherever "
                  f"the hardware keeps them",
            "parallelism": f"fleet replay: {world} GPU(s) x {B} lanes, no collective on the data path",
        },
        "gpu_launches": int(launches),
    }
    if e2e is None and not args.no_e2e:
        ms_e2e_max = max_over_ranks(ms_e2e)
        out["e2e"] = {"value": world * B * K / (ms_e2e_max * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                      "ms_per_step": ms_e2e_max / K,
                      "api": "ptk_fleet_replay -> ptk_register_scan_batch per context and scan (C ABI, ctypes) with pinned "
                             "host RANGE images (uint32 mm); poses and counters of every scan read back to the host"}

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peak, peak_src = measured_peaks()
    algo = {}
    for step_stats in stats_acc:
        for st in step_stats:
            for k, v in algorithmic_bytes(st, int(ranges[W][0].numel()) if use_range else 0).items():
                algo[k] = algo.get(k, 0) + v
    kern = {k: v for k, v in prof.items() if v[1] > 0 and k in algo}
    dom = max(kern, key=lambda k: kern[k][0])
    dms, dn = kern[dom]
    per_launch_bytes = algo[dom] / dn
    achieved = per_launch_bytes / (dms / dn * 1e-3) / 1e9
    total_algo = sum(algo.values())
    out["roofline"] = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": recorded_traffic(dom, args.config, B, args.input, G), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": per_launch_bytes, "avg_launch_ms": dms / dn,
        "kernel_share_of_step": dms / sum(v[0] for v in prof.values()),
        "step_algorithmic_bytes_per_scan": total_algo / (B * K),
        "step_hbm_frac": (total_algo / (ms_dev * 1e-3) / 1e9) / peak,
        "kernels_ms_per_step": {k: v[0] / K for k, v in prof.items() if v[1] > 0},
    }
    if G > 1:
        # the G contexts' launches of the kernel overlap in time: a launch's duration is the time it SHARES the GPU
        # with G - 1 others, so the per-launch figure above is a lower bound.  What the kernel moves as a whole:
        # all its algorithmic bytes of the timed region over the region's duration (it is running somewhere on the
        # GPU for most of it).
        agg = algo[dom] / (ms_dev * 1e-3) / 1e9
        out["roofline"]["concurrent_launches"] = G
        out["roofline"]["achieved_all_launches"] = agg
        out["roofline"]["frac_all_launches"] = agg / peak
    last = stats_acc[-1][0]
    out["counters"] = {k: last[k] for k in ("n_in", "n_range", "n_ds", "n_src", "n_voxels", "map_points",
                                             "iterations", "n_corr")}
    if single is not None:
        out["single_sequence"] = single
    if side:
        out["side_runs"] = side
    out["counters"]["mean_icp_iterations"] = float(np.mean([st["iterations"] for ss in stats_acc for st in ss]))
    queries = sum(st["iterations"] * st["n_src"] for ss in stats_acc for st in ss)
    searches = sum(st["icp_searches"] for ss in stats_acc for st in ss)
    out["icp"] = {"nn_queries": int(queries), "full_searches": int(searches),
                  "cache_hit_rate": 1.0 - searches / max(queries, 1),
                  "block0_phase_cycles_last_scan": icp_phase_cycles}

    if rank == 0:
        time.sleep(0.2)
        clocks.stop()
        out["clocks"] = clocks.summary(*cspan)

    # ---- cpu baseline: the oracle on this box's host cores (rank 0, N = 1 only) -------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        C = max(1, min(B, os.cpu_count() or 1))
        h = [[ranges[s][l].cpu().numpy().astype(np.uint32) for s in range(T)] for l in range(C)]
        out["cpu_baseline"], ref_poses = cpu_baseline_leg(args, h, gens[0].seq.dirs, min_r, max_r, T)
        n = min(len(p) for p in ref_poses)
        ref = np.stack([np.stack(p[:n]) for p in ref_poses], axis=1) if n else None       # (n, C, 4, 4)
        out["parity_vs_oracle"] = {"scans": n, "lanes": C,
                                   "max_abs_pose_diff": float(np.abs(ref - poses_dev[:n, :C]).max()) if n else None}
        if single is not None:      # configs[1] literally on the CPU: ONE sequence, OpenMP over all cores inside it
            v1, n1, s1, _ = cpu_port_fleet(h[:1], gens[0].seq.dirs, min_r, max_r, W, min(args.cpu_seconds, 8.0),
                                           os.cpu_count() or 1)
            single["cpu_port"] = {"value": v1, "unit": UNIT, "cores": os.cpu_count() or 1,
                                  "sample": f"one sequence, {n1} scans, OpenMP threads = cores ({s1:.1f} s)"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for o in odos:
        o.close()
    if rank == 0:
        print(json.dumps(out), file=RESULT, flush=True)


def cpu_port_fleet(scans, dirs, min_r, max_r, warmup, budget_s, cores):
    """The CPU restatement of the reference path on host cores, charged what KissICPWrapper.register_frame
    does per scan: the NumPy projection of the RANGE image (`sel = range != 0; xyz = xyz_lut(scan)[sel];
    timestamps = self._timestamps[sel]`, kiss.py:59-61 = synth.project_scan) and then the kiss-icp step
    (oracle/kiss_port.c: plain C + OpenMP, upstream kiss-icp's structure, gcc -O3 -march=native built on this
    host).  `scans[l][s]` = (H, W) uint32 range image of lane l, scan s.  The lanes are independent sequences, so
    the best use of the cores is one sequence per thread (NumPy and the C call release the GIL); with fewer
    lanes than cores the spare cores go to OpenMP inside a lane.  A step advances every lane by one scan, like
    the CUDA arm.  Returns (scans/s, steps timed, seconds, poses[lane][scan])."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import port
    from ptudes_lab_b200 import synth
    flags = port.use_native()
    L = len(scans)
    inner = max(1, cores // L)
    ctxs = [port.PortKissICP(_min_range=min_r, _max_range=max_r, threads=inner) for _ in range(L)]
    poses = [[] for _ in range(L)]

    def advance(args):
        l, s = args
        xyz, ts = synth.project_scan(scans[l][s], dirs)              # kiss.py:59-61, timed
        poses[l].append(ctxs[l].register_points(xyz, ts, 0.1 * (s + 1)).copy())

    T = len(scans[0])
    t_used, n_steps = 0.0, 0
    with ThreadPoolExecutor(max_workers=min(L, cores)) as ex:
        for s in range(T):
            t0 = time.perf_counter()
            list(ex.map(advance, [(l, s) for l in range(L)]))
            dt = time.perf_counter() - t0
            if s >= warmup:
                t_used += dt
                n_steps += 1
                if t_used > budget_s:
                    break
    for c in ctxs:
        c.close()
    cpu_port_fleet.flags = flags
    return (L * n_steps / t_used if t_used > 0 else None), n_steps, t_used, poses


def cpu_baseline_leg(args, host_scans, dirs, min_r, max_r, T):
    """cpu_baseline of the own arm: lane 0..C-1 of this very workload on the box's host cores,
    bounded by --cpu-seconds.  The oracle is used here as the reported baseline only."""
    cores = os.cpu_count() or 1
    val, n_steps, secs, poses = cpu_port_fleet(host_scans, dirs, min_r, max_r, args.warmup, args.cpu_seconds, cores)
    L = len(host_scans)
    up = probe_upstream()
    return ({"value": val, "unit": UNIT, "cores": cores, "kind": "port",
             "upstream_kiss_icp": up if up else "not importable on this box (nor under baseline/_ref): the port stands in",
             "sample": f"NumPy projection of the RANGE image (kiss.py:59-61) + oracle/kiss_port.c (C restatement of the "
                       f"kiss-icp 0.2.x step, gcc {cpu_port_fleet.flags} + OpenMP): {L} of the lanes' sequences run "
                       f"concurrently, one per host thread, scans {args.warmup}..{args.warmup + n_steps - 1} timed "
                       f"({secs:.1f} s of CPU wall time)"}, poses)


# --------------------------------------------------------------------------------------
def probe_upstream():
    """SURVEY 8c run-time probe: version of the REAL kiss-icp if this box has it (environment or baseline/_ref)."""
    for extra in (None, os.path.join(ROOT, "baseline", "_ref")):
        if extra is not None:
            if not os.path.isdir(extra):
                continue
            if extra not in sys.path:
                sys.path.insert(0, extra)
        try:
            import kiss_icp
            return "kiss-icp " + str(getattr(kiss_icp, "__version__", "?"))
        except Exception:
            continue
    return None


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path on the box's host cores.
    kiss-icp 0.2.x is not installable here and /root/reference holds no compilable source for the
    path, so this is the C port of oracle/ with every host thread in use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from ptudes_lab_b200 import synth
    K, W = args.steps, args.warmup
    min_r, max_r, _, _ = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    L = max(1, min(args.lanes, cores))
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    scans = []
    dirs = None
    for l in range(L):          # same sequences as the CUDA arm's lanes 0..L-1 (synthetic-data plumbing)
        g = synth.TorchScanGenerator(synth.make_sequence(args.config, l), dev)
        dirs = g.seq.dirs
        scans.append([g.range_image(s)[0].cpu().numpy().astype(np.uint32) for s in range(W + K)])
    val, n, secs, _ = cpu_port_fleet(scans, dirs, min_r, max_r, W, 240.0, cores)
    sample = (f"NumPy projection of the RANGE image (kiss.py:59-61) + oracle/kiss_port.c (C restatement of the kiss-icp "
              f"0.2.x step, gcc {cpu_port_fleet.flags} + OpenMP) on {cores} host cores: {L} of the workload's "
              f"{args.lanes} sequences per step, one per thread; {n} of the requested {K} steps timed")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n,
           "warmup": W, "ms_per_step": 1e3 * secs / max(n, 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"configs[1] 100-scan OS0-128 1024x10 sequence shape ({args.config}), full "
                                  f"odometry step on the CPU, {L} independent sequences per step",
                      "lanes": L},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), file=RESULT, flush=True)


def _guard_stdout():
    """Libraries (NCCL's version banner, for one) write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the whole run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


if __name__ == "__main__":
    RESULT = _guard_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ptk(a)
