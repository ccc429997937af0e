"""Python face of libptk: one class per kiss-icp object the reference touches.

`Odometry` owns a ptk context (= kiss_icp.kiss_icp.KissICP: poses, adaptive threshold, local map);
the free functions / small classes mirror the kiss-icp 0.2.x Python API that
/root/reference/src/ptudes/kiss.py imports or reaches through `self._kiss`:
`load_config`, `KissICP.{poses, config, compensator, preprocess, voxelize,
get_adaptive_threshold, get_prediction_model, adaptive_threshold, local_map}` and
`kiss_icp.registration.register_frame`.

All arithmetic runs in the CUDA library; arrays cross as raw pointers (numpy = host,
torch.cuda tensors = device).  No CPU fallback exists.
"""
import ctypes as C

import numpy as np

from . import _ffi
from ._ffi import PtkConfig, PtkError, PtkStats, addr


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __repr__(self):
        return "Config(" + ", ".join(f"{k}={v!r}" for k, v in self.__dict__.items()) + ")"


def load_config(config_file=None, deskew=False, max_range=100.0):
    """kiss_icp.config.load_config defaults; voxel_size = max_range / 100."""
    if config_file is not None:
        raise NotImplementedError("config files are outside the hot path; pass keyword options")
    return _NS(
        data=_NS(preprocess=True, correct_scan=True, max_range=float(max_range), min_range=5.0,
                 deskew=bool(deskew)),
        mapping=_NS(voxel_size=float(max_range) / 100.0, max_points_per_voxel=20),
        adaptive_threshold=_NS(fixed_threshold=None, initial_threshold=2.0, min_motion_th=0.1),
    )


def _mat16(T):
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).reshape(4, 4))


class StatsBatch:
    """The per-lane ptk_stats of one batched step.  Behaves like a list of dicts, but the dicts are only
    built when asked for: converting 48 structs x 16 fields eagerly costs more host time per step than
    the launches do, and the GPU idles meanwhile."""

    def __init__(self, raw):
        self._raw = raw

    def __len__(self):
        return len(self._raw)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._raw[k].as_dict() for k in range(*i.indices(len(self._raw)))]
        return self._raw[i].as_dict()

    def __iter__(self):
        return (self._raw[k].as_dict() for k in range(len(self._raw)))

    def field(self, name):
        """One field of every lane as a list (no dicts)."""
        return [getattr(self._raw[k], name) for k in range(len(self._raw))]


class Odometry:
    """A ptk context with `batch` lanes.  Lane 0 is the default sequence."""

    def __init__(self, config=None, *, device=0, max_points=262144, map_capacity=262144, batch=1,
                 trace_iterations=0, max_iterations=500):
        self._lib = _ffi.load()
        self.config = config if config is not None else load_config()
        c = PtkConfig()
        self._lib.ptk_default_config(C.byref(c))
        c.max_range = self.config.data.max_range
        c.min_range = self.config.data.min_range
        c.voxel_size = self.config.mapping.voxel_size
        c.max_points_per_voxel = self.config.mapping.max_points_per_voxel
        c.deskew = 1 if self.config.data.deskew else 0
        c.initial_threshold = self.config.adaptive_threshold.initial_threshold
        c.min_motion_th = self.config.adaptive_threshold.min_motion_th
        c.max_iterations = max_iterations
        c.max_points = max_points
        c.map_capacity = map_capacity
        c.batch = batch
        c.trace_iterations = trace_iterations
        self._cfg = c
        self.batch = batch
        self.max_points = max_points
        self.map_capacity = map_capacity
        h = C.c_void_p()
        rc = self._lib.ptk_ctx_create(C.byref(h), device, C.byref(c))
        if rc != 0:
            raise PtkError(rc, self._lib.ptk_last_error(None).decode())
        self._h = h

    # -- plumbing ------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise PtkError(rc, self._lib.ptk_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ptk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, lane=-1):
        self._check(self._lib.ptk_reset(self._h, lane))

    # -- measurement taps ----------------------------------------------------------
    def set_profiling(self, on=True):
        self._check(self._lib.ptk_set_profiling(self._h, 1 if on else 0))

    def get_profile(self):
        """{kernel name: (device ms, launches)} accumulated since set_profiling()."""
        ms = np.zeros(12)
        n = np.zeros(12, dtype=np.int64)
        self._check(self._lib.ptk_get_profile(self._h, addr(ms), addr(n)))
        out = {}
        for k in range(12):
            name = self._lib.ptk_kernel_name(k)
            if name is None:
                break
            out[name.decode()] = (float(ms[k]), int(n[k]))
        return out

    def icp_phases(self, lane=0):
        """clock64 cycles block 0 of the lane's last ICP launch spent per phase (measurement tap)."""
        out = np.zeros(6, dtype=np.int64)
        self._check(self._lib.ptk_get_icp_phases(self._h, lane, addr(out)))
        return dict(zip(("cache_pass", "searches", "sums", "barrier", "tree", "solve"), out.tolist()))

    def launch_count(self):
        return int(self._lib.ptk_launch_count(self._h))

    def set_icp_blocks_per_lane(self, blocks: int):
        """Cap the blocks per lane of the cooperative ICP launch (0: as many as the device holds), so that the
        ICP launches of several contexts fit the device side by side (fleet replay over several contexts)."""
        self._check(self._lib.ptk_set_icp_blocks_per_lane(self._h, int(blocks)))

    # -- the step ------------------------------------------------------------------
    def register_frame(self, frame, timestamps, initial_guess=None, lane=0, stream=0):
        """One odometry step (kiss.py:83-131).  Returns (pose 4x4, stats dict)."""
        frame = _ffi.f64(frame, 3)
        timestamps = _ffi.f64(timestamps)
        n = int(frame.shape[0])
        if int(np.prod(tuple(timestamps.shape))) != n:
            raise TypeError(f"timestamps: {tuple(timestamps.shape)} for {n} points")
        g = _mat16(initial_guess) if initial_guess is not None else None
        pose = np.empty((4, 4))
        st = PtkStats()
        rc = self._lib.ptk_register_frame(self._h, lane, addr(frame), addr(timestamps), n, addr(g),
                                          addr(pose), C.byref(st), stream)
        self._check(rc)
        return pose, st.as_dict()

    def register_frame_batch(self, frames, timestamps, guesses=None, stream=0):
        """Advance every lane by one scan in one set of launches."""
        B = self.batch
        assert len(frames) == B and len(timestamps) == B
        frames = [_ffi.f64(f, 3) for f in frames]
        timestamps = [_ffi.f64(t) for t in timestamps]
        xs = (C.c_void_p * B)(*[addr(f) for f in frames])
        ts = (C.c_void_p * B)(*[addr(t) for t in timestamps])
        ns = (C.c_int * B)(*[int(f.shape[0]) for f in frames])
        gbuf, hg = None, None
        if guesses is not None:
            gbuf = np.zeros((B, 4, 4))
            flags = bytearray(B)
            for i, g in enumerate(guesses):
                if g is not None:
                    gbuf[i] = _mat16(g)
                    flags[i] = 1
            hg = bytes(flags)
        poses = np.empty((B, 4, 4))
        stats = (PtkStats * B)()
        rc = self._lib.ptk_register_frame_batch(self._h, xs, ts, ns, addr(gbuf), hg, addr(poses), stats, stream)
        self._check(rc)
        return poses, StatsBatch(stats)

    # -- the step fed with range images (kiss.py:54-74 incl. the XYZLut projection) ---------
    def set_sensor(self, direction, offset=None, col_timestamps=None, range_unit=0.001):
        """client.XYZLut + the timestamp table of KissICPWrapper.__init__ (kiss.py:28-35).
        direction (H, W, 3); xyz = direction * (range * range_unit) [+ offset]."""
        d = np.ascontiguousarray(direction, dtype=np.float64)
        H, W = int(d.shape[0]), int(d.shape[1])
        o = None if offset is None else np.ascontiguousarray(offset, dtype=np.float64).reshape(H, W, 3)
        t = None if col_timestamps is None else np.ascontiguousarray(col_timestamps, dtype=np.float64).reshape(W)
        self._check(self._lib.ptk_set_sensor(self._h, H, W, addr(d), addr(o), addr(t), float(range_unit)))
        self.sensor_shape = (H, W)

    def _u32(self, a):
        """RANGE image for the C ABI: a torch tensor goes by address (int32/uint32 storage, contiguous, H*W
        elements when the sensor is known); anything else becomes a contiguous uint32 ndarray."""
        if hasattr(a, "data_ptr") and not isinstance(a, np.ndarray):
            _ffi._check_tensor(a, ("int32", "uint32"), "range image")
            shape = getattr(self, "sensor_shape", None)
            if shape is not None and int(a.numel()) != shape[0] * shape[1]:
                raise TypeError(f"range image: {int(a.numel())} elements, sensor is {shape[0]}x{shape[1]}")
            return a
        a = np.ascontiguousarray(a, dtype=np.uint32)
        shape = getattr(self, "sensor_shape", None)
        if shape is not None and a.size != shape[0] * shape[1]:
            raise TypeError(f"range image: {a.size} elements, sensor is {shape[0]}x{shape[1]}")
        return a

    def register_scan(self, range_mm, initial_guess=None, lane=0, stream=0):
        """One odometry step from the RANGE field (H, W) uint32 mm.  Returns (pose 4x4, stats)."""
        r = self._u32(range_mm)
        g = _mat16(initial_guess) if initial_guess is not None else None
        pose = np.empty((4, 4))
        st = PtkStats()
        self._check(self._lib.ptk_register_scan(self._h, lane, addr(r), addr(g), addr(pose), C.byref(st), stream))
        return pose, st.as_dict()

    def prefetch_scan_batch(self, ranges):
        """Start the H2D copy of the next step's (pinned host) range images; see ptk_prefetch_scan_batch."""
        B = self.batch
        assert len(ranges) == B
        self._pf_keep = [None if r is None else self._u32(r) for r in ranges]     # keep the arrays alive
        ptrs = (C.c_void_p * B)(*[None if r is None else addr(r) for r in self._pf_keep])
        self._check(self._lib.ptk_prefetch_scan_batch(self._h, ptrs))

    def register_scan_batch(self, ranges, guesses=None, stream=0):
        B = self.batch
        assert len(ranges) == B
        rs = [self._u32(r) for r in ranges]
        ptrs = (C.c_void_p * B)(*[addr(r) for r in rs])
        gbuf, hg = None, None
        if guesses is not None:
            gbuf = np.zeros((B, 4, 4))
            flags = bytearray(B)
            for i, g in enumerate(guesses):
                if g is not None:
                    gbuf[i] = _mat16(g)
                    flags[i] = 1
            hg = bytes(flags)
        poses = np.empty((B, 4, 4))
        stats = (PtkStats * B)()
        self._check(self._lib.ptk_register_scan_batch(self._h, ptrs, addr(gbuf), hg, addr(poses), stats, stream))
        return poses, StatsBatch(stats)

    # -- KissICP state -------------------------------------------------------------
    def num_poses(self, lane=0):
        return self._lib.ptk_num_poses(self._h, lane)

    def get_pose(self, index=-1, lane=0):
        out = np.empty((4, 4))
        self._check(self._lib.ptk_get_pose(self._h, lane, index, addr(out)))
        return out

    def get_prediction_model(self, lane=0):
        out = np.empty((4, 4))
        self._check(self._lib.ptk_get_prediction_model(self._h, lane, addr(out)))
        return out

    def get_adaptive_threshold(self, lane=0):
        """KissICP.get_adaptive_threshold() (kiss.py:99); like upstream it accumulates the last model deviation."""
        v = C.c_double()
        self._check(self._lib.ptk_get_adaptive_threshold(self._h, lane, C.byref(v)))
        return v.value

    def update_model_deviation(self, T, lane=0):
        """adaptive_threshold.update_model_deviation(T) (kiss.py:128)."""
        self._check(self._lib.ptk_update_model_deviation(self._h, lane, addr(_mat16(T))))

    def append_pose(self, T, lane=0):
        """KissICP.poses.append(T) (kiss.py:130) for the library's own pose list (prediction model, has_moved)."""
        self._check(self._lib.ptk_append_pose(self._h, lane, addr(_mat16(T))))

    def last_sigma(self, lane=0):
        return self._lib.ptk_last_sigma(self._h, lane)

    # -- pieces --------------------------------------------------------------------
    def deskew_scan(self, frame, timestamps, start_pose, finish_pose, stream=0):
        frame = _ffi.f64(frame)
        timestamps = _ffi.f64(timestamps)
        n = int(frame.shape[0])
        out = np.empty((n, 3))
        self._check(self._lib.ptk_deskew_scan(self._h, addr(frame), addr(timestamps), n, addr(_mat16(start_pose)),
                                              addr(_mat16(finish_pose)), addr(out), stream))
        return out

    def preprocess(self, frame, max_range=None, min_range=None, stream=0):
        frame = _ffi.f64(frame)
        n = int(frame.shape[0])
        out = np.empty((n, 3))
        m = C.c_int(0)
        self._check(self._lib.ptk_preprocess(
            self._h, addr(frame), n, self.config.data.max_range if max_range is None else max_range,
            self.config.data.min_range if min_range is None else min_range, addr(out), C.byref(m), stream))
        return out[:m.value].copy()

    def voxel_down_sample(self, frame, voxel_size, return_index=False, stream=0):
        frame = _ffi.f64(frame)
        n = int(frame.shape[0])
        out = np.empty((n, 3))
        idx = np.empty(n, dtype=np.int32)
        m = C.c_int(0)
        self._check(self._lib.ptk_voxel_down_sample(self._h, addr(frame), n, float(voxel_size), addr(out),
                                                    addr(idx), C.byref(m), stream))
        if return_index:
            return out[:m.value].copy(), idx[:m.value].copy()
        return out[:m.value].copy()

    def voxelize(self, frame):
        v = self.config.mapping.voxel_size
        frame_downsample = self.voxel_down_sample(frame, v * 0.5)
        source = self.voxel_down_sample(frame_downsample, v * 1.5)
        return source, frame_downsample

    # -- taps on the last step -----------------------------------------------------
    def get_points(self, which, lane=0, with_index=False, stream=0):
        cap = self.max_points
        out = np.empty((cap, 3))
        idx = np.empty(cap, dtype=np.int32)
        m = C.c_int(0)
        self._check(self._lib.ptk_get_points(self._h, lane, which, addr(out), addr(idx), cap, C.byref(m), stream))
        if with_index:
            return out[:m.value].copy(), idx[:m.value].copy()
        return out[:m.value].copy()

    def get_frame(self, lane=0, stream=0):
        cap = self.max_points
        out = np.empty((cap, 3))
        m = C.c_int(0)
        self._check(self._lib.ptk_get_frame(self._h, lane, addr(out), cap, C.byref(m), stream))
        return out[:m.value].copy()

    def get_trace(self, lane=0, stream=0):
        ni, ns = C.c_int(0), C.c_int(0)
        self._check(self._lib.ptk_get_trace(self._h, lane, None, 0, C.byref(ni), C.byref(ns), stream))
        out = np.empty((max(ni.value, 1), max(ns.value, 1)), dtype=np.int32)
        self._check(self._lib.ptk_get_trace(self._h, lane, addr(out), ni.value, C.byref(ni), C.byref(ns), stream))
        return out[:ni.value, :ns.value]


class VoxelHashMap:
    """kiss_icp.mapping.VoxelHashMap face of one lane's local map."""

    def __init__(self, odo: Odometry, lane=0):
        self._o = odo
        self._lane = lane

    def clear(self):
        self._o._check(self._o._lib.ptk_map_clear(self._o._h, self._lane, 0))

    def empty(self):
        rc = self._o._lib.ptk_map_empty(self._o._h, self._lane)
        if rc < 0:
            self._o._check(rc)
        return bool(rc)

    def update(self, points, pose):
        points = _ffi.f64(points)
        self._o._check(self._o._lib.ptk_map_update(self._o._h, self._lane, addr(points), int(points.shape[0]),
                                                   addr(_mat16(pose)), 0))

    def add_points(self, points):
        points = _ffi.f64(points)
        self._o._check(self._o._lib.ptk_map_add_points(self._o._h, self._lane, addr(points), int(points.shape[0]), 0))

    def remove_far_away_points(self, origin):
        o = np.ascontiguousarray(origin, dtype=np.float64).reshape(3)
        self._o._check(self._o._lib.ptk_map_remove_far(self._o._h, self._lane, addr(o), 0))

    def counts(self):
        npts, nvox = C.c_int(0), C.c_int(0)
        self._o._check(self._o._lib.ptk_map_num_points(self._o._h, self._lane, C.byref(npts), C.byref(nvox)))
        return npts.value, nvox.value

    def point_cloud(self):
        npts, _ = self.counts()
        out = np.empty((max(npts, 1), 3))
        m = C.c_int(0)
        self._o._check(self._o._lib.ptk_map_point_cloud(self._o._h, self._lane, addr(out), npts, C.byref(m), 0))
        return out[:m.value].copy()

    def dump(self):
        """(keys (V,3) int32, counts (V,), points (V,20,3)) sorted by key for comparisons."""
        _, nvox = self.counts()
        cap = max(nvox, 1)
        keys = np.empty((cap, 3), dtype=np.int32)
        cnt = np.empty(cap, dtype=np.int32)
        pts = np.empty((cap, 20, 3))
        m = C.c_int(0)
        self._o._check(self._o._lib.ptk_map_dump(self._o._h, self._lane, addr(keys), addr(cnt), addr(pts), nvox,
                                                 C.byref(m), 0))
        keys, cnt, pts = keys[:m.value], cnt[:m.value], pts[:m.value]
        order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
        return keys[order], cnt[order], pts[order]

    def get_correspondences(self, points, max_correspondance_distance, return_index=False):
        points = _ffi.f64(points)
        n = int(points.shape[0])
        order = np.empty(max(n, 1), dtype=np.int32)
        tgt = np.empty((max(n, 1), 3))
        nc = C.c_int(0)
        self._o._check(self._o._lib.ptk_map_get_correspondences(
            self._o._h, self._lane, addr(points), n, float(max_correspondance_distance), addr(order), addr(tgt),
            C.byref(nc), 0))
        order, tgt = order[:n], tgt[:n]
        acc = order >= 0
        if return_index:
            return acc, tgt, order
        return np.asarray(points)[acc], tgt[acc]


def register_frame(points, voxel_map: VoxelHashMap, initial_guess, max_correspondance_distance, kernel,
                   return_stats=False):
    """kiss_icp.registration.register_frame (kiss.py:8,108-114)."""
    o = voxel_map._o
    points = _ffi.f64(points)
    pose = np.empty((4, 4))
    st = PtkStats()
    rc = o._lib.ptk_register_point_cloud(o._h, voxel_map._lane, addr(points), int(points.shape[0]),
                                         addr(_mat16(initial_guess)), float(max_correspondance_distance),
                                         float(kernel), addr(pose), C.byref(st), 0)
    o._check(rc)
    if return_stats:
        return pose, st.as_dict()
    return pose


def fleet_replay(odos, ranges, streams, want_stats=False):
    """ptk_fleet_replay: contexts `odos` (each advanced by its own host thread inside the library, on the
    non-default stream handle streams[g]) through the scans `ranges[g][s][l]` = RANGE image of scan s, lane l of
    context g (pinned host arrays or device tensors).  Returns poses[g] (n_scans, batch_g, 4, 4) [, stats[g][s]]."""
    lib = odos[0]._lib
    G = len(odos)
    T = len(ranges[0])
    keep, ptr_arrays = [], []
    for g, o in enumerate(odos):
        assert len(ranges[g]) == T
        flat = [o._u32(r) for s in range(T) for r in ranges[g][s]]
        assert len(flat) == T * o.batch
        keep.append(flat)
        ptr_arrays.append((C.c_void_p * len(flat))(*[addr(r) for r in flat]))
    rng = (C.POINTER(C.c_void_p) * G)(*[C.cast(a, C.POINTER(C.c_void_p)) for a in ptr_arrays])
    poses = [np.empty((T, o.batch, 4, 4)) for o in odos]
    pp = (C.c_void_p * G)(*[addr(p) for p in poses])
    stats = [(PtkStats * (T * o.batch))() for o in odos] if want_stats else None
    sp = (C.POINTER(PtkStats) * G)(*[C.cast(s_, C.POINTER(PtkStats)) for s_ in stats]) if want_stats else None
    hs = (C.c_void_p * G)(*[o._h for o in odos])
    st = (C.c_void_p * G)(*[int(x) for x in streams])
    rc = lib.ptk_fleet_replay(hs, G, rng, T, pp, sp, st)
    if rc != 0:
        for o in odos:
            o._check(rc)
    if want_stats:
        out_stats = [[StatsBatch((PtkStats * o.batch).from_buffer(stats[g], s * o.batch * C.sizeof(PtkStats))) for s in range(T)]
                     for g, o in enumerate(odos)]
        return poses, out_stats
    return poses
