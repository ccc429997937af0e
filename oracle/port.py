"""ctypes face of oracle/libkiss_port.so - the C restatement of the kiss-icp step
(oracle/kiss_port.c).  TEST INFRASTRUCTURE ONLY: the second CPU oracle (cross-checked bit for
bit against oracle/kiss_oracle.py in tests/test_port.py) and the CPU baseline bench.py times.
"""
import ctypes as C
import os

import numpy as np

from . import build_port

_LIB = None


class KpStats(C.Structure):
    _fields_ = [("status", C.c_int), ("n_in", C.c_int), ("n_range", C.c_int), ("n_ds", C.c_int),
                ("n_src", C.c_int), ("n_voxels", C.c_int), ("iterations", C.c_int), ("n_corr", C.c_int),
                ("dx_norm", C.c_double), ("sigma", C.c_double), ("err_dt", C.c_double),
                ("err_drot", C.c_double), ("map_points", C.c_int), ("reserved", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_V = C.c_void_p
_SIG = {
    "kp_create": (_V, [C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]),
    "kp_destroy": (None, [_V]),
    "kp_set_iteration_limits": (None, [_V, C.c_int, C.c_double]),
    "kp_set_threads": (None, [_V, C.c_int]),
    "kp_max_threads": (C.c_int, []),
    "kp_deskew_scan": (C.c_int, [_V, _V, _V, C.c_int, _V, _V, _V]),
    "kp_preprocess": (C.c_int, [_V, C.c_int, C.c_double, C.c_double, _V]),
    "kp_voxel_down_sample": (C.c_int, [_V, _V, C.c_int, C.c_double, _V, _V]),
    "kp_map_clear": (None, [_V]),
    "kp_map_num_voxels": (C.c_int, [_V]),
    "kp_map_num_points": (C.c_int, [_V]),
    "kp_map_add_points": (None, [_V, _V, C.c_int]),
    "kp_map_remove_far": (None, [_V, _V]),
    "kp_map_update": (None, [_V, _V, C.c_int, _V]),
    "kp_map_dump": (C.c_int, [_V, _V, _V, _V, C.c_int]),
    "kp_map_get_correspondences": (C.c_int, [_V, _V, C.c_int, C.c_double, _V, _V]),
    "kp_register_point_cloud": (C.c_int, [_V, _V, C.c_int, _V, C.c_double, C.c_double, _V, C.POINTER(KpStats)]),
    "kp_reset": (None, [_V]),
    "kp_num_poses": (C.c_int, [_V]),
    "kp_get_pose": (C.c_int, [_V, C.c_int, _V]),
    "kp_get_prediction_model": (None, [_V, _V]),
    "kp_register_frame": (C.c_int, [_V, _V, _V, C.c_int, _V, _V, C.POINTER(KpStats)]),
    "kp_get_points": (C.c_int, [_V, C.c_int, _V, _V, C.c_int]),
    "kp_get_trace": (C.c_int, [_V, _V, C.c_int, C.POINTER(C.c_int)]),
    "kp_det_sincos": (None, [_V, C.c_int, _V, _V]),
    "kp_se3_exp": (None, [_V, _V]),
    "kp_se3_log": (None, [_V, _V]),
}


_NATIVE = None


def _bind(path):
    lib = C.CDLL(path)
    for name, (res, args) in _SIG.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def load(build_if_missing=True):
    global _LIB
    if _LIB is not None:
        return _LIB
    if build_if_missing:
        try:
            build_port.build()
        except Exception:
            if not os.path.exists(build_port.LIB):
                raise
    _LIB = _bind(build_port.LIB)
    return _LIB


def use_native():
    """Switch every PortKissICP created from now on to the -march=native build, compiled on this host (the
    timed CPU baseline of bench.py).  Returns the flags in use; falls back to the portable build on any failure."""
    global _LIB, _NATIVE
    if _NATIVE is None:
        try:
            _NATIVE = _bind(build_port.build_native())
        except Exception:
            _NATIVE = False
    if _NATIVE:
        _LIB = _NATIVE
        return " ".join(build_port.FLAGS_NATIVE[:1] + build_port.FLAGS_NATIVE[-1:])
    load()
    return " ".join(build_port.FLAGS[:1] + build_port.FLAGS[-1:])


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else a.ctypes.data


def det_sincos(x):
    lib = load()
    x = _f64(x).reshape(-1)
    s, c = np.empty_like(x), np.empty_like(x)
    lib.kp_det_sincos(_p(x), x.size, _p(s), _p(c))
    return s, c


def se3_exp_mat(tangent):
    out = np.empty((4, 4))
    load().kp_se3_exp(_p(_f64(tangent)), _p(out))
    return out


def se3_log(T):
    out = np.empty(6)
    load().kp_se3_log(_p(_f64(T)), _p(out))
    return out


class PortKissICP:
    """KissICP + the ouster-free half of KissICPWrapper (kiss.py:83-166) on the C port; the same
    members as oracle.kiss_oracle.OracleKissICPWrapper so tests can swap them."""

    def __init__(self, *, _min_range=5, _max_range=100, voxel_size=None, max_points_per_voxel=20, deskew=True,
                 threads=1, trace_iterations=0, max_iterations=500):
        self._lib = load()
        self._max_range, self._min_range = float(_max_range), float(_min_range)
        self.voxel_size = float(_max_range) / 100.0 if voxel_size is None else float(voxel_size)
        self._h = self._lib.kp_create(self._max_range, self._min_range, self.voxel_size, max_points_per_voxel,
                                      1 if deskew else 0, 2.0, 0.1, threads, trace_iterations)
        if max_iterations != 500:
            self._lib.kp_set_iteration_limits(self._h, max_iterations, 1e-4)
        self.poses = []
        self._poses_ts, self._err_dt, self._err_drot, self._sigmas = [], [], [], []
        self.last_stats = None
        self.last_counts = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_threads(self, n):
        self._lib.kp_set_threads(self._h, int(n))

    # -- the step ------------------------------------------------------------------------
    def register_points(self, frame, timestamps, ts, initial_guess=None):
        frame, timestamps = _f64(frame), _f64(timestamps)
        g = None if initial_guess is None else _f64(initial_guess).reshape(4, 4)
        pose = np.empty((4, 4))
        st = KpStats()
        rc = self._lib.kp_register_frame(self._h, _p(frame), _p(timestamps), int(frame.shape[0]), _p(g), _p(pose),
                                         C.byref(st))
        if rc != 0 and rc != -5:
            raise RuntimeError(f"kp_register_frame failed: {rc}")
        d = st.as_dict()
        self.poses.append(pose)
        self._poses_ts.append(ts)
        self._err_dt.append(d["err_dt"])
        self._err_drot.append(d["err_drot"])
        self._sigmas.append(d["sigma"])
        self.last_stats = {k: d[k] for k in ("iterations", "n_corr", "dx_norm", "status")}
        self.last_counts = {"n": d["n_in"], "n_range": d["n_range"], "n_ds": d["n_ds"], "n_src": d["n_src"],
                            "n_vox": d["n_voxels"], "map_points": d["map_points"]}
        return pose

    @property
    def pose(self):
        return self.poses[-1] if self.poses else np.eye(4)

    def get_prediction_model(self):
        out = np.empty((4, 4))
        self._lib.kp_get_prediction_model(self._h, _p(out))
        return out

    def get_points(self, which, with_index=False):
        cap = 1 << 19
        out = np.empty((cap, 3))
        idx = np.empty(cap, dtype=np.int32)
        m = self._lib.kp_get_points(self._h, which, _p(out), _p(idx), cap)
        if m < 0:
            raise RuntimeError("kp_get_points: buffer too small")
        return (out[:m].copy(), idx[:m].copy()) if with_index else out[:m].copy()

    def get_trace(self):
        ns = C.c_int(0)
        it = self._lib.kp_get_trace(self._h, None, 0, C.byref(ns))
        out = np.empty((max(it, 1), max(ns.value, 1)), dtype=np.int32)
        self._lib.kp_get_trace(self._h, _p(out), it, C.byref(ns))
        return out[:it, :ns.value]

    # -- pieces --------------------------------------------------------------------------
    def deskew_scan(self, frame, timestamps, start_pose, finish_pose):
        frame, timestamps = _f64(frame), _f64(timestamps)
        out = np.empty_like(frame)
        self._lib.kp_deskew_scan(self._h, _p(frame), _p(timestamps), int(frame.shape[0]), _p(_f64(start_pose)),
                                 _p(_f64(finish_pose)), _p(out))
        return out

    def preprocess(self, frame, max_range=None, min_range=None):
        frame = _f64(frame)
        out = np.empty_like(frame)
        m = self._lib.kp_preprocess(_p(frame), int(frame.shape[0]),
                                    self._max_range if max_range is None else max_range,
                                    self._min_range if min_range is None else min_range, _p(out))
        return out[:m].copy()

    def voxel_down_sample(self, frame, voxel_size, return_index=False):
        frame = _f64(frame)
        n = int(frame.shape[0])
        out = np.empty((max(n, 1), 3))
        idx = np.empty(max(n, 1), dtype=np.int32)
        m = self._lib.kp_voxel_down_sample(self._h, _p(frame), n, float(voxel_size), _p(out), _p(idx))
        if m < 0:
            raise ValueError("voxel coordinate out of range")
        return (out[:m].copy(), idx[:m].copy()) if return_index else out[:m].copy()

    # -- map -----------------------------------------------------------------------------
    def map_clear(self):
        self._lib.kp_map_clear(self._h)

    def map_update(self, points, pose):
        points = _f64(points)
        self._lib.kp_map_update(self._h, _p(points), int(points.shape[0]), _p(_f64(pose)))

    def map_add_points(self, points):
        points = _f64(points)
        self._lib.kp_map_add_points(self._h, _p(points), int(points.shape[0]))

    def map_remove_far(self, origin):
        self._lib.kp_map_remove_far(self._h, _p(_f64(origin)))

    def voxel_table(self):
        V = self._lib.kp_map_num_voxels(self._h)
        cap = max(V, 1)
        keys = np.empty((cap, 3), dtype=np.int32)
        cnt = np.empty(cap, dtype=np.int32)
        pts = np.empty((cap, 20, 3))
        m = self._lib.kp_map_dump(self._h, _p(keys), _p(cnt), _p(pts), cap)
        return keys[:m], cnt[:m], pts[:m]

    def get_correspondences(self, points, max_dist):
        points = _f64(points)
        n = int(points.shape[0])
        order = np.empty(max(n, 1), dtype=np.int32)
        tgt = np.empty((max(n, 1), 3))
        self._lib.kp_map_get_correspondences(self._h, _p(points), n, float(max_dist), _p(order), _p(tgt))
        return order[:n], tgt[:n]

    def register_point_cloud(self, points, initial_guess, max_dist, kernel):
        points = _f64(points)
        pose = np.empty((4, 4))
        st = KpStats()
        self._lib.kp_register_point_cloud(self._h, _p(points), int(points.shape[0]), _p(_f64(initial_guess)),
                                          float(max_dist), float(kernel), _p(pose), C.byref(st))
        return pose, st.as_dict()
