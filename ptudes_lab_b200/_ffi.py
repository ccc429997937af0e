"""ctypes binding of libptk.so (C ABI in include/ptk.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is
present, construction fails loudly.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

PTK_OK = 0
ERRORS = {-1: "PTK_E_ARG", -2: "PTK_E_CUDA", -3: "PTK_E_CAPACITY", -4: "PTK_E_KEYRANGE",
          -5: "PTK_E_NUMERIC", -6: "PTK_E_STATE"}


class PtkConfig(C.Structure):
    _fields_ = [("max_range", C.c_double), ("min_range", C.c_double), ("voxel_size", C.c_double),
                ("max_points_per_voxel", C.c_int), ("deskew", C.c_int),
                ("initial_threshold", C.c_double), ("min_motion_th", C.c_double),
                ("max_iterations", C.c_int), ("convergence_eps", C.c_double),
                ("max_points", C.c_int), ("map_capacity", C.c_int), ("batch", C.c_int),
                ("trace_iterations", C.c_int)]


class PtkStats(C.Structure):
    _fields_ = [("status", C.c_int), ("n_in", C.c_int), ("n_range", C.c_int), ("n_ds", C.c_int),
                ("n_src", C.c_int), ("n_voxels", C.c_int), ("iterations", C.c_int), ("n_corr", C.c_int),
                ("dx_norm", C.c_double), ("sigma", C.c_double), ("err_dt", C.c_double),
                ("err_drot", C.c_double), ("map_points", C.c_int), ("icp_searches", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class PtkPacketFormat(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("profile", "pixels_per_column", "columns_per_packet", "columns_per_frame",
                                       "packet_header_size", "col_header_size", "channel_data_size", "col_footer_size",
                                       "packet_footer_size", "col_size", "lidar_packet_size", "packets_per_frame")]


class PtkScanFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("range", "range2", "reflectivity", "signal", "near_ir", "timestamp", "status",
                                          "measurement_id")]


# every symbol include/ptk.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_D = C.c_void_p      # double* that may be host or device: pass raw addresses
_I = C.c_void_p
SYMBOLS = {
    "ptk_default_config": (None, [C.POINTER(PtkConfig)]),
    "ptk_version": (C.c_int, []),
    "ptk_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(PtkConfig)]),
    "ptk_ctx_destroy": (C.c_int, [_P]),
    "ptk_reset": (C.c_int, [_P, C.c_int]),
    "ptk_last_error": (C.c_char_p, [_P]),
    "ptk_register_frame": (C.c_int, [_P, C.c_int, _D, _D, C.c_int, _D, _D, C.POINTER(PtkStats), _P]),
    "ptk_register_frame_batch": (C.c_int, [_P, C.POINTER(_D), C.POINTER(_D), C.POINTER(C.c_int), _D,
                                           C.c_char_p, _D, C.POINTER(PtkStats), _P]),
    "ptk_set_sensor": (C.c_int, [_P, C.c_int, C.c_int, _D, _D, _D, C.c_double]),
    "ptk_register_scan": (C.c_int, [_P, C.c_int, _I, _D, _D, C.POINTER(PtkStats), _P]),
    "ptk_register_scan_batch": (C.c_int, [_P, C.POINTER(_I), _D, C.c_char_p, _D, C.POINTER(PtkStats), _P]),
    "ptk_shard_config": (C.c_int, [_P, C.c_int, C.c_int]),
    "ptk_shard_begin": (C.c_int, [_P, C.c_int, _D, _D, C.c_int, _I, _D, C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "ptk_shard_search": (C.c_int, [_P, C.c_int, C.c_int, _D, _P]),
    "ptk_shard_system": (C.c_int, [_P, C.c_int, _D, C.c_int, _D, _P]),
    "ptk_shard_solve": (C.c_int, [_P, C.c_int, _D, C.c_int, C.c_int, C.POINTER(C.c_int), _P]),
    "ptk_shard_end": (C.c_int, [_P, C.c_int, _D, C.POINTER(PtkStats), _P]),
    "ptk_prefetch_scan_batch": (C.c_int, [_P, C.POINTER(_I)]),
    "ptk_num_poses": (C.c_int, [_P, C.c_int]),
    "ptk_get_pose": (C.c_int, [_P, C.c_int, C.c_int, _D]),
    "ptk_get_prediction_model": (C.c_int, [_P, C.c_int, _D]),
    "ptk_last_sigma": (C.c_double, [_P, C.c_int]),
    "ptk_get_adaptive_threshold": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "ptk_update_model_deviation": (C.c_int, [_P, C.c_int, _D]),
    "ptk_append_pose": (C.c_int, [_P, C.c_int, _D]),
    "ptk_deskew_scan": (C.c_int, [_P, _D, _D, C.c_int, _D, _D, _D, _P]),
    "ptk_preprocess": (C.c_int, [_P, _D, C.c_int, C.c_double, C.c_double, _D, C.POINTER(C.c_int), _P]),
    "ptk_voxel_down_sample": (C.c_int, [_P, _D, C.c_int, C.c_double, _D, _I, C.POINTER(C.c_int), _P]),
    "ptk_map_clear": (C.c_int, [_P, C.c_int, _P]),
    "ptk_map_empty": (C.c_int, [_P, C.c_int]),
    "ptk_map_update": (C.c_int, [_P, C.c_int, _D, C.c_int, _D, _P]),
    "ptk_map_add_points": (C.c_int, [_P, C.c_int, _D, C.c_int, _P]),
    "ptk_map_remove_far": (C.c_int, [_P, C.c_int, _D, _P]),
    "ptk_map_num_points": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ptk_map_point_cloud": (C.c_int, [_P, C.c_int, _D, C.c_int, C.POINTER(C.c_int), _P]),
    "ptk_map_dump": (C.c_int, [_P, C.c_int, _I, _I, _D, C.c_int, C.POINTER(C.c_int), _P]),
    "ptk_map_get_correspondences": (C.c_int, [_P, C.c_int, _D, C.c_int, C.c_double, _I, _D,
                                              C.POINTER(C.c_int), _P]),
    "ptk_register_point_cloud": (C.c_int, [_P, C.c_int, _D, C.c_int, _D, C.c_double, C.c_double, _D,
                                           C.POINTER(PtkStats), _P]),
    "ptk_get_points": (C.c_int, [_P, C.c_int, C.c_int, _D, _I, C.c_int, C.POINTER(C.c_int), _P]),
    "ptk_get_frame": (C.c_int, [_P, C.c_int, _D, C.c_int, C.POINTER(C.c_int), _P]),
    "ptk_get_trace": (C.c_int, [_P, C.c_int, _I, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "ptk_set_profiling": (C.c_int, [_P, C.c_int]),
    "ptk_get_profile": (C.c_int, [_P, _D, _I]),
    "ptk_kernel_name": (C.c_char_p, [C.c_int]),
    "ptk_launch_count": (C.c_longlong, [_P]),
    "ptk_get_icp_phases": (C.c_int, [_P, C.c_int, _I]),
    "ptk_control_bytes": (None, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ptk_ekf_create": (C.c_int, [C.POINTER(_P), _D, _D, _D]),
    "ptk_ekf_destroy": (C.c_int, [_P]),
    "ptk_ekf_process_imu": (C.c_int, [_P, _D, _D, C.c_double]),
    "ptk_ekf_process_imu_batch": (C.c_int, [_P, _D, _D, _D, C.c_int]),
    "ptk_ekf_process_pose": (C.c_int, [_P, _D, _D]),
    "ptk_ekf_get_nav": (C.c_int, [_P, _D, _D, _D, _D, _D, _D]),
    "ptk_ekf_get_pose": (C.c_int, [_P, _D]),
    "ptk_ekf_get_cov": (C.c_int, [_P, _D]),
    "ptk_ekf_ts": (C.c_double, [_P]),
    "ptk_shard_peer_export": (C.c_int, [_P, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_ulonglong)]),
    "ptk_shard_peer_attach": (C.c_int, [_P, C.c_int, C.c_char_p, _P]),
    "ptk_set_icp_blocks_per_lane": (C.c_int, [_P, C.c_int]),
    "ptk_fleet_replay": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(C.POINTER(_I)), C.c_int, C.POINTER(_D),
                                   C.POINTER(C.POINTER(PtkStats)), C.POINTER(_P)]),
    "ptk_packet_format_init": (C.c_int, [C.POINTER(PtkPacketFormat), C.c_int, C.c_int, C.c_int, C.c_int]),
    "ptk_packet_frame_id": (C.c_int, [C.POINTER(PtkPacketFormat), C.c_void_p]),
    "ptk_decode_packets": (C.c_int, [C.POINTER(PtkPacketFormat), C.c_int, C.c_void_p, C.c_int, C.POINTER(PtkScanFields), _P]),
    "ptk_batcher_create": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(PtkPacketFormat), C.c_int]),
    "ptk_batcher_destroy": (C.c_int, [_P]),
    "ptk_batcher_push": (C.c_int, [_P, C.c_void_p, C.POINTER(C.c_int)]),
    "ptk_batcher_flush": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "ptk_batcher_decode": (C.c_int, [_P, C.POINTER(PtkScanFields), C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "ptk_batcher_peek": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ptk_batcher_pop": (C.c_int, [_P]),
    "ptk_pcap_open": (C.c_int, [C.POINTER(_P), C.c_char_p]),
    "ptk_pcap_close": (C.c_int, [_P]),
    "ptk_pcap_next": (C.c_int, [_P, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "ptk_lz4_frame_decompress": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_void_p, C.c_ulonglong, C.POINTER(C.c_ulonglong)]),
    "ptk_ingest_last_error": (C.c_char_p, []),
    "ptk_host_alloc": (C.c_int, [C.POINTER(_P), C.c_ulonglong]),
    "ptk_host_free": (C.c_int, [_P]),
}


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load libptk.so; builds it in-tree first when the sources are newer (needs nvcc)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if build_if_missing:
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError(f"libptk.so not found at {path}; run `python -m ptudes_lab_b200.build`")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class PtkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def addr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def _check_tensor(a, dtypes, what):
    """A torch tensor is handed to the C ABI by address: refuse anything the library would misread."""
    name = str(getattr(a, "dtype", "")).replace("torch.", "")
    if name not in dtypes:
        raise TypeError(f"{what}: tensor dtype {name} not accepted (need {' or '.join(dtypes)})")
    if hasattr(a, "is_contiguous") and not a.is_contiguous():
        raise TypeError(f"{what}: tensor must be contiguous")
    return a


def f64(a, shape_last=None):
    """C-contiguous float64 ndarray view/copy of array-like `a`; torch tensors pass through by address after a
    dtype / contiguity / shape check (`shape_last`: required size of the last dimension)."""
    if hasattr(a, "data_ptr") and not isinstance(a, np.ndarray):
        _check_tensor(a, ("float64",), "f64 input")
        if shape_last is not None and (a.dim() < 1 or int(a.shape[-1]) != shape_last):
            raise TypeError(f"f64 input: last dimension must be {shape_last}, got shape {tuple(a.shape)}")
        return a
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None and (a.ndim < 1 or a.shape[-1] != shape_last):
        raise TypeError(f"f64 input: last dimension must be {shape_last}, got shape {a.shape}")
    return a


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by CUDA pinned host memory (freed when the array is collected)."""
    lib = load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = lib.ptk_host_alloc(C.byref(p), max(n, 1))
    if rc != PTK_OK:
        raise PtkError(rc, "ptk_host_alloc failed")
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                lib.ptk_host_free(self.ptr)
            except Exception:
                pass
    owner = _Owner(p.value)
    # keep the owner alive as long as any view of the buffer is
    buf._ptk_owner = owner
    return arr
