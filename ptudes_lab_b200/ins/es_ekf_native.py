"""ESEKF with the filter itself in libptk's host-native C++ (`ptk_ekf_*`, csrc/ptk_ekf.cpp): same public
surface as ptudes_lab_b200.ins.ESEKF / the reference's ptudes.ins.es_ekf.ESEKF (processImu, processPose,
nav, ts), ~50x less host time per sample - what a fleet of sequences at 100 Hz IMU needs once the lidar
step is on the GPU (the reference's author flags the Python filter as wanting a C++ core, es_ekf.py:60-62)."""
import ctypes as C
from typing import Optional

import numpy as np

from .. import _ffi
from .._ffi import addr
from .data import IMU, NavState


class ESEKFNative:
    STATE_RANK = 18

    def __init__(self, *, init_grav=None, init_bacc=None, init_bgyr=None, _logging: bool = False):
        if _logging:
            raise NotImplementedError("history logging is a feature of the Python filter (ins.ESEKF)")
        self._lib = _ffi.load()
        h = C.c_void_p()
        f = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64).reshape(3)   # noqa: E731
        g, ba, bg = f(init_grav), f(init_bacc), f(init_bgyr)
        rc = self._lib.ptk_ekf_create(C.byref(h), addr(g), addr(ba), addr(bg))
        if rc:
            raise RuntimeError(f"ptk_ekf_create failed: {rc}")
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ptk_ekf_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def processImu(self, imu: IMU) -> None:
        la = np.ascontiguousarray(imu.lacc, dtype=np.float64)
        av = np.ascontiguousarray(imu.avel, dtype=np.float64)
        prev = self.ts
        rc = self._lib.ptk_ekf_process_imu(self._h, addr(la), addr(av), float(imu.ts))
        if rc:
            raise RuntimeError(f"ptk_ekf_process_imu failed: {rc}")
        imu.dt = imu.ts - prev

    def processImuBatch(self, lacc, avel, ts) -> None:
        """n samples in one call: lacc (n,3), avel (n,3), ts (n,)."""
        la = np.ascontiguousarray(lacc, dtype=np.float64).reshape(-1, 3)
        av = np.ascontiguousarray(avel, dtype=np.float64).reshape(-1, 3)
        t = np.ascontiguousarray(ts, dtype=np.float64).reshape(-1)
        rc = self._lib.ptk_ekf_process_imu_batch(self._h, addr(la), addr(av), addr(t), int(t.shape[0]))
        if rc:
            raise RuntimeError(f"ptk_ekf_process_imu_batch failed: {rc}")

    def processPose(self, pose_corr: np.ndarray, meas_cov: Optional[np.ndarray] = None) -> None:
        p = np.ascontiguousarray(pose_corr, dtype=np.float64).reshape(4, 4)
        m = None if meas_cov is None else np.ascontiguousarray(meas_cov, dtype=np.float64).reshape(6, 6)
        rc = self._lib.ptk_ekf_process_pose(self._h, addr(p), addr(m))
        if rc:
            raise RuntimeError(f"ptk_ekf_process_pose failed: {rc}")

    @property
    def nav(self) -> NavState:
        pos, att, vel, bg, ba, gr = (np.empty(3), np.empty((3, 3)), np.empty(3), np.empty(3), np.empty(3), np.empty(3))
        self._lib.ptk_ekf_get_nav(self._h, addr(pos), addr(att), addr(vel), addr(bg), addr(ba), addr(gr))
        return NavState(pos, att, vel, bg, ba, gr)

    def pose_mat(self) -> np.ndarray:
        out = np.empty((4, 4))
        self._lib.ptk_ekf_get_pose(self._h, addr(out))
        return out

    @property
    def _cov(self) -> np.ndarray:
        out = np.empty((18, 18))
        self._lib.ptk_ekf_get_cov(self._h, addr(out))
        return out

    @property
    def ts(self) -> float:
        return float(self._lib.ptk_ekf_ts(self._h))
