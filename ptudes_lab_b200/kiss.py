"""Drop-in for /root/reference/src/ptudes/kiss.py: same class, members and meaning, with the
kiss-icp calls replaced by the CUDA step behind libptk's C ABI.

Reference surface mirrored (file:line in /root/reference/src/ptudes/kiss.py):
  KissICPWrapper.__init__(metadata, *, _min_range=5, _max_range=100, _use_extrinsics=False)  :21-52
  register_frame(scan, initial_guess=None) -> pose                                          :54-74
  deskew(frame, timestamps)                                                                 :76-78
  _kiss_register_frame(frame, timestamps, ts, initial_guess=None) -> (frame, source)        :83-131
  velocity / pose / poses / poses_ts / local_map_points / _config                           :133-166
  privates the CLI reads: _kiss.poses, _kiss.get_prediction_model(), _poses_ts, _err_dt,
  _err_drot, _sigmas (cli/ekf_bench.py:545-547,652-655)
"""
from typing import List, Optional

import numpy as np

from . import odometry as _odo
from .ouster_compat import ChanField, XYZLut, last_valid_column_ts

Vec3 = np.ndarray
PoseH = np.ndarray


class _Compensator:
    """kiss_icp.deskew.MotionCompensator face (kiss.py:77,90)."""

    def __init__(self, odo):
        self._odo = odo

    def deskew_scan(self, frame, poses, timestamps):
        if len(poses) < 2:
            return frame
        return self._odo.deskew_scan(frame, timestamps, poses[-2], poses[-1])


class _AdaptiveThreshold:
    """kiss_icp AdaptiveThreshold face (kiss.py:128): the state itself lives in the library."""

    def __init__(self, odo):
        self._odo = odo

    def update_model_deviation(self, model_deviation):
        self._odo.update_model_deviation(model_deviation)


class _PoseList(list):
    """`KissICP.poses`: a list whose `append` (kiss.py:130) also reaches the library's pose list, which the
    prediction model, the deskew twist and has_moved() are computed from.  The fused step appends on the
    library side itself and records the pose here with `_record`."""

    def __init__(self, odo):
        super().__init__()
        self._odo = odo

    def append(self, pose):
        self._odo.append_pose(pose)
        super().append(pose)

    def _record(self, pose):
        super().append(pose)


class _Kiss:
    """The members of kiss_icp.kiss_icp.KissICP that ptudes reads through `wrapper._kiss` - enough of them that the
    reference's own `_kiss_register_frame` body (kiss.py:83-131) runs over this object unchanged, piece by piece
    (`tests/test_gpu_parity.py::test_reference_body_runs_piecewise`); the wrapper below uses the fused step."""

    def __init__(self, config, **odo_kw):
        self.config = config
        self._odo = _odo.Odometry(config, **odo_kw)
        self.poses: List[PoseH] = _PoseList(self._odo)
        self.compensator = _Compensator(self._odo)
        self.local_map = _odo.VoxelHashMap(self._odo, 0)
        self.adaptive_threshold = _AdaptiveThreshold(self._odo)

    def get_adaptive_threshold(self):
        return self._odo.get_adaptive_threshold()

    def get_prediction_model(self):
        # same arithmetic as the library uses for its own constant-velocity guess
        return self._odo.get_prediction_model()

    def preprocess(self, frame):
        return self._odo.preprocess(frame)

    def voxelize(self, frame):
        return self._odo.voxelize(frame)


class KissICPWrapper:
    """Thin wrapper to use with Ouster SDK LidarScans objects (CUDA step underneath)."""

    def __init__(self,
                 metadata,
                 *,
                 _min_range: float = 5,
                 _max_range: float = 100,
                 _use_extrinsics: bool = False,
                 _device: int = 0,
                 _max_points: Optional[int] = None,
                 _map_capacity: int = 262144,
                 _trace_iterations: int = 0):
        self._metadata = metadata
        self._xyz_lut = XYZLut(self._metadata, use_extrinsics=_use_extrinsics)

        w = self._metadata.format.columns_per_frame
        h = self._metadata.format.pixels_per_column

        self._timestamps = np.tile(np.linspace(0, 1.0, w, endpoint=False), (h, 1))

        self._max_range = _max_range
        self._min_range = _min_range

        self._kiss_config = _odo.load_config(None, deskew=True, max_range=self._max_range)
        self._kiss_config.data.min_range = self._min_range

        self._kiss = _Kiss(self._kiss_config, device=_device,
                           max_points=_max_points if _max_points else max(w * h, 1024),
                           map_capacity=_map_capacity, trace_iterations=_trace_iterations)

        # With a LUT that exposes its direction/offset tables the projection, the RANGE != 0 mask
        # and the timestamp gather of register_frame run on the device (ptk_register_scan).
        self._device_projection = hasattr(self._xyz_lut, "direction") and hasattr(self._xyz_lut, "range_unit")
        if self._device_projection:
            self._kiss._odo.set_sensor(self._xyz_lut.direction, getattr(self._xyz_lut, "offset", None),
                                       self._timestamps[0], self._xyz_lut.range_unit)

        # using last valid column timestamp as a pose ts
        self._poses_ts = []

        self._err_dt = []
        self._err_drot = []
        self._sigmas = []
        self._last_stats = None

    def register_frame(self, scan, initial_guess: Optional[PoseH] = None) -> PoseH:
        """Register scan with kiss icp"""
        ts = last_valid_column_ts(scan) * 1e-09

        if self._device_projection:
            new_pose, st = self._kiss._odo.register_scan(scan.field(ChanField.RANGE), initial_guess=initial_guess)
            self._record(new_pose, st)
        else:
            sel_flag = scan.field(ChanField.RANGE) != 0
            xyz = self._xyz_lut(scan)[sel_flag]
            timestamps = self._timestamps[sel_flag]
            self._kiss_register_frame(xyz, timestamps, ts, initial_guess=initial_guess, _want_frames=False)

        self._poses_ts.append(ts)

        return self.pose

    def deskew(self, frame, timestamps) -> np.ndarray:
        return self._kiss.compensator.deskew_scan(frame, self._kiss.poses, timestamps)

    def _kiss_register_frame(self, frame, timestamps, ts: float, initial_guess: Optional[PoseH] = None,
                             _want_frames: bool = True):
        """One odometry step.  Returns (preprocessed frame, source) like the reference; the two
        arrays are fetched from the device only when `_want_frames` (the reference's callers
        ignore them)."""
        odo = self._kiss._odo
        new_pose, st = odo.register_frame(frame, timestamps, initial_guess=initial_guess)
        self._record(new_pose, st)
        if not _want_frames:
            return None, None
        return odo.get_frame(), odo.get_points(1)

    def _record(self, new_pose, st):
        self._err_dt.append(st["err_dt"])           # kiss.py:122-124
        self._err_drot.append(st["err_drot"])
        self._sigmas.append(st["sigma"])
        self._kiss.poses._record(new_pose)          # kiss.py:130 (the library appended its own copy in the step)
        self._last_stats = st

    @property
    def velocity(self) -> Vec3:
        """Get linear velocity estimate from kiss icp poses"""
        if len(self.poses) < 2:
            return np.zeros(3)
        prediction = self._kiss.get_prediction_model()
        dt = self.poses_ts[-1] - self.poses_ts[-2]
        return prediction[:3, 3] / dt

    @property
    def pose(self) -> PoseH:
        """Get the last pose"""
        if not self.poses:
            return np.eye(4)
        return self.poses[-1]

    @property
    def poses(self) -> List[PoseH]:
        """Get all poses"""
        return self._kiss.poses

    @property
    def poses_ts(self) -> List[float]:
        """Get all poses"""
        return self._poses_ts

    @property
    def local_map_points(self) -> np.ndarray:
        return self._kiss.local_map.point_cloud()

    @property
    def _config(self):
        """Get underlying kiss icp config"""
        return self._kiss.config
