"""Seeded synthetic Ouster-shaped lidar (+IMU) data.

There is no dataset and no ouster-sdk in the build/bench environment, so the bench and
the tests run on analytically ray-cast scenes shaped like the sensor modes the reference is
used with (/root/reference/src/ptudes/utils.py:163-167: "1024x10", "2048x10").

A scan is a range image (H, W) uint32 millimetres, 0 = no return
(/root/reference/src/ptudes/kiss.py:59), columns swept over 0.1 s with the sensor moving
during the sweep, so deskew has something to undo.
"""
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np


@dataclass
class SensorModel:
    name: str
    H: int
    W: int
    fov_up_deg: float
    fov_down_deg: float
    scan_period: float = 0.1


OS0_128_1024 = SensorModel("OS0-128 1024x10", 128, 1024, 45.0, -45.0)
OS2_128_2048 = SensorModel("OS2-128 2048x10", 128, 2048, 11.25, -11.25)


def beam_directions(sensor: SensorModel) -> np.ndarray:
    """(H, W, 3) float64 unit directions in the sensor frame; column c looks at azimuth
    2*pi*(1 - c/W) (Ouster encoder convention), row 0 is the top beam."""
    alt = np.deg2rad(np.linspace(sensor.fov_up_deg, sensor.fov_down_deg, sensor.H))
    az = 2.0 * np.pi * (1.0 - np.arange(sensor.W) / sensor.W)
    ca, sa = np.cos(alt)[:, None], np.sin(alt)[:, None]
    d = np.empty((sensor.H, sensor.W, 3))
    d[..., 0] = ca * np.cos(az)[None, :]
    d[..., 1] = ca * np.sin(az)[None, :]
    d[..., 2] = sa * np.ones((1, sensor.W))
    return d


@dataclass
class Scene:
    """Room (ground + 4 walls, optionally a ceiling) with axis-aligned boxes inside."""
    room_min: Tuple[float, float, float]
    room_max: Tuple[float, float, float]
    boxes: List[Tuple[Tuple[float, float, float], Tuple[float, float, float]]] = field(default_factory=list)
    closed_top: bool = False


def quad_scene() -> Scene:
    """'quad' 50 x 35 m courtyard with a few buildings/pillars (config 1, 2, 4)."""
    boxes = [((8, 6, -1.5), (12, 9, 4.0)), ((-14, -9, -1.5), (-10, -5, 6.0)),
             ((-6, 10, -1.5), (-4, 12, 3.0)), ((15, -12, -1.5), (19, -10, 5.0)),
             ((-20, 8, -1.5), (-18, 9, 2.5)), ((2, -14, -1.5), (3, -13, 8.0)),
             ((20, 4, -1.5), (21, 5, 8.0)), ((-2, -3, -1.5), (-1, -2, 1.0))]
    return Scene((-25.0, -17.5, -1.5), (25.0, 17.5, 18.0), boxes, closed_top=True)


def street_scene() -> Scene:
    """'street' 400 x 60 m corridor with buildings on both sides (config 3)."""
    boxes = []
    rng = np.random.default_rng(7)
    x = -195.0
    while x < 190.0:
        w = float(rng.uniform(8, 25))
        h = float(rng.uniform(6, 30))
        d = float(rng.uniform(4, 12))
        boxes.append(((x, 30.0 - d, -2.0), (x + w, 30.0, h)))
        boxes.append(((x + 3.0, -30.0, -2.0), (x + w - 1.0, -30.0 + d, h * 0.8)))
        x += w + float(rng.uniform(3, 10))
    for px in np.arange(-180.0, 181.0, 24.0):
        boxes.append(((px, 6.0, -2.0), (px + 0.4, 6.4, 7.0)))
        boxes.append(((px + 11.0, -6.4, -2.0), (px + 11.4, -6.0, 7.0)))
    return Scene((-200.0, -30.0, -2.0), (200.0, 30.0, 40.0), boxes)


def hall_scene() -> Scene:
    return Scene((-60.0, -40.0, -1.5), (60.0, 40.0, 20.0),
                 [((10, 10, -1.5), (14, 14, 10.0)), ((-30, -20, -1.5), (-25, -10, 12.0)),
                  ((-10, 25, -1.5), (0, 28, 6.0)), ((35, -30, -1.5), (40, -22, 15.0))])


# --------------------------------------------------------------------------- trajectories
def _rot_zyx(yaw, pitch, roll):
    cy, sy = np.cos(yaw), np.sin(yaw)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    R = np.empty(np.shape(yaw) + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


def _soft_start(t, T=0.3):
    """Time warp with zero velocity at t=0 (the platform starts from rest, so the first two,
    un-deskewed, scans do not poison the map) and unit rate after ~1 s."""
    return t - T * (1.0 - np.exp(-t / T))


class LoopTrajectory:
    """~1 m/s around a 30 x 20 m rounded loop with +-0.2 rad/s yaw-rate wiggle (config 2)."""

    def __init__(self, seed=0, a=15.0, b=10.0, speed=1.0):
        rng = np.random.default_rng(1000 + seed)
        self.a, self.b = a, b
        self.om = speed / (0.5 * (a + b))
        self.phase = float(rng.uniform(0, 2 * np.pi)) if seed else 0.0
        self.wig = 0.1 + (0.02 * float(rng.uniform(-1, 1)) if seed else 0.0)

    def pose(self, t):
        t = _soft_start(np.asarray(t, dtype=np.float64))
        ph = self.om * t + self.phase
        p = np.stack([self.a * np.cos(ph),
                      self.b * np.sin(ph), 0.05 * np.sin(0.7 * t)], axis=-1)
        vx, vy = -self.a * np.sin(ph), self.b * np.cos(ph)
        yaw = np.arctan2(vy, vx) + self.wig * np.sin(2.0 * t)
        pitch = 0.01 * np.sin(1.3 * t)
        roll = 0.015 * np.sin(0.9 * t + 0.5)
        return _rot_zyx(yaw, pitch, roll), p


class StreetTrajectory:
    """10 m/s along x with a gentle lateral curve (config 3)."""

    def __init__(self, seed=0, speed=10.0, x0=-60.0):
        self.speed, self.x0 = speed, x0
        self.phase = 0.1 * seed

    def pose(self, t):
        t = _soft_start(np.asarray(t, dtype=np.float64))
        x = self.x0 + self.speed * t
        y = 2.0 * np.sin(0.03 * x + self.phase)
        p = np.stack([x, y, 0.03 * np.sin(1.1 * t)], axis=-1)
        yaw = np.arctan2(2.0 * 0.03 * np.cos(0.03 * x + self.phase), 1.0)
        return _rot_zyx(yaw, 0.005 * np.sin(1.7 * t), 0.005 * np.sin(1.3 * t)), p


def pose_mat(R, p):
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = p
    return T


# --------------------------------------------------------------------------- ray casting
def _raycast(scene: Scene, o, d):
    """o, d: (..., 3) world ray origins/directions -> range in metres, 0 = no return."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        lo = (np.asarray(scene.room_min) - o) * inv
        hi = (np.asarray(scene.room_max) - o) * inv
        t_exit_axes = np.maximum(lo, hi)               # exit parameter per axis (inside the room)
        t_exit = np.min(t_exit_axes, axis=-1)
        axis = np.argmin(t_exit_axes, axis=-1)
        # leaving through the open top -> no return
        open_top = (axis == 2) & (d[..., 2] > 0) & (not scene.closed_top)
        rng = np.where(open_top, np.inf, t_exit)
        for bmin, bmax in scene.boxes:
            l2 = (np.asarray(bmin) - o) * inv
            h2 = (np.asarray(bmax) - o) * inv
            tn = np.max(np.minimum(l2, h2), axis=-1)
            tf = np.min(np.maximum(l2, h2), axis=-1)
            hit = (tn < tf) & (tn > 0)
            rng = np.where(hit & (tn < rng), tn, rng)
    return np.where(np.isfinite(rng), rng, 0.0)


@dataclass
class SynthScan:
    range_mm: np.ndarray          # (H, W) uint32
    timestamp_ns: np.ndarray      # (W,) int64, per column
    gt_pose: np.ndarray           # 4x4 pose of the sensor at mid sweep
    scan_idx: int


class SynthSequence:
    """Deterministic scan generator: sensor + scene + trajectory + seed."""

    def __init__(self, sensor=OS0_128_1024, scene=None, trajectory=None, seed=0,
                 range_sigma=0.01, t0=0.0):
        self.sensor = sensor
        self.scene = scene if scene is not None else quad_scene()
        self.traj = trajectory if trajectory is not None else LoopTrajectory(seed)
        self.seed = seed
        self.range_sigma = range_sigma
        self.t0 = t0
        self.dirs = beam_directions(sensor)

    def scan(self, k: int) -> SynthScan:
        s = self.sensor
        tcol = self.t0 + (k + np.arange(s.W) / s.W) * s.scan_period
        R, p = self.traj.pose(tcol)                                    # (W,3,3), (W,3)
        d_world = np.einsum("wij,hwj->hwi", R, self.dirs)
        o_world = np.broadcast_to(p[None, :, :], d_world.shape)
        r = _raycast(self.scene, o_world, d_world)
        rng = np.random.default_rng([self.seed, k])
        noise = rng.normal(0.0, self.range_sigma, size=r.shape)
        r = np.where(r > 0, np.maximum(r + noise, 0.0), 0.0)
        range_mm = np.rint(r * 1000.0).astype(np.uint32)
        Rm, pm = self.traj.pose(self.t0 + (k + 0.5) * s.scan_period)
        ts = np.rint(tcol * 1e9).astype(np.int64)
        return SynthScan(range_mm, ts, pose_mat(Rm, pm), k)

    def points(self, k: int):
        """(xyz (N,3) f64, timestamps (N,) f64 in [0,1), ts seconds, gt pose) exactly as
        KissICPWrapper.register_frame builds them (kiss.py:59-65)."""
        sc = self.scan(k)
        xyz, tnorm = project_scan(sc.range_mm, self.dirs)
        return xyz, tnorm, float(sc.timestamp_ns[-1]) * 1e-9, sc.gt_pose


def project_scan(range_mm, dirs):
    """XYZLut stand-in + RANGE != 0 mask + per-column normalised timestamps
    (kiss.py:34-35,59-61); float64, row-major over (H, W)."""
    H, W = range_mm.shape
    sel = range_mm != 0
    r = range_mm.astype(np.float64) * 0.001
    xyz = (dirs * r[..., None])[sel]
    ts = np.tile(np.linspace(0, 1.0, W, endpoint=False), (H, 1))[sel]
    return np.ascontiguousarray(xyz), np.ascontiguousarray(ts)


def make_sequence(config: str, seed=0) -> SynthSequence:
    """Named workloads of BASELINE.json / SURVEY 8d."""
    if config in ("os0_quad", "config1", "config2", "config4"):
        return SynthSequence(OS0_128_1024, quad_scene(), LoopTrajectory(seed), seed)
    if config in ("os2_street", "config3"):
        return SynthSequence(OS2_128_2048, street_scene(), StreetTrajectory(seed), seed)
    if config == "os0_hall":
        return SynthSequence(OS0_128_1024, hall_scene(), LoopTrajectory(seed, 30.0, 20.0), seed)
    if config == "tiny":   # small sensor for fast CPU tests
        return SynthSequence(SensorModel("tiny 32x256", 32, 256, 30.0, -30.0), quad_scene(),
                             LoopTrajectory(seed), seed)
    raise ValueError(config)
