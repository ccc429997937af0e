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


MOUNT_YPR = (0.4, 0.07, -0.05)   # fixed mounting attitude of the sensor on the platform (rad)


def beam_directions(sensor: SensorModel) -> np.ndarray:
    """(H, W, 3) float64 unit directions in the platform frame; column c looks at azimuth
    2*pi*(1 - c/W) (Ouster encoder convention) plus the per-beam azimuth offset of the four
    staggered emitter columns (Ouster beam_azimuth_angles), row 0 is the top beam; beam
    altitudes carry a small fixed calibration irregularity; the sensor is mounted with a
    slight tilt so that its axes are not aligned with the walls of the synthetic scenes."""
    H, W = sensor.H, sensor.W
    rng = np.random.default_rng(4242)
    alt = np.deg2rad(np.linspace(sensor.fov_up_deg, sensor.fov_down_deg, H) + rng.uniform(-0.1, 0.1, H))
    stag = np.deg2rad(np.tile([4.2, 1.4, -1.4, -4.2], H // 4 + 1)[:H] + rng.uniform(-0.05, 0.05, H))
    az = 2.0 * np.pi * (1.0 - np.arange(W) / W)
    A = az[None, :] + stag[:, None]
    ca, sa = np.cos(alt)[:, None], np.sin(alt)[:, None]
    d = np.empty((H, W, 3))
    d[..., 0] = ca * np.cos(A)
    d[..., 1] = ca * np.sin(A)
    d[..., 2] = sa * np.ones((1, W))
    R = _rot_zyx(np.float64(MOUNT_YPR[0]), np.float64(MOUNT_YPR[1]), np.float64(MOUNT_YPR[2]))
    return d @ R.T


@dataclass
class Scene:
    """Room (ground + 4 walls, optionally a ceiling) with axis-aligned boxes inside."""
    room_min: Tuple[float, float, float]
    room_max: Tuple[float, float, float]
    boxes: List[Tuple[Tuple[float, float, float], Tuple[float, float, float]]] = field(default_factory=list)
    closed_top: bool = False


def _clutter(rng, n, xlim, ylim, z0, a, b, lo=0.72, hi=1.32):
    """n seeded boxes inside xlim x ylim that stay clear of the corridor around the ellipse
    x = a cos, y = b sin (the LoopTrajectory path)."""
    boxes = []
    while len(boxes) < n:
        cx, cy = rng.uniform(*xlim), rng.uniform(*ylim)
        sx, sy, h = rng.uniform(0.3, 3.0), rng.uniform(0.3, 3.0), rng.uniform(0.5, 9.0)
        pts = [(cx + dx * sx / 2, cy + dy * sy / 2) for dx in (-1, 0, 1) for dy in (-1, 0, 1)]
        if any(lo < np.hypot(px / a, py / b) < hi for px, py in pts):
            continue
        boxes.append(((cx - sx / 2, cy - sy / 2, z0), (cx + sx / 2, cy + sy / 2, z0 + h)))
    return boxes


def quad_scene() -> Scene:
    """'quad' 50 x 35 m courtyard with buildings, pillars and clutter; the 30 x 20 m loop the
    platform drives stays clear of every obstacle (config 1, 2, 4)."""
    boxes = [((-3, -2, -1.5), (3, 2, 6.0)), ((-6, 3, -1.5), (-4, 5, 3.0)),
             ((21, -14, -1.5), (24, -11, 5.0)), ((-24, 13, -1.5), (-22, 16, 2.5)),
             ((21.5, 12, -1.5), (22.5, 13, 8.0)), ((-23, -15, -1.5), (-22, -14, 8.0))]
    boxes += _clutter(np.random.default_rng(7), 54, (-24.0, 24.0), (-16.5, 16.5), -1.5, 15.0, 10.0)
    return Scene((-25.0, -17.5, -1.5), (25.0, 17.5, 18.0), boxes, closed_top=True)


RELIEF_AMP = 0.04   # metres; world-fixed surface texture added along the ray


def surface_relief(x, y, z, amp=RELIEF_AMP, xp=np):
    """Smooth pseudo-random displacement as a function of the WORLD hit position: it gives the
    flat synthetic surfaces the sub-voxel structure real surfaces have (without it a
    point-to-point ICP on a sensor-fixed sampling lattice drags the estimate towards zero
    motion).  `xp` is numpy or torch."""
    return amp * (xp.sin(7.1 * x + 1.3 * xp.sin(3.3 * y)) * xp.sin(6.3 * y + 0.7 * z)
                  + 0.6 * xp.sin(13.7 * z + 2.1 * x) * xp.cos(11.9 * y - 3.0 * x))


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
    boxes = [((-4, -4, -1.5), (4, 4, 10.0)), ((-52, -30, -1.5), (-47, -20, 12.0)),
             ((-10, 34, -1.5), (0, 37, 6.0)), ((48, -36, -1.5), (53, -28, 15.0))]
    boxes += _clutter(np.random.default_rng(11), 80, (-58.0, 58.0), (-38.0, 38.0), -1.5, 30.0, 20.0)
    return Scene((-60.0, -40.0, -1.5), (60.0, 40.0, 20.0), boxes)


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
        hit = o_world + r[..., None] * d_world
        rng = np.random.default_rng([self.seed, k])
        noise = rng.normal(0.0, self.range_sigma, size=r.shape)
        rel = surface_relief(hit[..., 0], hit[..., 1], hit[..., 2])
        r = np.where(r > 0, np.maximum(r + noise + rel, 0.0), 0.0)
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


# --------------------------------------------------------------------------- torch generator
# Same scene/trajectory/sensor model evaluated with torch tensors so that a bench can build
# hundreds of scans per second on the GPU it is about to measure (data plumbing only; the
# numpy generator above stays the one the parity tests and golden fixtures use - the two agree
# except for the noise stream, which comes from torch's generator here).
def _raycast_torch(scene: Scene, o, d):
    import torch
    inf = float("inf")
    inv = 1.0 / d
    rmin = torch.as_tensor(scene.room_min, dtype=d.dtype, device=d.device)
    rmax = torch.as_tensor(scene.room_max, dtype=d.dtype, device=d.device)
    lo = (rmin - o) * inv
    hi = (rmax - o) * inv
    t_exit_axes = torch.maximum(lo, hi)
    t_exit, axis = torch.min(t_exit_axes, dim=-1)
    rng = t_exit
    if not scene.closed_top:
        open_top = (axis == 2) & (d[..., 2] > 0)
        rng = torch.where(open_top, torch.full_like(rng, inf), rng)
    for bmin, bmax in scene.boxes:
        l2 = (torch.as_tensor(bmin, dtype=d.dtype, device=d.device) - o) * inv
        h2 = (torch.as_tensor(bmax, dtype=d.dtype, device=d.device) - o) * inv
        tn = torch.max(torch.minimum(l2, h2), dim=-1).values
        tf = torch.min(torch.maximum(l2, h2), dim=-1).values
        hit = (tn < tf) & (tn > 0) & (tn < rng)
        rng = torch.where(hit, tn, rng)
    return torch.where(torch.isfinite(rng), rng, torch.zeros_like(rng))


class TorchScanGenerator:
    """SynthSequence evaluated on a torch device.  `points(k)` returns what
    KissICPWrapper.register_frame hands to the step (kiss.py:59-65) as device tensors."""

    def __init__(self, seq: SynthSequence, device):
        import torch
        self.seq = seq
        self.device = torch.device(device)
        self.dirs = torch.as_tensor(seq.dirs, dtype=torch.float64, device=self.device)
        W, H = seq.sensor.W, seq.sensor.H
        self.tnorm = torch.as_tensor(np.tile(np.linspace(0, 1.0, W, endpoint=False), (H, 1)),
                                     dtype=torch.float64, device=self.device)

    def range_image(self, k: int):
        """(range_mm (H,W) int32 tensor, ts of the last column in seconds, gt pose 4x4 ndarray)."""
        import torch
        seq, s = self.seq, self.seq.sensor
        tcol = seq.t0 + (k + np.arange(s.W) / s.W) * s.scan_period
        R, p = seq.traj.pose(tcol)
        Rt = torch.as_tensor(R, dtype=torch.float64, device=self.device)
        pt = torch.as_tensor(p, dtype=torch.float64, device=self.device)
        d_world = torch.einsum("wij,hwj->hwi", Rt, self.dirs)
        o_world = pt[None, :, :].expand_as(d_world)
        r = _raycast_torch(seq.scene, o_world, d_world)
        hit = o_world + r[..., None] * d_world
        g = torch.Generator(device=self.device)
        g.manual_seed(int(seq.seed) * 1000003 + int(k))
        noise = torch.randn(r.shape, generator=g, dtype=torch.float64, device=self.device) * seq.range_sigma
        rel = surface_relief(hit[..., 0], hit[..., 1], hit[..., 2], xp=torch)
        r = torch.where(r > 0, torch.clamp(r + noise + rel, min=0.0), torch.zeros_like(r))
        range_mm = torch.round(r * 1000.0).to(torch.int32)
        Rm, pm = seq.traj.pose(seq.t0 + (k + 0.5) * s.scan_period)
        ts_last = float(np.rint(tcol[-1] * 1e9)) * 1e-9
        return range_mm, ts_last, pose_mat(Rm, pm)

    def project(self, range_mm):
        sel = range_mm != 0
        r = range_mm.to(torch_float64()) * 0.001
        xyz = (self.dirs * r[..., None])[sel]
        return xyz.contiguous(), self.tnorm[sel].contiguous()

    def points(self, k: int):
        range_mm, ts_last, gt = self.range_image(k)
        xyz, tn = self.project(range_mm)
        return xyz, tn, ts_last, gt


def torch_float64():
    import torch
    return torch.float64
