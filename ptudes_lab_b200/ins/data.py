"""IMU sample, navigation state and trajectory error, with the members `ptudes ekf-bench` uses
(reference: src/ptudes/ins/data.py:10-17 GRAV/IMU, :34-93 NavState, :124-153 calc_ate,
:156-168 calc_ate_from_navs).  Written for Python >= 3.12 (the reference's dataclasses carry
ndarray defaults, which 3.11+ rejects)."""
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

GRAV = 9.782940329221166          # data.py:10


def skew(v) -> np.ndarray:
    """Cross-product matrix [v]x (the reference's utils.vee, utils.py:28-36)."""
    x, y, z = float(v[0]), float(v[1]), float(v[2])
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def so3_exp(w) -> np.ndarray:
    """Rotation matrix of the rotation vector w (Rodrigues); the reference gets it from
    ouster.sdk.pose_util.exp_rot_vec / scipy Rotation.from_rotvec."""
    w = np.asarray(w, dtype=np.float64)
    th = float(np.sqrt(w @ w))
    K = skew(w)
    if th < 1e-8:                        # series: I + K + K^2/2
        return np.eye(3) + K + 0.5 * (K @ K)
    return np.eye(3) + (np.sin(th) / th) * K + ((1.0 - np.cos(th)) / (th * th)) * (K @ K)


def so3_log(Rm) -> np.ndarray:
    """Rotation vector of a rotation matrix (ouster.sdk.pose_util.log_rot_mat in the reference)."""
    Rm = np.asarray(Rm, dtype=np.float64)
    # via the unit quaternion: stable for every angle in [0, pi]
    tr = Rm[0, 0] + Rm[1, 1] + Rm[2, 2]
    v = np.array([Rm[2, 1] - Rm[1, 2], Rm[0, 2] - Rm[2, 0], Rm[1, 0] - Rm[0, 1]])
    s = float(np.sqrt(v @ v))            # 2 sin(theta)
    c = tr - 1.0                          # 2 cos(theta)
    th = float(np.arctan2(s, c))
    if s > 1e-8:
        return v * (th / s)
    if c > 0.0:                           # theta ~ 0
        return 0.5 * v
    # theta ~ pi: axis from the largest diagonal entry of (R + I) / 2
    A = 0.5 * (Rm + np.eye(3))
    k = int(np.argmax(np.diag(A)))
    axis = A[:, k] / np.sqrt(A[k, k])
    return axis * th


@dataclass
class IMU:
    """One inertial sample: specific force (m/s^2), angular rate (rad/s), time, step (data.py:12-17)."""
    lacc: np.ndarray = field(default_factory=lambda: np.zeros(3))
    avel: np.ndarray = field(default_factory=lambda: np.zeros(3))
    ts: float = 0
    dt: float = 0


@dataclass
class NavState:
    """Position, attitude, velocity, sensor biases, gravity (data.py:34-93).  The attitude is held
    as a rotation matrix; `att_q` (xyzw) and `att_v` (rotation vector) are derived views."""
    pos: np.ndarray = field(default_factory=lambda: np.zeros(3))
    att_h: np.ndarray = field(default_factory=lambda: np.eye(3))
    vel: np.ndarray = field(default_factory=lambda: np.zeros(3))
    bias_gyr: np.ndarray = field(default_factory=lambda: np.zeros(3))
    bias_acc: np.ndarray = field(default_factory=lambda: np.zeros(3))
    grav: np.ndarray = field(default_factory=lambda: GRAV * np.array([0.0, 0.0, -1.0]))
    update: bool = False
    cov: Optional[np.ndarray] = None
    kiss_pose: Optional[np.ndarray] = None

    def pose_mat(self) -> np.ndarray:
        T = np.eye(4)
        T[:3, :3] = self.att_h
        T[:3, 3] = self.pos
        return T

    @property
    def att_v(self) -> np.ndarray:
        return so3_log(self.att_h)

    @att_v.setter
    def att_v(self, w) -> None:
        self.att_h = so3_exp(w)

    @property
    def att_q(self) -> np.ndarray:
        w = so3_log(self.att_h)
        th = float(np.sqrt(w @ w))
        if th < 1e-12:
            return np.array([0.5 * w[0], 0.5 * w[1], 0.5 * w[2], 1.0])
        return np.concatenate([np.sin(0.5 * th) * w / th, [np.cos(0.5 * th)]])

    def copy(self) -> "NavState":
        return NavState(self.pos.copy(), self.att_h.copy(), self.vel.copy(), self.bias_gyr.copy(),
                        self.bias_acc.copy(), self.grav.copy(), self.update,
                        None if self.cov is None else self.cov.copy(), self.kiss_pose)

    def __repr__(self) -> str:
        return (f"NavState:\n  pos: {self.pos}\n  vel: {self.vel}\n  att_v: {self.att_v}\n"
                f"  bg: {self.bias_gyr}\n  ba: {self.bias_acc}\n  grav: {self.grav}\n")


def calc_ate(navs_poses: Sequence[np.ndarray], gt_poses: Sequence[np.ndarray]) -> Tuple[float, float]:
    """The reference's trajectory error (data.py:124-153): the ground truth is moved so that its
    first pose coincides with the first estimated pose (no further alignment); returns
    (mean SQUARED rotation error * 180/pi, mean SQUARED translation error) - squared, as upstream."""
    assert len(navs_poses) == len(gt_poses) and len(navs_poses)
    A = np.asarray(navs_poses, dtype=np.float64)
    G = np.asarray(gt_poses, dtype=np.float64)
    align = A[0] @ np.linalg.inv(G[0])
    G = align[None] @ G
    dt = np.linalg.norm(G[:, :3, 3] - A[:, :3, 3], axis=1)
    dr = np.array([np.linalg.norm(so3_log(a[:3, :3].T @ g[:3, :3])) for a, g in zip(A, G)])
    return float(np.mean(dr * dr) * 180.0 / np.pi), float(np.mean(dt * dt))


def calc_ate_from_navs(navs, gt_poses) -> Tuple[float, float]:
    return calc_ate([n.pose_mat() for n in navs], gt_poses)
