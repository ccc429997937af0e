"""18-state error-state EKF: IMU mechanisation + covariance propagation, 6-D pose update.

Restates the filter of the reference (src/ptudes/ins/es_ekf.py:57-329) with the same public
surface - ESEKF(init_grav=, init_bacc=, init_bgyr=, _logging=), processImu(imu), processPose(pose,
meas_cov=None), .nav, .ts, and the log lists ekf-bench's plots read (_navs, _navs_pred, _navs_t,
_nav_update_idxs, _lg_t, _lg_acc, _lg_gyr) - and the same equations, state order and noise
constants, so that it can stand in for it under `ekf-bench` on Python 3.12.
tests/test_ekf.py checks it against vectors produced by the reference's own class
(tests/golden/make_ekf_golden.py).

State order (es_ekf.py:65-71): pos 0, vel 3, phi 6, gyro bias 9, accel bias 12, gravity 15.
"""
from typing import Optional

import numpy as np

from .data import GRAV, IMU, NavState, skew, so3_exp, so3_log

N = 18
POS, VEL, PHI, BG, BA, GR = 0, 3, 6, 9, 12, 15


def _initial_covariance() -> np.ndarray:
    """es_ekf.py:99-135: diag of squared initial sigmas; the attitude sigma is the rotation vector
    of the 'XYZ' Euler rotation by (10, 10, 10) degrees."""
    a = np.deg2rad(10.0)
    Rx = so3_exp([a, 0.0, 0.0])
    Ry = so3_exp([0.0, a, 0.0])
    Rz = so3_exp([0.0, 0.0, a])
    att_sigma = so3_log(Rx @ Ry @ Rz)      # scipy 'XYZ' (intrinsic) = Rx Ry Rz
    sig = np.concatenate([[10.0] * 3, [5.0] * 3, att_sigma, [1.5] * 3, [0.5] * 3, [2.5] * 3])
    return np.diag(sig * sig)


class ESEKF:
    STATE_RANK = N
    POS_ID, VEL_ID, PHI_ID, BG_ID, BA_ID, G_ID = POS, VEL, PHI, BG, BA, GR

    # IMU noise densities (es_ekf.py:115-118)
    ACC_BIAS_STD = 0.049
    GYR_BIAS_STD = 0.38
    ACC_VRW = 0.0043
    GYR_ARW = 0.000466

    def __init__(self, *, init_grav=None, init_bacc=None, init_bgyr=None, _logging: bool = False):
        self._logging = _logging
        self._cov = _initial_covariance()
        self._cov_init = self._cov.copy()
        self._nav_curr = NavState(
            grav=GRAV * np.array([0.0, 0.0, -1.0]) if init_grav is None else np.array(init_grav, dtype=np.float64),
            bias_acc=np.zeros(3) if init_bacc is None else np.array(init_bacc, dtype=np.float64),
            bias_gyr=np.zeros(3) if init_bgyr is None else np.array(init_bgyr, dtype=np.float64))
        self._nav_init = self._nav_curr.copy()
        self._nav_prev = self._nav_curr.copy()
        self._imu_prev = IMU()
        self._imu_curr = IMU()
        self._imu_initialized = False
        self._imu_idx = 0
        self._lg_t, self._lg_acc, self._lg_gyr = [], [], []
        self._navs, self._navs_pred, self._navs_t, self._nav_update_idxs = [], [], [], []

    @property
    def nav(self) -> NavState:
        return self._nav_curr

    @property
    def ts(self) -> float:
        return self._imu_curr.ts

    # ---------------------------------------------------------------- predict (es_ekf.py:191-257)
    def processImu(self, imu: IMU) -> None:
        self._imu_prev = self._imu_curr
        imu.dt = imu.ts - self._imu_prev.ts
        self._imu_idx += 1
        self._imu_curr = imu
        if not self._imu_initialized:           # the first sample only sets the clock
            self._imu_initialized = True
            return
        nav = self._nav_curr
        self._nav_prev = nav.copy()
        dt = imu.dt
        R_prev = nav.att_h
        f_body = imu.lacc - nav.bias_acc
        w_body = imu.avel - nav.bias_gyr
        dR = so3_exp(w_body * dt)
        # mechanisation
        a_nav = R_prev @ f_body + nav.grav
        nav.pos = nav.pos + nav.vel * dt + 0.5 * a_nav * dt * dt
        nav.vel = nav.vel + a_nav * dt
        nav.att_h = R_prev @ dR
        # error-state transition and process noise
        F = np.eye(N)
        F[POS:POS + 3, VEL:VEL + 3] = dt * np.eye(3)
        F[VEL:VEL + 3, PHI:PHI + 3] = -dt * (R_prev @ skew(f_body))
        F[VEL:VEL + 3, BA:BA + 3] = -dt * R_prev
        F[PHI:PHI + 3, PHI:PHI + 3] = dR.T
        F[PHI:PHI + 3, BG:BG + 3] = -dt * np.eye(3)
        q = np.zeros(N)
        q[VEL:VEL + 3] = (dt * self.ACC_BIAS_STD) ** 2
        q[PHI:PHI + 3] = (dt * self.GYR_BIAS_STD) ** 2
        q[BA:BA + 3] = dt * self.ACC_VRW ** 2
        q[BG:BG + 3] = dt * self.GYR_ARW ** 2
        self._cov = F @ self._cov @ F.T + np.diag(q)
        if self._logging:
            self._lg_t.append(imu.ts)
            self._lg_acc.append(imu.lacc.copy())
            self._lg_gyr.append(imu.avel.copy())
            self._navs.append(nav.copy())
            self._navs_t.append(imu.ts)
            pred = nav.copy()
            pred.cov = self._cov.copy()
            self._navs_pred.append(pred)

    # ---------------------------------------------------------------- update (es_ekf.py:259-329)
    def processPose(self, pose_corr: np.ndarray, meas_cov: Optional[np.ndarray] = None) -> None:
        nav = self._nav_curr
        if self._logging:
            pred = nav.copy()
            pred.cov = self._cov.copy()
            self._navs_pred.append(pred)
        self._nav_prev = nav.copy()
        pose_corr = np.asarray(pose_corr, dtype=np.float64)
        if meas_cov is None:                    # 2 cm / 0.01 rad (es_ekf.py:289-292)
            meas_cov = np.diag([0.02 ** 2] * 3 + [0.01 ** 2] * 3)
        H = np.zeros((6, N))
        H[0:3, POS:POS + 3] = np.eye(3)
        H[3:6, PHI:PHI + 3] = np.eye(3)
        resid = np.concatenate([pose_corr[:3, 3] - nav.pos, so3_log(nav.att_h.T @ pose_corr[:3, :3])])
        S = H @ self._cov @ H.T + meas_cov
        K = self._cov @ H.T @ np.linalg.inv(S)
        dx = K @ resid
        self._cov = (np.eye(N) - K @ H) @ self._cov
        # inject the error state
        nav.pos = nav.pos + dx[POS:POS + 3]
        nav.vel = nav.vel + dx[VEL:VEL + 3]
        nav.att_h = nav.att_h @ so3_exp(dx[PHI:PHI + 3])
        nav.bias_gyr = nav.bias_gyr + dx[BG:BG + 3]
        nav.bias_acc = nav.bias_acc + dx[BA:BA + 3]
        nav.grav = nav.grav + dx[GR:GR + 3]
        # reset: project the attitude block through G = I - [dphi/2]x
        G = np.eye(3) - skew(0.5 * dx[PHI:PHI + 3])
        self._cov[PHI:PHI + 3, PHI:PHI + 3] = G @ self._cov[PHI:PHI + 3, PHI:PHI + 3] @ G.T
        if self._logging:
            st = nav.copy()
            st.cov = self._cov.copy()
            st.update = True
            st.kiss_pose = pose_corr
            self._navs.append(st)
            self._navs_t.append(self._imu_curr.ts)
            self._nav_update_idxs.append(len(self._navs) - 1)
