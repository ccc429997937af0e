"""The lidar/IMU loop of `ptudes ekf-bench ouster` (reference: src/ptudes/cli/ekf_bench.py:493-563)
as a callable, plus a synthetic stand-in for the packet source it iterates.

Only the plumbing of the odometry step is mirrored here - what feeds KissICPWrapper.register_frame
its initial guess and what consumes its pose; the CLI, the stream statistics, the plots and the
file writers around it are out of scope.
"""
import time
from typing import Iterator, Optional, Tuple, Union

import numpy as np

from .ins import ESEKF, GRAV, IMU
from .ins.data import so3_log
from .ouster_compat import LidarScan, scan_from_synth


class SynthLidarImuSource:
    """Stand-in for ptudes.data.OusterLidarData (data.py:12-77): `withScanIdx()` yields
    (scan_idx, IMU | LidarScan) in packet order - the IMU samples of a sweep, then the finished scan.

    IMU samples (100 Hz) are the true specific force and angular rate of the synthetic trajectory
    in the body frame plus white noise and constant biases, in the manner of the reference's own
    `sim_imu` (cli/ekf_bench.py:44-79), seeded."""

    def __init__(self, seq, n_scans: int, imu_rate: float = 100.0, seed: int = 1,
                 acc_noise_std: float = 0.02, gyr_noise_std: float = 0.002,
                 acc_bias=(0.03, -0.02, 0.01), gyr_bias=(0.001, 0.003, -0.0012)):
        self.seq, self.n_scans, self.imu_rate = seq, n_scans, imu_rate
        self.rng = np.random.default_rng(seed)
        self.acc_noise_std, self.gyr_noise_std = acc_noise_std, gyr_noise_std
        self.acc_bias, self.gyr_bias = np.array(acc_bias, dtype=float), np.array(gyr_bias, dtype=float)

    def imu_at(self, t: float) -> IMU:
        """Specific force and angular rate in the body frame at time t (central differences)."""
        h = 1e-3
        tt = np.array([t - h, t, t + h]) + self.seq.t0
        R, p = self.seq.traj.pose(tt)
        acc_w = (p[2] - 2.0 * p[1] + p[0]) / (h * h)
        g_w = GRAV * np.array([0.0, 0.0, -1.0])
        lacc = R[1].T @ (acc_w - g_w)
        avel = so3_log(R[0].T @ R[2]) / (2.0 * h)
        lacc = lacc + self.rng.normal(0.0, self.acc_noise_std, 3) + self.acc_bias
        avel = avel + self.rng.normal(0.0, self.gyr_noise_std, 3) + self.gyr_bias
        return IMU(lacc, avel, float(t), 1.0 / self.imu_rate)

    def withScanIdx(self, *, start_scan: int = 0, end_scan: Optional[int] = None
                    ) -> Iterator[Tuple[int, Union[LidarScan, IMU]]]:
        period = self.seq.sensor.scan_period
        n_imu = int(round(period * self.imu_rate))
        last = self.n_scans - 1 if end_scan is None else min(end_scan, self.n_scans - 1)
        for k in range(last + 1):
            for j in range(n_imu):
                imu = self.imu_at(k * period + j / self.imu_rate)     # drawn for every sweep: seeded stream
                if k >= start_scan:
                    yield k, imu
            if k >= start_scan:
                yield k, scan_from_synth(self.seq.scan(k))

    def gt_poses(self, start_scan: int = 0):
        """Mid-sweep ground-truth poses relative to the first one (what the odometry estimates)."""
        T = [self.seq.scan(k).gt_pose for k in range(start_scan, self.n_scans)]
        T0i = np.linalg.inv(T[0])
        return [T0i @ t for t in T]


def run_ekf_ouster(data_source, kiss_icp, ekf: Optional[ESEKF] = None, *, use_imu_prediction: bool = False,
                   gt_guess=None, start_scan: int = 0, end_scan: Optional[int] = None):
    """The scan/IMU loop of ptudes_ekf_ouster (cli/ekf_bench.py:493-563).

    `kiss_icp` is anything with the KissICPWrapper surface (`register_frame(scan, initial_guess=)`,
    `.pose`, `._kiss.poses`, `._kiss.get_prediction_model()`); `gt_guess(ts) -> 4x4` stands in for the
    --use-gt-guess TrajectoryEvaluator (:536-542).  Returns the lists the CLI collects plus the
    per-stage mean timings it prints (:590-595)."""
    from .ouster_compat import last_valid_column_ts
    ekf = ekf if ekf is not None else ESEKF()
    res_t, kiss_poses, res_poses = [], [], []
    t_imu = t_corr = t_kiss = 0.0
    n_imu = n_corr = 0
    imus_per_scan = 1
    gt0 = None
    for scan_idx, d in data_source.withScanIdx(start_scan=start_scan, end_scan=end_scan):
        if isinstance(d, IMU):
            t1 = time.monotonic()
            ekf.processImu(d)                                                    # :500-503
            t_imu += time.monotonic() - t1
            n_imu += 1
            imus_per_scan += 1
            continue
        if not imus_per_scan:                                                    # :512-518
            continue
        imus_per_scan = 0
        ls = d
        ts = last_valid_column_ts(ls) * 1e-09
        if use_imu_prediction:                                                   # :533-535
            pose_guess = ekf.nav.pose_mat()
        elif gt_guess is not None:                                               # :536-542
            g = gt_guess(ts)
            if gt0 is None:
                gt0 = np.linalg.inv(g)
            pose_guess = gt0 @ g
        else:                                                                    # :543-548
            prediction = kiss_icp._kiss.get_prediction_model()
            last_pose = kiss_icp._kiss.poses[-1] if kiss_icp._kiss.poses else np.eye(4)
            pose_guess = last_pose @ prediction
        t1 = time.monotonic()
        kiss_icp.register_frame(ls, initial_guess=pose_guess)                    # :550-552
        t_kiss += time.monotonic() - t1
        t1 = time.monotonic()
        ekf.processPose(kiss_icp.pose)                                           # :554-557
        t_corr += time.monotonic() - t1
        n_corr += 1
        kiss_poses.append(kiss_icp.pose)                                         # :560-563
        res_poses.append(ekf.nav.pose_mat())
        res_t.append(ekf.ts)
    timings = {"esekf_imu_s_per_step": t_imu / n_imu if n_imu else None,
               "esekf_update_s_per_update": t_corr / n_corr if n_corr else None,
               "kiss_register_frame_s_per_frame": t_kiss / n_corr if n_corr else None}
    return {"res_t": res_t, "kiss_poses": kiss_poses, "res_poses": res_poses, "ekf": ekf, "timings": timings}


# ---- trajectory files: the KITTI pose format ekf-bench writes (utils.py:189-194, cli/ekf_bench.py:579-581)
def save_poses_kitti_format(filename: str, poses, header: str = "") -> None:
    """One line per pose: the 12 entries of the top three rows, row-major (what `ptudes flyby` and
    `ekf-bench cmp` read back)."""
    rows = np.array([np.asarray(p, dtype=np.float64)[:3, :].reshape(12) for p in poses]).reshape(-1, 12)
    np.savetxt(fname=filename, X=rows, header=header)


def load_poses_kitti_format(filename: str):
    rows = np.loadtxt(filename).reshape(-1, 12)
    out = np.tile(np.eye(4), (rows.shape[0], 1, 1))
    out[:, :3, :] = rows.reshape(-1, 3, 4)
    return list(out)
