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


def reduce_active_beams(ls, beams_num: int) -> None:
    """utils.py:328-341: keep `beams_num` uniformly spread beams (rows) of a lidar scan by zeroing the RANGE of
    all the others.  On the CUDA path a zeroed pixel is simply a pixel without return (kiss.py:59), so the row
    mask costs nothing beyond this host-side write."""
    from .ouster_compat import ChanField
    beam_idxs = np.linspace(0, ls.h, num=beams_num, endpoint=False, dtype=int)
    clean_mask = np.ones(ls.h, dtype=bool)
    clean_mask[beam_idxs] = 0
    ls.field(ChanField.RANGE)[clean_mask, :] = 0


def run_ekf_ouster(data_source, kiss_icp, ekf: Optional[ESEKF] = None, *, use_imu_prediction: bool = False,
                   gt_guess=None, start_scan: int = 0, end_scan: Optional[int] = None, beams: int = 0):
    """The scan/IMU loop of ptudes_ekf_ouster (cli/ekf_bench.py:493-563).

    `beams` is the --beams option (:364-368,526-527).  `kiss_icp` is anything with the KissICPWrapper surface (`register_frame(scan, initial_guess=)`,
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
        if beams:                                                                # :526-527
            reduce_active_beams(ls, beams)
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


# ---- the Newer College ground-truth format ekf-bench also writes and `ekf-bench cmp` reads (utils.py:199-252)
# newer_college_2021/os_imu_lidar_transforms.yaml, as utils.py:20-26 states them
NC_OS_IMU_TO_OS_SENSOR = np.eye(4)
NC_OS_IMU_TO_OS_SENSOR[:3, 3] = [-0.014, 0.012, 0.015]
NC_OS_SENSOR_TO_BASE = np.eye(4)
NC_OS_SENSOR_TO_BASE[:3, 3] = [0.001, 0.000, 0.091]
NC_OS_IMU_TO_BASE = NC_OS_SENSOR_TO_BASE @ NC_OS_IMU_TO_OS_SENSOR


def save_poses_nc_gt_format(filename: str, t, poses, header: str = "") -> None:
    """utils.py:199-228: `sec, nsec, x, y, z, qx, qy, qz, qw` per pose, poses moved from the IMU (nav) frame to
    the BASE frame on the way out (read_newer_college_gt undoes it)."""
    from scipy.spatial.transform import Rotation
    t_arr = np.asarray(t, dtype=np.float64)
    poses_arr = np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4)
    poses_arr = np.einsum("nij,jk->nik", poses_arr, np.linalg.inv(NC_OS_IMU_TO_BASE))
    res = np.zeros((len(t_arr), 9))
    res[:, 0] = np.floor(t_arr)
    res[:, 1] = np.floor((t_arr - res[:, 0]) * 1e+9)
    res[:, 2:5] = poses_arr[:, :3, 3]
    res[:, 5:9] = Rotation.from_matrix(poses_arr[:, :3, :3]).as_quat()
    if header:
        header += "\n\n" + "sec,nsec,x,y,z,qx,qy,qz,qw"
    np.savetxt(fname=filename, X=res, delimiter=", ", header=header)


def read_newer_college_gt(data_path: str, to_os_imu: bool = True):
    """utils.py:231-252: [(ts, pose4x4)] with the poses in the Ouster IMU nav frame (to_os_imu)."""
    from scipy.spatial.transform import Rotation
    gt = np.loadtxt(data_path, delimiter=",").reshape(-1, 9)
    ts = gt[:, 0] + gt[:, 1] * 1e-9
    pos = np.tile(np.eye(4), reps=(gt.shape[0], 1, 1))
    pos[:, :3, 3] = gt[:, 2:5]
    pos[:, :3, :3] = Rotation.from_quat(gt[:, 5:9]).as_matrix()
    if to_os_imu:
        pos = np.einsum("nij,jk->nik", pos, NC_OS_IMU_TO_BASE)
    return [(float(a), p) for a, p in zip(ts, pos)]


def calc_ate_fleet(est, gt, device=None):
    """calc_ate (ins/data.py:124-153) for a whole fleet at once: `est`, `gt` (S, T, 4, 4) trajectories of S
    sequences -> (ate_rot (S,), ate_trans (S,)), the reference's definition (first poses aligned, mean of the
    SQUARED errors, rotation term * 180/pi).  Runs as batched torch ops on `device` (the GPU of the rank when the
    trajectories of a fleet replay live there); float64 throughout."""
    import torch
    A = torch.as_tensor(np.asarray(est) if not torch.is_tensor(est) else est, dtype=torch.float64, device=device)
    G = torch.as_tensor(np.asarray(gt) if not torch.is_tensor(gt) else gt, dtype=torch.float64, device=device)
    assert A.shape == G.shape and A.dim() == 4 and A.shape[1] > 0
    align = A[:, 0] @ torch.linalg.inv(G[:, 0])                     # (S,4,4)
    G = align[:, None] @ G
    dt = torch.linalg.norm(G[..., :3, 3] - A[..., :3, 3], dim=-1)   # (S,T)
    R = A[..., :3, :3].transpose(-1, -2) @ G[..., :3, :3]
    cosang = ((R[..., 0, 0] + R[..., 1, 1] + R[..., 2, 2]) - 1.0) * 0.5
    # |rotvec| = angle; atan2 of the skew part keeps precision for the tiny angles an ATE is made of
    sk = torch.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], dim=-1)
    ang = torch.atan2(0.5 * torch.linalg.norm(sk, dim=-1), cosang)
    ate_r = (ang * ang).mean(dim=1) * (180.0 / np.pi)
    ate_t = (dt * dt).mean(dim=1)
    return ate_r.cpu().numpy(), ate_t.cpu().numpy()
