"""The ekf-bench scan/IMU loop (ptudes_lab_b200.ekf_bench.run_ekf_ouster, reference
cli/ekf_bench.py:493-563): plumbing on the CPU with the oracle behind the KissICPWrapper surface,
and (gpu) the CUDA wrapper against the oracle wrapper under the same loop - config 4 of
BASELINE.json (ekf-bench ranges min 1 / max 70 -> v = 0.7, 100 Hz IMU)."""
import numpy as np
import pytest

from oracle import kiss_oracle as ko
from ptudes_lab_b200 import synth
from ptudes_lab_b200.ekf_bench import SynthLidarImuSource, run_ekf_ouster
from ptudes_lab_b200.ins import ESEKF, IMU, calc_ate
from ptudes_lab_b200.ouster_compat import ChanField, XYZLut, sensor_info_from_synth


class OracleWrapper:
    """The KissICPWrapper surface (kiss.py:18-166) on the NumPy oracle, LidarScan in."""

    def __init__(self, meta, _min_range=1, _max_range=70):
        self._lut = XYZLut(meta)
        w, h = meta.format.columns_per_frame, meta.format.pixels_per_column
        self._timestamps = np.tile(np.linspace(0, 1.0, w, endpoint=False), (h, 1))
        self._o = ko.OracleKissICPWrapper(_min_range=_min_range, _max_range=_max_range)
        self._kiss = self._o._kiss

    def register_frame(self, scan, initial_guess=None):
        from ptudes_lab_b200.ouster_compat import last_valid_column_ts
        sel = scan.field(ChanField.RANGE) != 0
        return self._o.register_points(self._lut(scan)[sel], self._timestamps[sel],
                                       last_valid_column_ts(scan) * 1e-9, initial_guess=initial_guess)

    @property
    def pose(self):
        return self._o.pose


def test_source_interleaves_imu_and_scans(tiny_seq):
    src = SynthLidarImuSource(tiny_seq, 3)
    items = list(src.withScanIdx())
    kinds = ["I" if isinstance(d, IMU) else "S" for _, d in items]
    assert "".join(kinds) == ("I" * 10 + "S") * 3
    ts = [d.ts for _, d in items if isinstance(d, IMU)]
    assert np.allclose(np.diff(ts), 0.01)
    # at rest the accelerometer reads +g on z (specific force), up to the mount-free body tilt
    first = SynthLidarImuSource(tiny_seq, 1, acc_noise_std=0, gyr_noise_std=0, acc_bias=(0, 0, 0), gyr_bias=(0, 0, 0)).imu_at(0.0)
    assert abs(np.linalg.norm(first.lacc) - 9.782940329221166) < 0.5 and first.lacc[2] > 9.0
    assert len(list(src.withScanIdx(start_scan=1, end_scan=1))) == 11


@pytest.mark.parametrize("use_imu", [False, True])
def test_loop_on_the_oracle(tiny_seq, use_imu):
    n = 12
    meta = sensor_info_from_synth(tiny_seq.sensor, tiny_seq.dirs)
    src = SynthLidarImuSource(tiny_seq, n)
    out = run_ekf_ouster(src, OracleWrapper(meta), ESEKF(), use_imu_prediction=use_imu)
    assert len(out["kiss_poses"]) == len(out["res_poses"]) == len(out["res_t"]) == n
    gt = src.gt_poses()
    ate_r, ate_t = calc_ate(out["kiss_poses"], gt)
    assert ate_t < 0.15 ** 2 and ate_r < 0.05          # odometry tracks the synthetic ground truth (sparse 32x256 sensor)
    ekf_r, ekf_t = calc_ate(out["res_poses"], gt)
    assert ekf_t < 0.2 ** 2                             # and the filter follows its measurements
    assert all(v is not None and v > 0 for v in out["timings"].values())


def test_scan_without_imu_in_between_is_skipped(tiny_seq):
    meta = sensor_info_from_synth(tiny_seq.sensor, tiny_seq.dirs)

    class Src(SynthLidarImuSource):
        def withScanIdx(self, **kw):
            for k, d in super().withScanIdx(**kw):
                if k == 2 and isinstance(d, IMU):
                    continue                            # drop the IMU packets of sweep 2
                yield k, d
    out = run_ekf_ouster(Src(tiny_seq, 4), OracleWrapper(meta))
    assert len(out["kiss_poses"]) == 3                  # cli/ekf_bench.py:512-518


@pytest.mark.gpu
@pytest.mark.parametrize("use_imu", [False, True])
def test_config4_cuda_wrapper_equals_oracle_under_the_loop(use_imu):
    """BASELINE config 4: OS0-128 + 100 Hz IMU, ekf-bench defaults min 1 / max 70, lidar update on the
    GPU, ESEKF on the host; with --use-imu-prediction the filter's pose is the injected guess."""
    from ptudes_lab_b200.kiss import KissICPWrapper
    seq = synth.make_sequence("os0_quad", 0)
    n = 10
    meta = sensor_info_from_synth(seq.sensor, seq.dirs)
    a = run_ekf_ouster(SynthLidarImuSource(seq, n), KissICPWrapper(meta, _min_range=1, _max_range=70, _use_extrinsics=True),
                       use_imu_prediction=use_imu)
    b = run_ekf_ouster(SynthLidarImuSource(seq, n), OracleWrapper(meta), use_imu_prediction=use_imu)
    assert np.array_equal(np.array(a["kiss_poses"]), np.array(b["kiss_poses"]))
    assert np.array_equal(np.array(a["res_poses"]), np.array(b["res_poses"]))
    ate_r, ate_t = calc_ate(a["kiss_poses"], SynthLidarImuSource(seq, n).gt_poses())
    assert ate_t < 0.1 ** 2


def test_kitti_pose_file_round_trip(tmp_path):
    from ptudes_lab_b200.ekf_bench import load_poses_kitti_format, save_poses_kitti_format
    from oracle import canon
    rng = np.random.default_rng(0)
    poses = [canon.se3_exp_mat(rng.normal(0, 1, 6)) for _ in range(5)]
    f = str(tmp_path / "poses.txt")
    save_poses_kitti_format(f, poses, header="ptk")
    back = load_poses_kitti_format(f)
    assert len(back) == 5 and all(np.array_equal(a, b) for a, b in zip(poses, back))
    assert open(f).readline().startswith("# ptk")


def test_loop_with_the_native_filter_matches_the_python_one(tiny_seq):
    """run_ekf_ouster with libptk's host-native ESEKF (ptk_ekf_*) in place of the Python filter; with
    --use-imu-prediction the filter's pose is the ICP initial guess, so the odometry depends on it too."""
    from ptudes_lab_b200.ins import ESEKFNative
    meta = sensor_info_from_synth(tiny_seq.sensor, tiny_seq.dirs)
    a = run_ekf_ouster(SynthLidarImuSource(tiny_seq, 8), OracleWrapper(meta), ESEKF(), use_imu_prediction=True)
    b = run_ekf_ouster(SynthLidarImuSource(tiny_seq, 8), OracleWrapper(meta), ESEKFNative(), use_imu_prediction=True)
    assert np.abs(np.array(a["res_poses"]) - np.array(b["res_poses"])).max() < 1e-6
    assert np.abs(np.array(a["kiss_poses"]) - np.array(b["kiss_poses"])).max() < 1e-6
    assert a["res_t"] == b["res_t"]


def test_nc_gt_pose_file_round_trip(tmp_path):
    """save_poses_nc_gt_format / read_newer_college_gt (utils.py:199-252): sec, nsec, xyz, quaternion per pose;
    the writer moves the poses from the IMU frame to BASE, the reader moves them back."""
    from ptudes_lab_b200.ekf_bench import NC_OS_IMU_TO_BASE, read_newer_college_gt, save_poses_nc_gt_format
    from oracle import canon
    rng = np.random.default_rng(3)
    poses = [canon.se3_exp_mat(rng.normal(0, 0.7, 6)) for _ in range(6)]
    ts = [1600000000.25 + 0.1 * k for k in range(6)]
    f = str(tmp_path / "nc_gt.csv")
    save_poses_nc_gt_format(f, ts, poses, header="ptk traj")
    back = read_newer_college_gt(f)
    assert len(back) == 6
    for (t, p), t0, p0 in zip(back, ts, poses):
        assert abs(t - t0) < 2e-6                           # nsec column is floor((t - sec) * 1e9) of a double
        assert np.abs(p - p0).max() < 1e-12
    raw = np.loadtxt(f, delimiter=",")
    assert raw.shape == (6, 9) and raw[0, 0] == 1600000000.0
    in_base = read_newer_college_gt(f, to_os_imu=False)
    assert np.abs(in_base[2][1] @ NC_OS_IMU_TO_BASE - poses[2]).max() < 1e-12
    head = open(f).read().splitlines()
    assert head[0] == "# ptk traj" and head[2] == "# sec,nsec,x,y,z,qx,qy,qz,qw"


def test_reduce_active_beams_and_the_beams_option(tiny_seq):
    """utils.py:328-341 / cli/ekf_bench.py:526-527: --beams N keeps N uniformly spread rows, the rest lose their
    RANGE; under the loop the odometry then sees only those beams."""
    from ptudes_lab_b200.ekf_bench import reduce_active_beams
    src = SynthLidarImuSource(tiny_seq, 1)
    ls = [d for _, d in src.withScanIdx() if not isinstance(d, IMU)][0]
    before = ls.field(ChanField.RANGE).copy()
    reduce_active_beams(ls, 4)
    after = ls.field(ChanField.RANGE)
    keep = np.linspace(0, ls.h, num=4, endpoint=False, dtype=int)
    assert np.array_equal(after[keep], before[keep])
    mask = np.ones(ls.h, bool)
    mask[keep] = False
    assert not after[mask].any()
    meta = sensor_info_from_synth(tiny_seq.sensor, tiny_seq.dirs)
    full = run_ekf_ouster(SynthLidarImuSource(tiny_seq, 4), OracleWrapper(meta))
    half = run_ekf_ouster(SynthLidarImuSource(tiny_seq, 4), OracleWrapper(meta), beams=tiny_seq.sensor.H // 2)
    assert len(half["kiss_poses"]) == 4 and not np.array_equal(np.array(full["kiss_poses"]), np.array(half["kiss_poses"]))


def test_fleet_ate_equals_the_reference_definition():
    """calc_ate_fleet: calc_ate (ins/data.py:124-153) over S trajectories in one batched call."""
    from ptudes_lab_b200.ekf_bench import calc_ate_fleet
    from oracle import canon
    rng = np.random.default_rng(5)
    S, T = 5, 12
    gt = np.array([[canon.se3_exp_mat(np.r_[0.1 * t, 0.02 * s, 0, 0, 0, 0.03 * t]) for t in range(T)] for s in range(S)])
    est = np.array([[g @ canon.se3_exp_mat(rng.normal(0, 1e-3 * (s + 1), 6)) for g in row] for s, row in enumerate(gt)])
    r, t = calc_ate_fleet(est, gt)
    for s in range(S):
        rr, tt = calc_ate(list(est[s]), list(gt[s]))
        assert abs(r[s] - rr) <= 1e-9 * max(rr, 1e-12) + 1e-15 and abs(t[s] - tt) <= 1e-12 * max(tt, 1e-12) + 1e-18


@pytest.mark.gpu
def test_beams_option_on_the_cuda_wrapper():
    """--beams through the CUDA wrapper: zeroed rows are pixels without return, poses equal the oracle's."""
    from ptudes_lab_b200.kiss import KissICPWrapper
    seq = synth.make_sequence("tiny", 0)
    meta = sensor_info_from_synth(seq.sensor, seq.dirs)
    a = run_ekf_ouster(SynthLidarImuSource(seq, 5), KissICPWrapper(meta, _min_range=1, _max_range=70), beams=seq.sensor.H // 2)
    b = run_ekf_ouster(SynthLidarImuSource(seq, 5), OracleWrapper(meta), beams=seq.sensor.H // 2)
    assert np.array_equal(np.array(a["kiss_poses"]), np.array(b["kiss_poses"]))
