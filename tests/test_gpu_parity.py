"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): voxel keys, downsample selections and correspondence index sets
bit-exact; poses within 1e-6 rad / 1e-5 m.  Because oracle and kernels share one canonical
operation order (oracle/canon.py <-> csrc/ptk_canon.cuh), these tests assert EXACT equality of
float64 outputs as well, which implies the toleranced bar.
"""
import numpy as np
import pytest

from oracle import canon
from oracle import kiss_oracle as ko

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def odo():
    from ptudes_lab_b200 import odometry
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    o = odometry.Odometry(cfg, max_points=140000, map_capacity=65536, trace_iterations=64)
    yield o
    o.close()


def _rand_pose(rng, t=1.0, r=0.2):
    return canon.se3_exp_mat(np.concatenate([rng.uniform(-t, t, 3), rng.uniform(-r, r, 3)]))


def test_deskew_bit_exact(odo, os0_seq):
    xyz, ts, _, _ = os0_seq.points(2)
    rng = np.random.default_rng(0)
    a, b = _rand_pose(rng), _rand_pose(rng)
    got = odo.deskew_scan(xyz, ts, a, b)
    ref = ko.deskew_scan(xyz, ts, a, b)
    assert np.array_equal(got, ref)
    # identical poses -> identity motion
    assert np.array_equal(odo.deskew_scan(xyz, ts, a, a), ko.deskew_scan(xyz, ts, a, a))


def test_deskew_large_rotation(odo):
    rng = np.random.default_rng(1)
    xyz = rng.uniform(-50, 50, (5000, 3))
    ts = rng.uniform(0, 1, 5000)
    a = np.eye(4)
    b = canon.se3_exp_mat(np.array([3.0, -2.0, 1.0, 1.5, -2.0, 1.0]))   # |omega| ~ 2.7 rad
    assert np.array_equal(odo.deskew_scan(xyz, ts, a, b), ko.deskew_scan(xyz, ts, a, b))


def test_preprocess_and_downsample(odo, os0_seq):
    xyz, ts, _, _ = os0_seq.points(0)
    fr = odo.preprocess(xyz, 100.0, 5.0)
    assert np.array_equal(fr, ko.preprocess(xyz, 100.0, 5.0))
    for size in (0.5, 0.35, 1.0499999999999998, 1.5):
        pts, idx = odo.voxel_down_sample(fr, size, return_index=True)
        ridx = ko.voxel_down_sample_idx(fr, size)
        assert np.array_equal(idx.astype(np.int64), ridx), size
        assert np.array_equal(pts, fr[ridx])
    # idempotence: every kept point is alone in its voxel
    ds = odo.voxel_down_sample(fr, 0.5)
    assert np.array_equal(odo.voxel_down_sample(ds, 0.5), ds)


def test_downsample_edge_cases(odo):
    assert odo.voxel_down_sample(np.zeros((0, 3)), 0.5).shape == (0, 3)
    assert odo.preprocess(np.zeros((0, 3))).shape == (0, 3)
    # truncation toward zero: -0.3 and 0.3 share voxel 0
    p = np.array([[-0.3, 0, 0], [0.3, 0, 0], [-1.2, 0, 0], [1.2, 0, 0], [-0.999999, 0, 0], [0.7, 0, 0], [1.4, 0, 0]])
    pts, idx = odo.voxel_down_sample(p, 1.0, return_index=True)
    assert idx.tolist() == [0, 2, 3]
    pts, idx = odo.voxel_down_sample(p, 0.7, return_index=True)
    assert idx.tolist() == ko.voxel_down_sample_idx(p, 0.7).tolist() == [0, 2, 3, 6]
    # one point, all points identical
    assert odo.voxel_down_sample(np.ones((1, 3)), 0.5).shape == (1, 3)
    assert odo.voxel_down_sample(np.ones((1000, 3)), 0.5).shape == (1, 3)
    # everything out of range
    assert odo.preprocess(np.full((100, 3), 1000.0)).shape == (0, 3)
    # range bounds are strict on both sides
    q = np.array([[5.0, 0, 0], [100.0, 0, 0], [5.000001, 0, 0], [99.999, 0, 0]])
    assert np.array_equal(odo.preprocess(q, 100.0, 5.0), q[2:])


def _map_equal(gpu_map, ref_map):
    keys, cnt, pts = gpu_map.dump()
    rkeys, rcnt, rpts = ref_map.voxel_table()
    assert np.array_equal(keys, rkeys)
    assert np.array_equal(cnt, rcnt)
    mask = np.arange(20)[None, :] < rcnt[:, None]
    assert np.array_equal(pts[mask], rpts[mask])


def test_map_update_and_prune(odo, os0_seq):
    from ptudes_lab_b200.odometry import VoxelHashMap
    gm = VoxelHashMap(odo, 0)
    gm.clear()
    assert gm.empty()
    rm = ko.VoxelHashMap(1.0, 100.0, 20)
    rng = np.random.default_rng(3)
    pose = np.eye(4)
    for k in range(4):
        xyz, ts, _, _ = os0_seq.points(k)
        ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)
        gm.update(ds, pose)
        rm.update(ds, pose)
        _map_equal(gm, rm)
        pose = canon.rigid_mul(pose, _rand_pose(rng, 0.5, 0.05))
    assert not gm.empty()
    # a far-away origin prunes nearly everything; boundary is strict '>'
    origin = np.array([80.0, 0.0, 0.0])
    gm.remove_far_away_points(origin)
    rm.remove_far_away_points(origin)
    _map_equal(gm, rm)
    pc = gm.point_cloud()
    rpc = rm.point_cloud()
    assert pc.shape == rpc.shape
    assert np.array_equal(pc[np.lexsort(pc.T[::-1])], rpc[np.lexsort(rpc.T[::-1])])


def test_map_cap_and_order(odo):
    """25 points into one voxel keep the first 20 in order; later scans append after them."""
    from ptudes_lab_b200.odometry import VoxelHashMap
    gm = VoxelHashMap(odo, 0)
    gm.clear()
    rm = ko.VoxelHashMap(1.0, 100.0, 20)
    rng = np.random.default_rng(5)
    a = rng.uniform(0.05, 0.95, (12, 3)) + np.array([3.0, 4.0, 5.0])
    b = rng.uniform(0.05, 0.95, (25, 3)) + np.array([3.0, 4.0, 5.0])
    for cloud in (a, b):
        gm.add_points(cloud)
        rm.add_points(cloud)
        _map_equal(gm, rm)
    keys, cnt, pts = gm.dump()
    assert cnt.tolist() == [20]
    assert np.array_equal(pts[0, :12], a) and np.array_equal(pts[0, 12:20], b[:8])
    # prune: a voxel whose first point is at distance exactly max_range is kept
    gm.clear()
    gm.add_points(np.array([[100.0, 0.0, 0.0], [100.00000001, 0.5, 1.5]]))
    gm.remove_far_away_points(np.zeros(3))
    assert gm.counts() == (1, 1)


def test_correspondences_exact(odo, os0_seq):
    from ptudes_lab_b200.odometry import VoxelHashMap
    gm = VoxelHashMap(odo, 0)
    gm.clear()
    rm = ko.VoxelHashMap(1.0, 100.0, 20)
    for k in range(3):
        xyz, ts, _, _ = os0_seq.points(k)
        ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)
        gm.update(ds, np.eye(4))
        rm.update(ds, np.eye(4))
    xyz, ts, _, _ = os0_seq.points(3)
    q = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 1.5)
    q = np.concatenate([q, np.array([[500.0, 500.0, 500.0]])])      # no neighbour at all (B.3)
    for max_dist in (6.0, 0.3):
        acc, tgt, order = gm.get_correspondences(q, max_dist, return_index=True)
        racc, rtgt, rorder = rm.get_correspondences(q, max_dist, return_index=True)
        assert np.array_equal(acc, racc)
        assert np.array_equal(order[acc], rorder[racc])
        assert np.array_equal(tgt[acc], rtgt[racc])
    assert not acc[-1]


def test_register_point_cloud_bit_exact(odo, os0_seq):
    from ptudes_lab_b200.odometry import VoxelHashMap, register_frame
    gm = VoxelHashMap(odo, 0)
    gm.clear()
    rm = ko.VoxelHashMap(1.0, 100.0, 20)
    xyz, ts, _, _ = os0_seq.points(0)
    ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)
    gm.update(ds, np.eye(4))
    rm.update(ds, np.eye(4))
    xyz, ts, _, _ = os0_seq.points(4)
    src = ko.voxel_down_sample(ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5), 1.5)
    guess = canon.se3_exp_mat(np.array([0.1, -0.05, 0.02, 0.01, -0.01, 0.02]))
    trace = []
    rpose, rst = ko.register_point_cloud(src, rm, guess, 6.0, 2.0 / 3.0, trace=trace)
    pose, st = register_frame(src, gm, guess, 6.0, 2.0 / 3.0, return_stats=True)
    assert st["iterations"] == rst["iterations"]
    assert st["n_corr"] == rst["n_corr"]
    assert np.array_equal(pose, rpose)
    tr = odo.get_trace()
    assert tr.shape[0] == min(len(trace), 64)
    for it in range(tr.shape[0]):
        assert np.array_equal(tr[it].astype(np.int64), trace[it]["order"]), it
    # known answer: a cloud against a moved copy of itself recovers the motion
    T = canon.se3_exp_mat(np.array([0.05, -0.03, 0.02, 0.004, -0.003, 0.01]))
    s2 = ko.voxel_down_sample(ds, 1.5)
    x, y, z = canon.transform_points(canon.rigid_inv(T), s2[:, 0], s2[:, 1], s2[:, 2])
    pose = register_frame(np.stack([x, y, z], 1), gm, np.eye(4), 6.0, 2.0 / 3.0)
    assert np.abs(pose - T).max() < 1e-9
    # empty map -> the guess comes back; no correspondences -> guess, status 1 (B.5)
    # (the guess crosses the boundary as a Sophus::SE3d upstream, so it comes back re-orthonormalised)
    gm.clear()
    rempty = ko.VoxelHashMap(1.0, 100.0, 20)
    pose = register_frame(src, gm, guess, 6.0, 0.6)
    assert np.array_equal(pose, ko.register_point_cloud(src, rempty, guess, 6.0, 0.6)[0])
    assert np.abs(pose - guess).max() < 1e-15
    far = np.array([[900.0, 900.0, 900.0]])
    gm.add_points(far)
    rempty.add_points(far)
    pose, st = register_frame(src, gm, guess, 6.0, 0.6, return_stats=True)
    assert st["status"] == 1 and st["n_corr"] == 0
    assert np.array_equal(pose, ko.register_point_cloud(src, rempty, guess, 6.0, 0.6)[0])


def _run_sequence(seq, n_scans, min_range, max_range, guesses=None, max_points=140000):
    from ptudes_lab_b200 import odometry
    cfg = odometry.load_config(None, deskew=True, max_range=max_range)
    cfg.data.min_range = min_range
    o = odometry.Odometry(cfg, max_points=max_points, map_capacity=131072, trace_iterations=8)
    ref = ko.OracleKissICPWrapper(_min_range=min_range, _max_range=max_range)
    try:
        for k in range(n_scans):
            xyz, ts, tsec, _ = seq.points(k)
            g = None if guesses is None else guesses(k, ref)
            trace = []
            ref.register_points(xyz, ts, tsec, initial_guess=g, trace=trace)
            pose, st = o.register_frame(xyz, ts, initial_guess=g)
            c = ref.last_counts
            assert (st["n_in"], st["n_range"], st["n_ds"], st["n_src"]) == (c["n"], c["n_range"], c["n_ds"], c["n_src"]), k
            # downsample selections (indices into the input scan) are bit-exact
            ds, ds_idx = o.get_points(0, with_index=True)
            s0, s_idx = o.get_points(1, with_index=True)
            fr = ref._kiss.compensator.deskew_scan(xyz, ref.poses[:-1], ts) if k >= 2 else xyz
            mask = ko.range_mask(fr, max_range, min_range)
            rid1 = np.flatnonzero(mask)[ko.voxel_down_sample_idx(fr[mask], cfg.mapping.voxel_size * 0.5)]
            assert np.array_equal(ds_idx.astype(np.int64), rid1), k
            assert np.array_equal(ds, fr[rid1]), k
            rid2 = ko.voxel_down_sample_idx(fr[rid1], cfg.mapping.voxel_size * 1.5)
            assert np.array_equal(s_idx.astype(np.int64), rid2), k
            assert np.array_equal(s0, fr[rid1][rid2]), k
            # correspondences per iteration, iteration count, sigma, pose: exact
            assert st["iterations"] == ref.last_stats["iterations"], k
            tr = o.get_trace()
            for it in range(tr.shape[0]):
                assert np.array_equal(tr[it].astype(np.int64), trace[it]["order"]), (k, it)
            assert st["sigma"] == ref._sigmas[-1], k
            assert np.array_equal(pose, ref.pose), k
            assert st["err_dt"] == ref._err_dt[-1]
            assert abs(st["err_drot"] - ref._err_drot[-1]) == 0.0
            assert st["n_voxels"] == c["n_vox"], k
        _map_equal(odometry.VoxelHashMap(o, 0), ref._kiss.local_map)
    finally:
        o.close()
    return ref


def test_sequence_os0_default_ranges(os0_seq):
    """config 2 shape (OS0-128 1024x10, v = 1.0), first 8 scans, constant-velocity guess."""
    _run_sequence(os0_seq, 8, 5.0, 100.0)


def test_sequence_ekf_bench_ranges(os0_seq):
    """config 4 ranges (min 1, max 70 -> v = 0.7, grids 0.35 / 1.0499999999999998)."""
    _run_sequence(os0_seq, 5, 1.0, 70.0)


def test_sequence_with_injected_guess(tiny_seq):
    """initial_guess injection (cli/ekf_bench.py:533-551): a perturbed constant-velocity guess."""
    rng = np.random.default_rng(11)

    def guesses(k, ref):
        if k < 2:
            return None
        g = canon.rigid_mul(ref.pose, ref._kiss.get_prediction_model())
        return canon.rigid_mul(g, canon.se3_exp_mat(rng.normal(0, 0.01, 6)))
    _run_sequence(tiny_seq, 12, 5.0, 100.0, guesses=guesses, max_points=16384)


def test_os2_long_range():
    """config 3 shape: OS2-128 2048x10, max_range 200 -> v = 2.0."""
    from ptudes_lab_b200 import synth
    _run_sequence(synth.make_sequence("os2_street", 0), 4, 5.0, 200.0, max_points=262144)


def test_empty_and_degenerate_scans(tiny_seq):
    from ptudes_lab_b200 import odometry
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    o = odometry.Odometry(cfg, max_points=16384, map_capacity=16384)
    ref = ko.OracleKissICPWrapper()
    try:
        xyz, ts, tsec, _ = tiny_seq.points(0)
        for frame, t in ((np.zeros((0, 3)), np.zeros(0)), (xyz, ts), (np.zeros((0, 3)), np.zeros(0)),
                         (np.full((50, 3), 500.0), np.full(50, 0.5)), (xyz[:100], ts[:100]), (xyz, ts)):
            ref.register_points(frame, t, tsec)
            pose, st = o.register_frame(frame, t)
            assert np.array_equal(pose, ref.pose)
            assert st["n_src"] == ref.last_counts["n_src"]
        with pytest.raises(Exception):
            o.register_frame(np.zeros((20000, 3)), np.zeros(20000))     # > max_points
    finally:
        o.close()


def test_wrapper_drop_in(tiny_seq):
    """KissICPWrapper (reference surface, kiss.py:18-166) against the oracle wrapper."""
    from ptudes_lab_b200 import synth
    from ptudes_lab_b200.kiss import KissICPWrapper
    from ptudes_lab_b200.ouster_compat import scan_from_synth, sensor_info_from_synth
    meta = sensor_info_from_synth(tiny_seq.sensor, tiny_seq.dirs)
    w = KissICPWrapper(meta, _min_range=1, _max_range=70)
    ref = ko.OracleKissICPWrapper(_min_range=1, _max_range=70)
    assert np.array_equal(w.pose, np.eye(4)) and np.array_equal(w.velocity, np.zeros(3))
    for k in range(6):
        sc = tiny_seq.scan(k)
        xyz, ts, tsec, _ = tiny_seq.points(k)
        ref.register_points(xyz, ts, tsec)
        pose = w.register_frame(scan_from_synth(sc))
        assert np.array_equal(pose, ref.pose)
    assert w._sigmas == ref._sigmas and w._err_dt == ref._err_dt
    assert np.allclose(w._err_drot, ref._err_drot, atol=0, rtol=0)
    assert w.poses_ts == ref.poses_ts
    assert np.array_equal(w.velocity, ref.velocity)
    assert np.array_equal(w._kiss.get_prediction_model(), ref._kiss.get_prediction_model())
    a, b = w.local_map_points, ref.local_map_points
    assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
    assert w._config.mapping.voxel_size == 0.7
    # public deskew() and the (frame, source) return of _kiss_register_frame
    xyz, ts, tsec, _ = tiny_seq.points(6)
    assert np.array_equal(w.deskew(xyz, ts), ref.deskew(xyz, ts))
    rf, rs = ref._kiss_register_frame(xyz, ts, tsec)
    f, s = w._kiss_register_frame(xyz, ts, tsec)
    assert np.array_equal(f, rf) and np.array_equal(s, rs)


def test_reference_body_runs_piecewise(tiny_seq):
    """The reference's own `_kiss_register_frame` (kiss.py:83-131) calls kiss-icp piece by piece: compensator.deskew_scan,
    preprocess, voxelize, get_adaptive_threshold, get_prediction_model, registration.register_frame,
    adaptive_threshold.update_model_deviation, local_map.update, poses.append.  The same sequence over this
    package's kiss-icp-shaped objects (`kiss._Kiss`, `odometry.register_frame`) gives the fused step's poses; the
    guess and the model deviation are NumPy products here as in the reference, hence a tolerance instead of `==`."""
    from ptudes_lab_b200 import odometry
    from ptudes_lab_b200.kiss import _Kiss
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    cfg.data.min_range = 5.0
    k = _Kiss(cfg, device=0, max_points=16384, map_capacity=16384)
    ref = ko.OracleKissICPWrapper()
    try:
        for s in range(6):
            xyz, ts, tsec, _ = tiny_seq.points(s)
            ref.register_points(xyz, ts, tsec)
            frame = k.compensator.deskew_scan(xyz, k.poses, ts)                      # kiss.py:90
            frame = k.preprocess(frame)                                              # :93
            source, frame_ds = k.voxelize(frame)                                     # :96
            sigma = k.get_adaptive_threshold()                                       # :99
            guess = (k.poses[-1] if k.poses else np.eye(4)) @ k.get_prediction_model()   # :102-105
            pose = odometry.register_frame(points=source, voxel_map=k.local_map, initial_guess=guess,
                                           max_correspondance_distance=3 * sigma, kernel=sigma / 3)   # :108-114
            k.adaptive_threshold.update_model_deviation(np.linalg.inv(guess) @ pose)     # :128
            k.local_map.update(frame_ds, pose)                                       # :129
            k.poses.append(pose)                                                     # :130
            assert abs(sigma - ref._sigmas[-1]) < 1e-9, s
            assert np.allclose(pose, ref.pose, rtol=0, atol=1e-8), s
            assert len(source) == ref.last_counts["n_src"] and len(frame_ds) == ref.last_counts["n_ds"]
        assert k._odo.num_poses() == 6 and len(k.poses) == 6
    finally:
        k._odo.close()


def test_register_scan_range_image_path(tiny_seq):
    """ptk_register_scan (projection + mask + column timestamps on the device, kiss.py:59-61)
    against the oracle fed with the host-projected cloud, and against the xyz entry point."""
    from ptudes_lab_b200 import odometry
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    a = odometry.Odometry(cfg, max_points=16384, map_capacity=32768, trace_iterations=4)
    b = odometry.Odometry(cfg, max_points=16384, map_capacity=32768, trace_iterations=4)
    ref = ko.OracleKissICPWrapper()
    try:
        a.set_sensor(tiny_seq.dirs)
        with pytest.raises(Exception):
            b.register_scan(tiny_seq.scan(0).range_mm)          # no sensor set on b
        for k in range(7):
            sc = tiny_seq.scan(k)
            rng = sc.range_mm.copy()
            if k == 4:
                rng[:] = 0                                       # a scan without a single return
            if k == 5:
                rng[::2, ::3] = 0                                # ragged: many dropped returns
            from ptudes_lab_b200.synth import project_scan
            xyz, ts = project_scan(rng, tiny_seq.dirs)
            ref.register_points(xyz, ts, 0.1 * (k + 1))
            pa, sa = a.register_scan(rng)
            pb, sb = b.register_frame(xyz, ts)
            assert np.array_equal(pa, ref.pose) and np.array_equal(pb, ref.pose), k
            for key in ("n_in", "n_range", "n_ds", "n_src", "n_voxels", "iterations", "n_corr", "sigma", "err_dt"):
                assert sa[key] == sb[key], (k, key)
            assert sa["n_in"] == xyz.shape[0]
            pix = np.flatnonzero(rng.reshape(-1) != 0)
            da, ia = a.get_points(0, with_index=True)
            db, ib = b.get_points(0, with_index=True)
            assert np.array_equal(da, db) and np.array_equal(ia, pix[ib]), k
            assert np.array_equal(a.get_points(1), b.get_points(1))
            assert np.array_equal(a.get_trace(), b.get_trace())
            assert np.array_equal(a.get_frame(), b.get_frame())
        _map_equal(odometry.VoxelHashMap(a, 0), ref._kiss.local_map)
    finally:
        a.close()
        b.close()


def test_wrapper_with_extrinsics_and_offsets(tiny_seq):
    """_use_extrinsics=True (cli/ekf_bench.py:451-454): the LUT carries direction AND offset."""
    from ptudes_lab_b200.kiss import KissICPWrapper
    from ptudes_lab_b200.ouster_compat import ChanField, scan_from_synth, sensor_info_from_synth
    meta = sensor_info_from_synth(tiny_seq.sensor, tiny_seq.dirs)
    meta.extrinsic = canon.se3_exp_mat(np.array([0.05, -0.02, 0.1, 0.02, -0.03, 0.4]))
    w = KissICPWrapper(meta, _min_range=1, _max_range=70, _use_extrinsics=True)
    assert w._device_projection and w._xyz_lut.offset is not None
    ref = ko.OracleKissICPWrapper(_min_range=1, _max_range=70)
    for k in range(5):
        scan = scan_from_synth(tiny_seq.scan(k))
        sel = scan.field(ChanField.RANGE) != 0                   # kiss.py:59-61 on the host for the oracle
        ref.register_points(w._xyz_lut(scan)[sel], w._timestamps[sel], 0.1 * (k + 1))
        assert np.array_equal(w.register_frame(scan), ref.pose), k
    assert w._sigmas == ref._sigmas and w._err_dt == ref._err_dt
    a, b = w.local_map_points, ref.local_map_points
    assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])


def test_register_scan_batch_matches_single(tiny_seq):
    from ptudes_lab_b200 import odometry, synth
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    B = 3
    seqs = [synth.make_sequence("tiny", s) for s in range(B)]
    ob = odometry.Odometry(cfg, max_points=16384, map_capacity=32768, batch=B)
    singles = [odometry.Odometry(cfg, max_points=16384, map_capacity=32768) for _ in range(B)]
    try:
        ob.set_sensor(seqs[0].dirs)
        for o in singles:
            o.set_sensor(seqs[0].dirs)
        for k in range(5):
            rngs = [s.scan(k).range_mm for s in seqs]
            poses, stats = ob.register_scan_batch(rngs)
            for l in range(B):
                p1, s1 = singles[l].register_scan(rngs[l])
                assert np.array_equal(poses[l], p1), (k, l)
                assert stats[l]["iterations"] == s1["iterations"] and stats[l]["n_src"] == s1["n_src"]
    finally:
        ob.close()
        for o in singles:
            o.close()


def test_prefetched_host_scans_give_the_same_poses(tiny_seq):
    """ptk_prefetch_scan_batch: the next step's H2D copy overlaps the running step; results unchanged,
    also when a prefetched image is not the one registered next (falls back to the in-line copy)."""
    from ptudes_lab_b200 import _ffi, odometry, synth
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    B = 2
    seqs = [synth.make_sequence("tiny", s) for s in range(B)]
    a = odometry.Odometry(cfg, max_points=16384, map_capacity=32768, batch=B)
    b = odometry.Odometry(cfg, max_points=16384, map_capacity=32768, batch=B)
    try:
        for o in (a, b):
            o.set_sensor(seqs[0].dirs)
        scans = []
        for k in range(6):
            row = []
            for s in seqs:
                h = _ffi.pinned_empty((s.sensor.H, s.sensor.W), dtype=np.uint32)
                h[...] = s.scan(k).range_mm
                row.append(h)
            scans.append(row)
        for k in range(6):
            if k + 1 < 6 and k != 2:
                a.prefetch_scan_batch(scans[k + 1])
            if k == 2:
                a.prefetch_scan_batch([scans[5][0], None])          # a wrong guess for lane 0, nothing for lane 1
            pa, _ = a.register_scan_batch(scans[k])
            pb, _ = b.register_scan_batch([torch_u32(x) for x in scans[k]])
            assert np.array_equal(pa, pb), k
    finally:
        a.close()
        b.close()


def torch_u32(host):
    import torch
    return torch.as_tensor(host.astype(np.int64), device="cuda").to(torch.int32).contiguous()


def test_100_scan_trajectory_equals_the_cpu_port():
    """north_star: per-frame poses and the trajectory ATE over 100 scans.  100 scans of the OS0-128 1024x10
    sequence through the range-image entry against the C port of the oracle: every pose bit-identical,
    hence ATE (reference definition, ins/data.py:124-153) of exactly zero between the two, and both
    close to the synthetic ground truth."""
    import torch
    from oracle import port
    from ptudes_lab_b200 import odometry, synth
    from ptudes_lab_b200.ins import calc_ate
    seq = synth.make_sequence("os0_quad", 0)
    gen = synth.TorchScanGenerator(seq, torch.device("cuda", 0))
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    o = odometry.Odometry(cfg, max_points=131072, map_capacity=65536)
    o.set_sensor(seq.dirs)
    ref = port.PortKissICP(threads=8)
    gpu, gts = [], []
    try:
        for k in range(100):
            rng, _, gt = gen.range_image(k)
            pose, st = o.register_scan(rng)
            xyz, ts = synth.project_scan(rng.cpu().numpy().astype(np.uint32), seq.dirs)
            ref.register_points(xyz, ts, 0.1 * (k + 1))
            assert np.array_equal(pose, ref.pose), k
            assert st["iterations"] == ref.last_stats["iterations"], k
            gpu.append(pose)
            gts.append(gt)
    finally:
        o.close()
    assert calc_ate(gpu, ref.poses) == (0.0, 0.0)
    g0 = np.linalg.inv(gts[0])
    ate_r, ate_t = calc_ate(gpu, [g0 @ g for g in gts])
    assert ate_t < 0.1 ** 2, (ate_r, ate_t)
    ref.close()


def test_os2_long_run_with_prune_equals_the_cpu_port():
    """config 3 as the survey meant it: OS2-128 2048x10, max_range 200 m (voxel 2 m), a platform that travels 240 m
    (30 m/s, 80 scans) so that the local map grows large and the 200 m prune erases voxels scan after scan.  Poses,
    iteration counts and the map's size equal the C port of the oracle at every scan."""
    import torch
    from oracle import port
    from ptudes_lab_b200 import odometry, synth
    seq = synth.SynthSequence(synth.OS2_128_2048, synth.street_scene(), synth.StreetTrajectory(0, speed=30.0, x0=-150.0), 0)
    gen = synth.TorchScanGenerator(seq, torch.device("cuda", 0))
    cfg = odometry.load_config(None, deskew=True, max_range=200.0)
    o = odometry.Odometry(cfg, max_points=262144, map_capacity=65536)
    o.set_sensor(seq.dirs)
    ref = port.PortKissICP(_max_range=200.0, threads=8)
    vox = []
    try:
        for k in range(80):
            rng, _, _ = gen.range_image(k)
            pose, st = o.register_scan(rng)
            xyz, ts = synth.project_scan(rng.cpu().numpy().astype(np.uint32), seq.dirs)
            ref.register_points(xyz, ts, 0.1 * (k + 1))
            assert np.array_equal(pose, ref.pose), k
            assert st["iterations"] == ref.last_stats["iterations"], k
            assert (st["n_voxels"], st["map_points"]) == (ref.last_counts["n_vox"], ref.last_counts["map_points"]), k
            vox.append(st["n_voxels"])
        # every scan adds voxels ahead; the map can only shrink between two scans if the prune erased more than that
        assert any(b < a for a, b in zip(vox, vox[1:])), vox
        assert np.linalg.norm(pose[:3, 3]) > 200.0
    finally:
        o.close()
        ref.close()


def test_wide_batch_uses_one_block_per_lane_and_the_global_cache(tiny_seq):
    """300 lanes > the co-resident block budget: the ICP launch is split into chunks with ONE block per
    lane, whose 34 groups (> 1024 points) no longer fit the shared-memory cache, so the global-memory
    cache arrays and the multi-chunk walk are exercised.  Every lane gets the same scans and must
    produce the oracle's poses."""
    from ptudes_lab_b200 import odometry
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    B = 300
    o = odometry.Odometry(cfg, max_points=8192, map_capacity=4096, batch=B, trace_iterations=0)
    ref = ko.OracleKissICPWrapper()
    try:
        o.set_sensor(tiny_seq.dirs)
        for k in range(4):
            sc = tiny_seq.scan(k)
            xyz, ts, tsec, _ = tiny_seq.points(k)
            ref.register_points(xyz, ts, tsec)
            poses, stats = o.register_scan_batch([sc.range_mm] * B)
            assert ref.last_counts["n_src"] > 1024
            for l in (0, 1, 147, 295, 296, 299):
                assert np.array_equal(poses[l], ref.pose), (k, l)
                assert stats[l]["iterations"] == ref.last_stats["iterations"]
            assert all(np.array_equal(poses[l], poses[0]) for l in range(B))
    finally:
        o.close()


def test_error_paths_and_reset(tiny_seq):
    """Errors come back as PtkError with the library's code; a reset lane starts a fresh sequence
    (error behaviour of the reference: exceptions propagate out of register_frame, kiss.py:54-74)."""
    from ptudes_lab_b200 import odometry
    from ptudes_lab_b200._ffi import PtkError
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    xyz, ts, tsec, _ = tiny_seq.points(0)
    # the local map does not fit: PTK_E_CAPACITY (-3)
    small = odometry.Odometry(cfg, max_points=16384, map_capacity=64)
    try:
        with pytest.raises(PtkError) as e:
            small.register_frame(xyz, ts)
        assert e.value.code == -3
        # the failed step committed nothing, and the lane refuses further steps (its map is incomplete) ...
        assert small.num_poses(0) == 0
        with pytest.raises(PtkError) as e:
            small.register_frame(xyz, ts)
        assert e.value.code == -6
        # ... until it is reset; a scan that fits then runs normally (no stale table entry, no hang)
        small.reset(0)
        few = xyz[:40]
        pose, st = small.register_frame(few, ts[:40])
        assert np.array_equal(pose, np.eye(4)) and small.num_poses(0) == 1
        pose, st = small.register_frame(few, ts[:40])
        assert st["n_voxels"] > 0 and small.num_poses(0) == 2
    finally:
        small.close()
    o = odometry.Odometry(cfg, max_points=16384, map_capacity=16384, batch=2)
    try:
        # voxel coordinate beyond +-2^20: PTK_E_KEYRANGE (-4); the oracle raises as well
        far = np.array([[3.0e6, 0.0, 0.0], [1.0, 2.0, 3.0]])
        with pytest.raises(PtkError) as e:
            o.voxel_down_sample(far, 1.0)
        assert e.value.code == -4
        with pytest.raises(ValueError):
            ko.voxel_down_sample(far, 1.0)
        # bad arguments: PTK_E_ARG (-1)
        with pytest.raises(PtkError) as e:
            o.register_frame(xyz, ts, lane=5)
        assert e.value.code == -1
        with pytest.raises(PtkError):
            o.voxel_down_sample(xyz, -1.0)
        with pytest.raises(PtkError):
            o.get_pose(0)                      # no pose yet
        # lanes are independent and resettable
        ref = ko.OracleKissICPWrapper()
        for k in range(3):
            f, t, _, _ = tiny_seq.points(k)
            ref.register_points(f, t, 0.1)
            poses, _ = o.register_frame_batch([f, f], [t, t])
            assert np.array_equal(poses[0], ref.pose) and np.array_equal(poses[1], ref.pose)
        o.reset(1)
        assert o.num_poses(1) == 0 and o.num_poses(0) == 3
        assert odometry.VoxelHashMap(o, 1).empty() and not odometry.VoxelHashMap(o, 0).empty()
        ref2 = ko.OracleKissICPWrapper()
        f, t, _, _ = tiny_seq.points(3)
        ref.register_points(f, t, 0.1)
        ref2.register_points(f, t, 0.1)
        poses, _ = o.register_frame_batch([f, f], [t, t])
        assert np.array_equal(poses[0], ref.pose) and np.array_equal(poses[1], ref2.pose)
        assert np.array_equal(o.get_pose(-1, lane=0), ref.pose) and np.array_equal(o.get_prediction_model(lane=1), np.eye(4))
    finally:
        o.close()


def test_tombstones_and_table_rebuild(tiny_seq):
    """A platform that keeps moving erases as many voxels as it adds: erased table slots become tombstones
    and the table is rebuilt when they pile up.  Ten hops of 300 m with a small map (table of 16384 slots),
    map contents and nearest-neighbour answers checked against the oracle after every hop."""
    from ptudes_lab_b200 import odometry
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    o = odometry.Odometry(cfg, max_points=16384, map_capacity=4096)
    gm = odometry.VoxelHashMap(o, 0)
    rm = ko.VoxelHashMap(1.0, 100.0, 20)
    rng = np.random.default_rng(2)
    try:
        xyz, _, _, _ = tiny_seq.points(0)
        ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)[:3000]
        total_erased = 0
        for hop in range(10):
            pose = canon.se3_exp_mat(np.concatenate([[300.0 * hop, 40.0 * (hop % 3), 0.0], rng.normal(0, 0.2, 3)]))
            before = rm.num_voxels()
            gm.update(ds, pose)
            rm.update(ds, pose)
            if hop:
                total_erased += before               # 300 m away with max_range 100: every old voxel goes
                assert rm.num_voxels() < 1.5 * before
            _map_equal(gm, rm)
            x, y, z = canon.transform_points(pose, ds[::7, 0], ds[::7, 1], ds[::7, 2])
            q = np.stack([x, y, z], 1) + 0.03
            acc, tgt, order = gm.get_correspondences(q, 2.0, return_index=True)
            racc, rtgt, rorder = rm.get_correspondences(q, 2.0, return_index=True)
            assert np.array_equal(acc, racc) and np.array_equal(order[acc], rorder[racc]) and np.array_equal(tgt[acc], rtgt[racc])
        assert total_erased > 8192          # more slots were erased than half the table: a rebuild must have run
    finally:
        o.close()


def test_iteration_cap(os0_seq):
    """MAX_NUM_ITERATIONS_ reached before convergence: the loop stops there and returns T_icp * guess."""
    from ptudes_lab_b200 import odometry
    from ptudes_lab_b200.odometry import VoxelHashMap, register_frame
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    o = odometry.Odometry(cfg, max_points=140000, map_capacity=65536, max_iterations=3)
    try:
        gm = VoxelHashMap(o, 0)
        rm = ko.VoxelHashMap(1.0, 100.0, 20)
        xyz, _, _, _ = os0_seq.points(0)
        ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)
        gm.update(ds, np.eye(4))
        rm.update(ds, np.eye(4))
        src = ko.voxel_down_sample(ds, 1.5)
        guess = canon.se3_exp_mat(np.array([0.2, -0.1, 0.05, 0.02, -0.01, 0.03]))
        rpose, rst = ko.register_point_cloud(src, rm, guess, 6.0, 2.0 / 3.0, max_iters=3)
        pose, st = register_frame(src, gm, guess, 6.0, 2.0 / 3.0, return_stats=True)
        assert rst["iterations"] == 3 and st["iterations"] == 3 and st["dx_norm"] > 1e-4
        assert np.array_equal(pose, rpose)
    finally:
        o.close()


def test_fleet_replay_over_several_contexts(tiny_seq):
    """ptk_fleet_replay: three contexts (2 + 2 + 1 lanes), each advanced by its own thread inside the library, from
    pinned host range images (prefetched one scan ahead) and from device images.  Free-running contexts must give
    exactly the poses of the lock-step batched call, i.e. the oracle's."""
    import torch
    from ptudes_lab_b200 import _ffi, odometry
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    n = 6
    ref = ko.OracleKissICPWrapper()
    want = []
    for k in range(n):
        xyz, ts, tsec, _ = tiny_seq.points(k)
        want.append(ref.register_points(xyz, ts, tsec).copy())
    batches = (2, 2, 1)
    odos = [odometry.Odometry(cfg, max_points=16384, map_capacity=16384, batch=b) for b in batches]
    try:
        for o in odos:
            o.set_sensor(tiny_seq.dirs)
            o.set_icp_blocks_per_lane(6)
        streams = [torch.cuda.Stream() for _ in odos]
        host = []
        for k in range(n):
            h = _ffi.pinned_empty(tiny_seq.scan(k).range_mm.shape, dtype=np.uint32)
            h[...] = tiny_seq.scan(k).range_mm
            host.append(h)
        dev = [torch.as_tensor(tiny_seq.scan(k).range_mm.astype(np.int32), device="cuda:0") for k in range(n)]
        # third pass: the workers outnumber the host cores (affinity narrowed to one core), so they wait for their
        # steps with the yielding poll instead of a spinning synchronize - as 8 ranks x 8 contexts do on a 32-core box
        import os
        cores = os.sched_getaffinity(0)
        for images, narrow in ((host, False), (dev, False), (host, True)):
            for o in odos:
                o.reset()
            ranges = [[[images[k]] * b for k in range(n)] for b in batches]
            if narrow:
                os.sched_setaffinity(0, {min(cores)})
            try:
                poses, stats = odometry.fleet_replay(odos, ranges, [s.cuda_stream for s in streams], want_stats=True)
            finally:
                os.sched_setaffinity(0, cores)
            for g, b in enumerate(batches):
                assert poses[g].shape == (n, b, 4, 4)
                for k in range(n):
                    for l in range(b):
                        assert np.array_equal(poses[g][k, l], want[k]), (g, k, l)
                assert stats[g][n - 1][0]["n_src"] == ref.last_counts["n_src"]
    finally:
        for o in odos:
            o.close()
