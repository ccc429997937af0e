"""The two CPU oracles against each other: oracle/kiss_oracle.py (NumPy) and oracle/kiss_port.c
(plain C, written independently from SURVEY Appendix A/B and oracle/canon.py's operation order).
With no runnable reference, two independent restatements agreeing bit for bit is the strongest
pin available."""
import numpy as np
import pytest

from oracle import canon, kiss_oracle as ko, port


@pytest.fixture(scope="module")
def prt():
    p = port.PortKissICP(threads=2, trace_iterations=8)
    yield p
    p.close()


def test_det_sincos_and_se3(prt):
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(-40, 40, 5000), [0.0, 1e-12, -1e-9, np.pi / 4, -np.pi / 2, 1e4]])
    s, c = port.det_sincos(x)
    s2, c2 = canon.det_sincos(x)
    assert np.array_equal(s, s2) and np.array_equal(c, c2)
    for _ in range(200):
        tg = np.concatenate([rng.uniform(-2, 2, 3), rng.uniform(-1, 1, 3) * rng.choice([1.0, 1e-3, 1e-11])])
        T = canon.se3_exp_mat(tg)
        assert np.array_equal(port.se3_exp_mat(tg), T)
        assert np.array_equal(port.se3_log(T), canon.se3_log(T))


def test_pieces_match(prt, tiny_seq):
    xyz, ts, _, _ = tiny_seq.points(3)
    rng = np.random.default_rng(5)
    a = canon.se3_exp_mat(rng.normal(0, 0.3, 6))
    b = canon.rigid_mul(a, canon.se3_exp_mat(rng.normal(0, 0.05, 6)))
    assert np.array_equal(prt.deskew_scan(xyz, ts, a, b), ko.deskew_scan(xyz, ts, a, b))
    assert np.array_equal(prt.preprocess(xyz, 100.0, 5.0), ko.preprocess(xyz, 100.0, 5.0))
    for v in (0.5, 1.5, 0.35, 1.0499999999999998):
        out, idx = prt.voxel_down_sample(xyz, v, return_index=True)
        ridx = ko.voxel_down_sample_idx(xyz, v)
        assert np.array_equal(idx, ridx) and np.array_equal(out, xyz[ridx])
    # edge cases: empty cloud, all points in one voxel, points on the axis planes (trunc, not floor)
    assert prt.voxel_down_sample(np.zeros((0, 3)), 1.0).shape == (0, 3)
    one = rng.uniform(0.1, 0.9, (50, 3))
    assert np.array_equal(prt.voxel_down_sample(one, 1.0), one[:1])
    planes = np.array([[-0.3, 0.3, 0.0], [0.3, -0.3, 0.0], [-1.2, 0.0, 0.0], [1.2, 0.0, 0.0], [-0.999999, 0.7, 1.4]])
    assert np.array_equal(prt.voxel_down_sample(planes, 1.0), ko.voxel_down_sample(planes, 1.0))
    with pytest.raises(ValueError):
        prt.voxel_down_sample(np.array([[2.0e6, 0.0, 0.0]]), 1.0)


def test_map_and_correspondences_match(prt, tiny_seq):
    prt.map_clear()
    rm = ko.VoxelHashMap(1.0, 100.0, 20)
    rng = np.random.default_rng(7)
    for k in range(4):
        xyz, _, _, _ = tiny_seq.points(k)
        ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)
        pose = canon.se3_exp_mat(np.concatenate([rng.uniform(-30, 30, 3) * (k > 1), rng.normal(0, 0.1, 3)]))
        prt.map_update(ds, pose)
        rm.update(ds, pose)
        for a, b in zip(prt.voxel_table(), rm.voxel_table()):
            assert np.array_equal(a, b)
    # 25 points into one voxel keep the first 20 in order; exact-radius voxel survives the prune
    prt.map_clear()
    rm.clear()
    pts = np.concatenate([rng.uniform(10.1, 10.9, (25, 3)), [[100.0, 0.0, 0.0]], [[100.0000001, 1.5, 0.0]]])
    prt.map_add_points(pts)
    rm.add_points(pts)
    prt.map_remove_far(np.zeros(3))
    rm.remove_far_away_points(np.zeros(3))
    for a, b in zip(prt.voxel_table(), rm.voxel_table()):
        assert np.array_equal(a, b)
    assert prt.voxel_table()[1].tolist() == [20, 1]
    # correspondences incl. a tie (two map points at the same distance) and a query without neighbours
    prt.map_clear()
    rm.clear()
    xyz, _, _, _ = tiny_seq.points(0)
    ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)
    tie = np.array([[20.25, 20.5, 0.5], [20.75, 20.5, 0.5]])
    prt.map_add_points(np.concatenate([ds, tie]))
    rm.add_points(np.concatenate([ds, tie]))
    q = np.concatenate([ko.voxel_down_sample(ds, 1.5) + 0.05, [[20.5, 20.5, 0.5]], [[500.0, 500.0, 500.0]]])
    order, tgt = prt.get_correspondences(q, 1.0)
    acc, rtgt, rorder = rm.get_correspondences(q, 1.0, return_index=True)
    assert np.array_equal(order, np.where(acc, rorder, -1))
    assert np.array_equal(tgt[acc], rtgt[acc])
    assert order[-1] == -1 and np.array_equal(tgt[-2], tie[0])


def test_registration_matches(prt, tiny_seq):
    prt.map_clear()
    rm = ko.VoxelHashMap(1.0, 100.0, 20)
    xyz, _, _, _ = tiny_seq.points(0)
    ds = ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5)
    prt.map_update(ds, np.eye(4))
    rm.update(ds, np.eye(4))
    xyz, _, _, _ = tiny_seq.points(3)
    src = ko.voxel_down_sample(ko.voxel_down_sample(ko.preprocess(xyz, 100.0, 5.0), 0.5), 1.5)
    guess = canon.se3_exp_mat(np.array([0.1, -0.05, 0.02, 0.01, -0.01, 0.02]))
    trace = []
    rpose, rst = ko.register_point_cloud(src, rm, guess, 6.0, 2.0 / 3.0, trace=trace)
    pose, st = prt.register_point_cloud(src, guess, 6.0, 2.0 / 3.0)
    assert np.array_equal(pose, rpose)
    assert (st["iterations"], st["n_corr"], st["status"]) == (rst["iterations"], rst["n_corr"], rst["status"])
    tr = prt.get_trace()
    for it in range(tr.shape[0]):
        assert np.array_equal(tr[it], trace[it]["order"])
    # empty map -> guess (through the quaternion, like Sophus); no correspondence -> status 1 (B.5)
    prt.map_clear()
    pose, st = prt.register_point_cloud(src, guess, 6.0, 0.6)
    assert np.array_equal(pose, ko.register_point_cloud(src, ko.VoxelHashMap(1.0, 100.0, 20), guess, 6.0, 0.6)[0])
    prt.map_add_points(np.array([[900.0, 900.0, 900.0]]))
    pose, st = prt.register_point_cloud(src, guess, 6.0, 0.6)
    assert st["status"] == 1 and st["n_corr"] == 0


@pytest.mark.parametrize("ranges", [(5.0, 100.0), (1.0, 70.0)])
def test_sequences_match(tiny_seq, ranges):
    mn, mx = ranges
    ref = ko.OracleKissICPWrapper(_min_range=mn, _max_range=mx)
    p = port.PortKissICP(_min_range=mn, _max_range=mx, threads=3)
    frames = [tiny_seq.points(k) for k in range(7)]
    # an empty scan and a scan entirely out of range in the middle of the run
    frames.insert(3, (np.zeros((0, 3)), np.zeros(0), 0.35, None))
    frames.insert(5, (np.full((40, 3), 500.0), np.full(40, 0.5), 0.55, None))
    for xyz, ts, tsec, _ in frames:
        a = ref.register_points(xyz, ts, tsec)
        b = p.register_points(xyz, ts, tsec)
        assert np.array_equal(a, b)
        assert ref.last_stats["iterations"] == p.last_stats["iterations"]
    assert ref._sigmas == p._sigmas and ref._err_dt == p._err_dt and ref._err_drot == p._err_drot
    assert np.array_equal(ref._kiss.get_prediction_model(), p.get_prediction_model())
    for a, b in zip(p.voxel_table(), ref._kiss.local_map.voxel_table()):
        assert np.array_equal(a, b)
    p.close()


def test_thread_count_does_not_change_results(tiny_seq):
    outs = []
    for th in (1, 4):
        p = port.PortKissICP(threads=th)
        for k in range(5):
            xyz, ts, tsec, _ = tiny_seq.points(k)
            p.register_points(xyz, ts, tsec)
        outs.append(np.stack(p.poses))
        p.close()
    assert np.array_equal(outs[0], outs[1])
