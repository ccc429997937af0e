"""order="robin_map": emulation of the hash-table iteration order upstream kiss-icp 0.2.x produces
(tsl::robin_map), used to MEASURE what the canonical ordering rules B.1/B.2/B.4 cost (DESIGN.md section 2;
profiles/r2_order_delta.py / .json hold the 100-scan numbers).  Reference call sites whose results depend
on that order: /root/reference/src/ptudes/kiss.py:96 (voxelize) and :129 (local_map.update)."""
import numpy as np
import pytest

from oracle import canon, kiss_oracle as ko
from ptudes_lab_b200 import synth


@pytest.fixture(scope="module")
def tiny_seq():
    return synth.make_sequence("tiny", 0)


def test_upstream_hash_known_answers():
    # ((1 << 20) - 1) & (x*73856093 ^ y*19349663 ^ z*83492791) on uint32 lanes
    assert ko.upstream_voxel_hash(np.array([[1, 2, 3], [-1, 0, 0], [0, 0, 0]])).tolist() == [363078, 592803, 0]


def test_bucket_counts():
    # reserve(n): next power of two >= 2 n (max_load_factor 0.5)
    assert [ko.robin_bucket_count_reserve(n) for n in (1, 2, 3, 1000, 91132, 131072)] == [2, 4, 8, 2048, 262144, 262144]
    # unreserved map: doubles from 0 -> 2 whenever size() >= bucket_count / 2 at the insertion of a new key
    t = ko.RobinTable()
    seen = []
    for k in range(100):
        t.insert_new(k, k * 7919)
        seen.append(t.bc)
    assert [seen[s - 1] for s in (1, 2, 3, 4, 5, 8, 9, 100)] == [2, 4, 8, 8, 16, 16, 32, 256]


def test_robin_hood_layout_invariants():
    """Ideal buckets ascend along every cluster, nobody sits before its ideal bucket, and a DISPLACED entry
    travels past the residents that are as far from home as it is (strict comparison in tsl's swap)."""
    rng = np.random.default_rng(1)
    for _ in range(50):
        bc = 64
        h = rng.integers(0, 1 << 20, size=int(rng.integers(1, 33)))
        order, pos = ko.robin_layout(h, bc)
        assert sorted(order.tolist()) == list(range(len(h))) and np.all(np.diff(pos) > 0)
        dist = (pos - (h[order] & (bc - 1))) % bc
        for a in range(len(pos) - 1):
            if pos[a + 1] == pos[a] + 1:                # neighbours in one cluster: distance grows by at most one
                assert dist[a + 1] <= dist[a] + 1
            else:
                pass
        first_of_cluster = np.r_[True, np.diff(pos) > 1]
        if pos[0] != 0 or pos[-1] != bc - 1:            # (a cluster that wraps starts before bucket 0)
            assert np.all(dist[first_of_cluster] == 0)
    # the worked example: 3 sits in bucket 60, 12 in 59, 15 and 16 (ideal 60) behind 3; inserting 31 (ideal 58,
    # bucket 58 taken by a poorer entry) pushes 12 forward, 12 pushes 3, and 3 lands BEHIND 15 and 16
    t = ko.RobinTable()
    t.bc = 64
    for ident, ideal in ((3, 60), (12, 59), (15, 60), (16, 60), (7, 57), (24, 57), (31, 58)):
        t._place(ident, ideal, ideal, 0)
    assert t.iteration() == [7, 24, 31, 12, 15, 16, 3]


def test_downsample_same_set_other_order(tiny_seq):
    xyz, ts, _, _ = tiny_seq.points(3)
    frame = ko.preprocess(xyz, 100.0, 5.0)
    a = ko.voxel_down_sample_idx(frame, 0.5, "index")
    b = ko.voxel_down_sample_idx(frame, 0.5, "robin_map")
    assert np.array_equal(np.sort(b), a) and not np.array_equal(a, b)
    # bucket order: ideal buckets of the emulated output never decrease
    keys = ko.voxel_keys(frame[b], 0.5)
    ideal = ko.upstream_voxel_hash(keys) & (ko.robin_bucket_count_reserve(len(frame)) - 1)
    assert np.all(np.diff(ideal) >= 0)
    # the second grid sees another "first" point in some voxels: same voxel set, different representatives
    sa = ko.voxel_down_sample(frame[a], 1.5, "index")
    sb = ko.voxel_down_sample(frame[b], 1.5, "robin_map")
    ka = {tuple(k) for k in ko.voxel_keys(sa, 1.5)}
    kb = {tuple(k) for k in ko.voxel_keys(sb, 1.5)}
    assert ka == kb and len(sa) == len(sb)
    assert {tuple(p) for p in sa} != {tuple(p) for p in sb}


def test_erase_while_iterating_skips_the_entry_that_slides_back():
    m = ko.VoxelHashMap(1.0, 10.0, 20, order="robin_map")
    # three voxels whose upstream hashes collide modulo the bucket count, all far from the new origin:
    # erasing the first shifts the second into its bucket, the range-for steps over it, the third goes too
    bc = None
    keys = []
    k = 0
    while len(keys) < 3:
        k += 1
        h = int(ko.upstream_voxel_hash(np.array([[k, 0, 0]]))[0])
        if bc is None:
            bc = 8
            want = h & (bc - 1)
        if (h & (bc - 1)) == want:
            keys.append(k)
    pts = np.array([[kk + 0.5, 0.5, 0.5] for kk in keys])
    m.add_points(pts)
    assert m.num_voxels() == 3 and m.table.bc == bc
    m.remove_far_away_points(np.array([1.0e3, 0.0, 0.0]))
    assert m.num_voxels() == 1 and m.skipped_last_prune == 1
    kept = ko.unpack_keys(m.keys)[0, 0]
    assert kept == keys[1]                       # the middle one (insertion order) survived this call
    m.remove_far_away_points(np.array([1.0e3, 0.0, 0.0]))
    assert m.num_voxels() == 0                   # ... and goes with the next one (B.4: "for at most one scan")
    c = ko.VoxelHashMap(1.0, 10.0, 20)           # canonical rule: everything far goes at once
    c.add_points(pts)
    c.remove_far_away_points(np.array([1.0e3, 0.0, 0.0]))
    assert c.num_voxels() == 0


def test_order_changes_poses_far_beyond_the_bit_level(tiny_seq):
    """The honest statement of rule B.1: both orders are valid kiss-icp runs, and they differ by millimetres to
    centimetres per scan - five orders of magnitude above the 1e-5 m bar - while tracking the ground truth
    equally well."""
    a = ko.OracleKissICPWrapper(order="index")
    b = ko.OracleKissICPWrapper(order="robin_map")
    worst = 0.0
    gts = []
    for k in range(8):
        xyz, ts, tsec, gt = tiny_seq.points(k)
        gts.append(gt)
        pa = a.register_points(xyz, ts, tsec)
        pb = b.register_points(xyz, ts, tsec)
        if k < 2:       # later scans are deskewed with each run's own poses, so even the inputs differ slightly
            assert a.last_counts["n_ds"] == b.last_counts["n_ds"]       # same first-grid voxel set
        d = canon.rigid_mul(canon.rigid_inv(pa), pb)
        worst = max(worst, float(np.linalg.norm(d[:3, 3])))
    assert 1e-5 < worst < 0.1, worst
    from ptudes_lab_b200.ins.data import calc_ate
    ea, eb = calc_ate(a.poses, gts)[1], calc_ate(b.poses, gts)[1]
    assert ea < 0.1 ** 2 and eb < 0.1 ** 2 and 0.5 < ea / eb < 2.0, (ea, eb)    # mean squared metres (calc_ate)
