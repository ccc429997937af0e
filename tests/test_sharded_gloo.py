"""World-size-2 gloo runs (CPU) of the multi-GPU host logic: the sharded-map registration loop with
its all-gather + all-reduce per iteration (backend: NumPy oracle shard), and the fleet helpers."""
import os
import sys
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_SCANS = 6
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from oracle import kiss_oracle as ko
        from ptudes_lab_b200 import sharded, synth
        from shard_oracle import OracleShardBackend
        seq = synth.make_sequence("tiny", 0)
        so = sharded.ShardedOdometry(OracleShardBackend(rank, world))
        ref = ko.OracleKissICPWrapper()
        for k in range(N_SCANS):
            xyz, ts, tsec, _ = seq.points(k)
            g = None
            if k == 4:      # an injected guess in the middle of the run
                g = ref.pose.copy()
            ref.register_points(xyz, ts, tsec, initial_guess=g)
            pose, st = so.register_frame(xyz, ts, initial_guess=g)
            assert np.array_equal(pose, ref.pose), (rank, k)
            assert st["iterations"] == ref.last_stats["iterations"], (rank, k)
        assert so.collectives >= 2 * sum(1 for _ in range(N_SCANS - 1))
        # the shards partition the unsharded map
        keys, cnt, pts = so.b.w._kiss.local_map.voxel_table()
        own = sharded.shard_owner(ko.pack_keys(keys), world)
        assert (own == rank).all()
        nv = torch.tensor([len(keys), int(cnt.sum())])
        dist.all_reduce(nv)
        rk, rc, _ = ref._kiss.local_map.voxel_table()
        assert nv.tolist() == [len(rk), int(rc.sum())]
        full = sharded.shard_owner(ko.pack_keys(rk), world) == rank
        assert np.array_equal(keys, rk[full]) and np.array_equal(pts, ref._kiss.local_map.voxel_table()[2][full])
        # fleet helpers: 5 sequences dealt to 2 ranks, trajectories gathered everywhere
        ids = sharded.fleet_assign(5, world, rank)
        assert ids == list(range(rank, 5, world))
        local = [np.tile(np.eye(4) * (i + 1), (3, 1, 1)) for i in ids]
        allp = sharded.gather_fleet_poses(ids, local, 5)
        assert allp.shape == (5, 3, 4, 4) and all(allp[i, 0, 0, 0] == i + 1 for i in range(5))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_loop_and_fleet_helpers(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_sharded_loop_world1_equals_oracle(tiny_seq):
    """Without a process group the loop degenerates to the plain registration."""
    from oracle import kiss_oracle as ko
    from ptudes_lab_b200 import sharded
    from shard_oracle import OracleShardBackend
    so = sharded.ShardedOdometry(OracleShardBackend(0, 1))
    ref = ko.OracleKissICPWrapper()
    for k in range(4):
        xyz, ts, tsec, _ = tiny_seq.points(k)
        ref.register_points(xyz, ts, tsec)
        pose, _ = so.register_frame(xyz, ts)
        assert np.array_equal(pose, ref.pose)


def test_slices_are_aligned_subtrees():
    from ptudes_lab_b200.sharded import shard_slice
    for n in (1, 31, 32, 33, 1000, 2244, 10222, 262144):
        for G in (1, 2, 4, 8):
            slices = [shard_slice(n, G, r) for r in range(G)]
            n_roots = slices[0][2]
            assert n_roots & (n_roots - 1) == 0 and 1 <= n_roots <= G
            covered = sum(c for _, c, _ in slices)
            assert covered * 32 >= n
            for lo, c, _ in slices:
                if c:
                    assert c & (c - 1) == 0 and lo % c == 0
