"""The hash-sharded map on real GPUs: world size 1 through the ptk_shard_* entry points (runs on
any GPU box) and world size 2 over NCCL (needs two GPUs).  Poses must equal the single-GPU step
and the oracle bit for bit (the partial sums are combined in the canonical tree order)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N_SCANS = 6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_rank(rank, world, port, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch
    import torch.distributed as dist
    from oracle import kiss_oracle as ko
    from ptudes_lab_b200 import odometry, sharded, synth
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
    try:
        seq = synth.make_sequence("os0_quad", 0)
        cfg = odometry.load_config(None, deskew=True, max_range=100.0)
        o = odometry.Odometry(cfg, device=rank, max_points=140000, map_capacity=65536, trace_iterations=4)
        single = odometry.Odometry(cfg, device=rank, max_points=140000, map_capacity=65536)
        o.set_sensor(seq.dirs)
        so = sharded.ShardedOdometry(sharded.PtkShardBackend(o, rank, world))
        ref = ko.OracleKissICPWrapper()
        for k in range(N_SCANS):
            xyz, ts, tsec, _ = seq.points(k)
            ref.register_points(xyz, ts, tsec)
            p1, s1 = single.register_frame(xyz, ts)
            if k % 2:
                pose, st = so.register_frame(None, None, range_mm=seq.scan(k).range_mm)   # range-image input
            else:
                pose, st = so.register_frame(xyz, ts)
            assert np.array_equal(pose, ref.pose), (rank, k)
            assert np.array_equal(pose, p1), (rank, k)
            assert st["iterations"] == s1["iterations"] and st["n_src"] == s1["n_src"] and st["sigma"] == s1["sigma"]
        # this rank's shard holds exactly its share of the single-GPU map
        keys, cnt, pts = odometry.VoxelHashMap(o, 0).dump()
        fk, fc, fp = odometry.VoxelHashMap(single, 0).dump()
        mine = sharded.shard_owner(ko.pack_keys(fk), world) == rank
        assert np.array_equal(keys, fk[mine]) and np.array_equal(cnt, fc[mine]) and np.array_equal(pts, fp[mine])
        if world > 1:
            assert so.collectives > 2 * N_SCANS
        o.close()
        single.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write(f"collectives {so.collectives}")
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_sharded_entry_points_world1(tmp_path):
    _run_rank(0, 1, 0, str(tmp_path))
    assert (tmp_path / "ok0").exists()


def test_sharded_map_two_gpus_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_run_rank, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
