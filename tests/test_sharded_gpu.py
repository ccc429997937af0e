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


def test_peer_exchange_two_ranks_on_one_gpu():
    """The in-kernel exchange of the sharded mode (ptk_shard_peer_*), exercised on ONE GPU: two contexts play rank 0
    and rank 1, each driven by its own host thread on its own non-default stream, their exchange buffers attached to
    each other by pointer.  The two cooperative ICP kernels run side by side and trade records through the buffers
    exactly as they do over NVLink between two processes.  Poses of both ranks must equal the oracle's and an
    unsharded context's bit for bit, and each shard must hold exactly its share of the map."""
    import threading
    import torch
    from oracle import kiss_oracle as ko
    from ptudes_lab_b200 import odometry, sharded, synth
    seq = synth.make_sequence("tiny", 0)
    cfg = odometry.load_config(None, deskew=True, max_range=100.0)
    n = 6
    odos = [odometry.Odometry(cfg, device=0, max_points=16384, map_capacity=16384, trace_iterations=3) for _ in range(2)]
    single = odometry.Odometry(cfg, device=0, max_points=16384, map_capacity=16384, trace_iterations=3)
    ranks = []
    try:
        for r, o in enumerate(odos):
            o.set_sensor(seq.dirs)
            o.set_icp_blocks_per_lane(8)                   # both launches must fit the device together
            ranks.append(sharded.PeerShardedOdometry(o, r, 2))
        ptrs = [rk.export()[1] for rk in ranks]
        for r, rk in enumerate(ranks):
            rk.attach(1 - r, pointer=ptrs[1 - r])
            rk.connected = True
        scans = [torch.as_tensor(seq.scan(k).range_mm.astype(np.int32), device="cuda:0") for k in range(n)]
        streams = [torch.cuda.Stream() for _ in range(2)]
        out = [[], []]
        errs = []

        def work(r):
            try:
                for k in range(n):
                    if k % 2:
                        xyz, ts, _, _ = seq.points(k)
                        out[r].append(ranks[r].register_frame(xyz, ts, stream=streams[r].cuda_stream))
                    else:
                        out[r].append(ranks[r].register_frame(None, None, range_mm=scans[k], stream=streams[r].cuda_stream))
            except Exception as e:      # noqa: BLE001
                errs.append((r, e))
        th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
        [t.start() for t in th]
        [t.join(timeout=120) for t in th]
        assert not any(t.is_alive() for t in th), "sharded ranks did not finish (exchange deadlock?)"
        assert not errs, errs
        ref = ko.OracleKissICPWrapper()
        for k in range(n):
            xyz, ts, tsec, _ = seq.points(k)
            ref.register_points(xyz, ts, tsec)
            p1, s1 = single.register_frame(xyz, ts)
            for r in range(2):
                pose, st = out[r][k]
                assert np.array_equal(pose, ref.pose), (r, k)
                assert np.array_equal(pose, p1), (r, k)
                assert st["iterations"] == s1["iterations"] and st["n_corr"] == s1["n_corr"]
        # per-iteration correspondence ids of the last scan: what the ranks agreed on is what one GPU finds
        t1 = single.get_trace()
        for r in range(2):
            assert np.array_equal(odos[r].get_trace(), t1)
        fk, fc, fp = odometry.VoxelHashMap(single, 0).dump()
        for r in range(2):
            keys, cnt, pts = odometry.VoxelHashMap(odos[r], 0).dump()
            mine = sharded.shard_owner(ko.pack_keys(fk), 2) == r
            assert 0 < mine.sum() < len(fk)
            assert np.array_equal(keys, fk[mine]) and np.array_equal(cnt, fc[mine]) and np.array_equal(pts, fp[mine])
    finally:
        for o in odos:
            o.close()
        single.close()


def _run_peer_rank(rank, world, port, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tests")]
    import torch
    import torch.distributed as dist
    from oracle import kiss_oracle as ko
    from ptudes_lab_b200 import odometry, sharded, synth
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        seq = synth.make_sequence("os0_quad", 0)
        cfg = odometry.load_config(None, deskew=True, max_range=100.0)
        so = sharded.make_sharded(cfg, rank, rank, world, max_points=140000, map_capacity=65536, dirs=seq.dirs, mode="peer")
        ref = ko.OracleKissICPWrapper()
        for k in range(N_SCANS):
            xyz, ts, tsec, _ = seq.points(k)
            ref.register_points(xyz, ts, tsec)
            pose, st = so.register_frame(None, None, range_mm=seq.scan(k).range_mm)
            assert np.array_equal(pose, ref.pose), (rank, k)
        so.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("peer")
    finally:
        dist.destroy_process_group()


def test_peer_exchange_two_gpus_ipc(tmp_path):
    """Two processes, two GPUs, CUDA IPC: the product form of the sharded mode."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_run_peer_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
