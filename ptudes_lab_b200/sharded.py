"""Multi-GPU forms of the odometry path, one process per GPU over torch.distributed.

1. Fleet replay: independent sequences are dealt to the ranks (`fleet_assign`); nothing crosses
   ranks on the data path; `gather_fleet_poses` collects the trajectories at the end.
2. One sequence with the voxel map sharded by hash key (`ShardedOdometry`): every rank keeps the
   voxels whose key hashes to it, searches only those, and per ICP iteration the ranks exchange
   the per-point nearest-neighbour records (all-gather) and the normal-equation partials
   (all-reduce).  The partial of rank r is the root of an ALIGNED subtree of the canonical reduction
   tree and the all-reduce only ever adds zeros to it, so every rank - and a single GPU - computes
   the same bits.

The orchestration is backend-agnostic: on GPUs the backend is `PtkShardBackend` (libptk's
ptk_shard_* entry points, NCCL); the CPU tests run the very same loop over gloo with a NumPy
backend of their own.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _ffi
from ._ffi import PtkStats, addr

NSUM = 17          # 16 distinct normal-equation sums + correspondence count
ROOT_COLS = 32     # columns of the partial table (ranks <= 32)
NO_ORD = 1 << 30


# ------------------------------------------------------------------------------- fleet replay
def fleet_assign(n_sequences: int, world_size: int, rank: int):
    """Sequence ids of `rank`: round-robin, so 64 sequences give 64/32/16/8 per GPU on 1/2/4/8."""
    return list(range(rank, n_sequences, world_size))


def gather_fleet_poses(local_ids, local_poses, n_sequences: int, group=None, device="cpu"):
    """All ranks' trajectories as one (n_sequences, n_scans, 4, 4) array on every rank.
    `local_poses[i]` is the (n_scans, 4, 4) trajectory of sequence `local_ids[i]`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_scans = int(np.asarray(local_poses[0]).shape[0]) if len(local_poses) else 0
    t = torch.tensor([n_scans], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    n_scans = int(t.item())
    full = torch.zeros((n_sequences, n_scans, 4, 4), dtype=torch.float64, device=device)
    for i, p in zip(local_ids, local_poses):
        full[i] = torch.as_tensor(np.asarray(p), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)     # disjoint rows: x + 0 is exact
    return full.cpu().numpy()


# ------------------------------------------------------------------------------- sharded map
def shard_slice(n_src: int, nranks: int, rank: int):
    """(first group, group count, number of slice roots) of `rank`: the source's 32-point groups,
    padded to a power of two, are cut into equal aligned runs - subtrees of the canonical tree."""
    n_groups = (n_src + 31) // 32
    P = 1
    while P < n_groups:
        P <<= 1
    sg = max(1, P // nranks)
    n_roots = P // sg
    return rank * sg, (sg if rank < n_roots else 0), n_roots


def shard_owner(packed_keys, nranks: int):
    """Owner rank of packed voxel keys (uint64): upper 32 bits of the 64-bit finaliser mix."""
    k = np.asarray(packed_keys).astype(np.uint64)
    with np.errstate(over="ignore"):
        k = k ^ (k >> np.uint64(33))
        k = k * np.uint64(0xff51afd7ed558ccd)
        k = k ^ (k >> np.uint64(33))
        k = k * np.uint64(0xc4ceb9fe1a85ec53)
        k = k ^ (k >> np.uint64(33))
    return ((k >> np.uint64(32)) % np.uint64(nranks)).astype(np.int64)


class PtkShardBackend:
    """libptk behind the five calls the sharded loop needs (device = the context's GPU)."""

    def __init__(self, odo, rank: int, nranks: int, lane: int = 0):
        self.odo, self.lane = odo, lane
        self.lib, self.h = odo._lib, odo._h
        self.rank, self.nranks = rank, nranks
        self.device = torch.device("cuda", torch.cuda.current_device())
        odo._check(self.lib.ptk_shard_config(self.h, rank, nranks))
        self.n_src = 0
        self.max_iterations = int(odo._cfg.max_iterations)

    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def begin(self, frame, timestamps, initial_guess=None, range_mm=None):
        g = None if initial_guess is None else np.ascontiguousarray(np.asarray(initial_guess, dtype=np.float64).reshape(4, 4))
        ns, nv = C.c_int(0), C.c_int(0)
        if range_mm is not None:
            r = self.odo._u32(range_mm)
            rc = self.lib.ptk_shard_begin(self.h, self.lane, None, None, 0, addr(r), addr(g), C.byref(ns), C.byref(nv), self._stream())
        else:
            f, t = _ffi.f64(frame), _ffi.f64(timestamps)
            rc = self.lib.ptk_shard_begin(self.h, self.lane, addr(f), addr(t), int(f.shape[0]), None, addr(g),
                                          C.byref(ns), C.byref(nv), self._stream())
        self.odo._check(rc)
        self.n_src = ns.value
        return ns.value, nv.value

    def search(self, it):
        rec = torch.empty((5, max(self.n_src, 1)), dtype=torch.float64, device=self.device)
        self.odo._check(self.lib.ptk_shard_search(self.h, self.lane, it, rec.data_ptr(), self._stream()))
        return rec

    def system(self, gathered, it):
        part = torch.empty((NSUM, ROOT_COLS), dtype=torch.float64, device=self.device)
        self.odo._check(self.lib.ptk_shard_system(self.h, self.lane, gathered.data_ptr(), it, part.data_ptr(), self._stream()))
        return part

    def solve(self, partials, it, map_empty=False):
        done = C.c_int(0)
        self.odo._check(self.lib.ptk_shard_solve(self.h, self.lane, None if partials is None else partials.data_ptr(), it,
                                                 1 if map_empty else 0, C.byref(done), self._stream()))
        return bool(done.value)

    def end(self):
        pose = np.empty((4, 4))
        st = PtkStats()
        self.odo._check(self.lib.ptk_shard_end(self.h, self.lane, addr(pose), C.byref(st), self._stream()))
        return pose, st.as_dict()


class ShardedOdometry:
    """register_frame over a hash-sharded map: the loop of kiss-icp's RegisterFrame with two
    collectives per iteration.  `backend` does the per-rank work, `group` the exchange."""

    def __init__(self, backend, group=None, max_iterations: int = None):
        self.b = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # the loop must run exactly as long as the device-side cap (L.max_iters) allows: k_shard_solve only writes
        # the pose on its last iteration
        self.max_iterations = max_iterations if max_iterations is not None else int(getattr(backend, "max_iterations", 500))
        self.collectives = 0
        self.scans = 0

    def _all_gather(self, rec):
        if self.world == 1:
            return rec.unsqueeze(0).contiguous()
        out = torch.empty((self.world * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
        dist.all_gather_into_tensor(out, rec.contiguous(), group=self.group)     # concatenated along dim 0
        self.collectives += 1
        return out.view((self.world,) + tuple(rec.shape))

    def _all_reduce(self, t, op=dist.ReduceOp.SUM):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)
            self.collectives += 1
        return t

    def describe(self):
        """What the exchange is and how much of it there was (for bench.py's sharded line)."""
        return {"exchange": "host-driven loop: NCCL all-gather of the per-point records + all-reduce of the [17][32] "
                            "partial table per ICP iteration",
                "collectives_per_scan": self.collectives / max(self.scans, 1)}

    def close(self):
        odo = getattr(self.b, "odo", None)
        if odo is not None:
            odo.close()

    def register_frame(self, frame, timestamps, initial_guess=None, range_mm=None):
        self.scans += 1
        n_src, n_vox_local = self.b.begin(frame, timestamps, initial_guess, range_mm=range_mm)
        total_vox = int(self._all_reduce(torch.tensor([n_vox_local], dtype=torch.int64, device=self.b.device)).item())
        if total_vox == 0:                       # RegisterFrame: empty map -> the guess
            self.b.solve(None, 0, map_empty=True)
            return self.b.end()
        for it in range(self.max_iterations):
            rec = self.b.search(it)
            gathered = self._all_gather(rec)
            part = self.b.system(gathered, it)
            part = self._all_reduce(part)
            if self.b.solve(part, it):
                break
        return self.b.end()


class PeerShardedOdometry:
    """One rank of a hash-sharded single-sequence odometry whose per-iteration exchange happens INSIDE the ICP kernel,
    through peer memory (NVLink): every rank steps the same scans with the ordinary register_scan / register_frame
    calls, the kernels of the ranks trade per-point records and stamps directly - no collective, no host in the loop.
    torch.distributed is used once, to pass the CUDA IPC handles around (`connect`)."""

    def __init__(self, odo, rank: int, nranks: int):
        self.odo, self.rank, self.nranks = odo, rank, nranks
        self.lib, self.h = odo._lib, odo._h
        odo._check(self.lib.ptk_shard_config(self.h, rank, nranks))
        self.scans = 0
        self.connected = nranks == 1

    def export(self):
        """(IPC handle bytes, device pointer) of this rank's exchange buffer."""
        handle = C.create_string_buffer(64)
        ptr = C.c_void_p()
        self.odo._check(self.lib.ptk_shard_peer_export(self.h, handle, C.byref(ptr), None))
        return handle.raw, ptr.value

    def attach(self, rank: int, handle: bytes = None, pointer: int = None):
        self.odo._check(self.lib.ptk_shard_peer_attach(self.h, rank, handle, pointer))

    def connect(self, group=None):
        """Exchange the IPC handles over torch.distributed (one all-gather of 64 bytes per rank) and map every peer."""
        if self.nranks == 1:
            return self
        handle, _ = self.export()
        t = torch.tensor(list(handle), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        out = [torch.empty_like(t) for _ in range(self.nranks)]
        dist.all_gather(out, t, group=group)
        for r, h in enumerate(out):
            if r != self.rank:
                self.attach(r, handle=bytes(h.cpu().tolist()))
        dist.barrier(group=group)
        self.connected = True
        return self

    def register_frame(self, frame, timestamps, initial_guess=None, range_mm=None, stream=None):
        """The ordinary step; `stream` must be a non-default stream (the ranks' ICP kernels wait for each other)."""
        assert self.connected, "PeerShardedOdometry.connect() first"
        if stream is None:
            if not hasattr(self, "_stream"):
                self._stream = torch.cuda.Stream()
            stream = self._stream.cuda_stream
        self.scans += 1
        if range_mm is not None:
            return self.odo.register_scan(range_mm, initial_guess=initial_guess, stream=stream)
        return self.odo.register_frame(frame, timestamps, initial_guess=initial_guess, stream=stream)

    def describe(self):
        return {"exchange": "inside the ICP kernel: per-point (d2, order id, target) records written straight into the "
                            "peers' memory (CUDA IPC over NVLink) with stamp flags; no collective, no host round trip",
                "collectives_per_scan": 0.0}

    def close(self):
        self.odo.close()


def make_sharded(cfg, device: int, rank: int, world: int, *, max_points: int, map_capacity: int, dirs=None, group=None,
                 mode: str = "peer"):
    """One rank's share of a hash-sharded single-sequence odometry.  mode="peer": exchange inside the kernel through
    peer memory (the product path on NVLink-connected GPUs); mode="nccl": the host-driven loop with two NCCL
    collectives per ICP iteration (kept as the portable comparison point, and what the gloo tests exercise)."""
    from . import odometry
    if mode == "peer" and world > 1:
        odo = odometry.Odometry(cfg, device=device, max_points=max_points, map_capacity=map_capacity)
        if dirs is not None:
            odo.set_sensor(dirs)
        return PeerShardedOdometry(odo, rank, world).connect(group)
    return _make_sharded_nccl(cfg, device, rank, world, max_points=max_points, map_capacity=map_capacity, dirs=dirs, group=group)


def _make_sharded_nccl(cfg, device: int, rank: int, world: int, *, max_points: int, map_capacity: int, dirs=None, group=None):
    """One rank's share of a hash-sharded single-sequence odometry: its own context (the local map holds only
    the voxels this rank owns) behind the sharded loop.  `dirs`: XYZLut directions for range-image input."""
    from . import odometry
    odo = odometry.Odometry(cfg, device=device, max_points=max_points, map_capacity=map_capacity)
    if dirs is not None:
        odo.set_sensor(dirs)
    return ShardedOdometry(PtkShardBackend(odo, rank, world), group=group)
