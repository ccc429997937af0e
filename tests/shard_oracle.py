"""NumPy backend of the sharded-map loop (ptudes_lab_b200.sharded.ShardedOdometry) for the CPU/gloo
tests: the per-rank work of one rank, restated on the oracle's functions.  TEST CODE (imports oracle/)."""
import math

import numpy as np
import torch

from oracle import canon, kiss_oracle as ko
from ptudes_lab_b200.sharded import NO_ORD, ROOT_COLS, shard_owner, shard_slice

NT = 27


def _tree(a):
    """adjacent-pairs tree over axis 0 of a power-of-two number of rows"""
    while a.shape[0] > 1:
        a = a[0::2] + a[1::2]
    return a[0]


class OracleShardBackend:
    def __init__(self, rank, nranks, min_range=5.0, max_range=100.0):
        self.rank, self.nranks = rank, nranks
        self.device = torch.device("cpu")
        self.w = ko.OracleKissICPWrapper(_min_range=min_range, _max_range=max_range)
        self.iterations = 0

    def begin(self, frame, timestamps, initial_guess=None, range_mm=None):
        k = self.w._kiss
        frame = k.compensator.deskew_scan(frame, self.w.poses, timestamps)
        frame = k.preprocess(frame)
        source, self.frame_ds = k.voxelize(frame)
        self.sigma = k.get_adaptive_threshold()
        if initial_guess is None:
            last = k.poses[-1] if k.poses else np.eye(4)
            initial_guess = canon.rigid_mul(last, k.get_prediction_model())
        self.guess = np.array(initial_guess, dtype=np.float64)
        self.guess_q = canon.SE3q.from_matrix(self.guess)
        self.T = canon.SE3q()
        x, y, z = canon.transform_points(self.guess, source[:, 0], source[:, 1], source[:, 2])
        self.src = np.stack([x, y, z], axis=1)
        self.E = None
        self.status = 0
        self.iterations = 0
        return len(source), k.local_map.num_voxels()

    def search(self, it):
        if it > 0:
            x, y, z = canon.transform_points(self.E, self.src[:, 0], self.src[:, 1], self.src[:, 2])
            self.src = np.stack([x, y, z], axis=1)
        found, best, d2, order = self.w._kiss.local_map.nearest(self.src)
        n = len(self.src)
        rec = np.zeros((5, max(n, 1)))
        rec[0, :n] = np.where(found, d2, np.inf)
        rec[1, :n] = np.where(found, order, NO_ORD)
        rec[2:5, :n] = np.where(found[:, None], best, 0.0).T
        return torch.from_numpy(rec)

    def system(self, gathered, it):
        G = gathered.numpy()
        n = len(self.src)
        d2, od = G[:, 0, :n], G[:, 1, :n]
        br = np.lexsort((od, d2), axis=0)[0]                     # per point: rank with the smallest (d2, ord)
        cols = np.arange(n)
        bd2, bord = d2[br, cols], od[br, cols]
        tgt = G[br, 2:5, cols] if n else np.zeros((0, 3))
        with np.errstate(invalid="ignore"):
            acc = (bord < NO_ORD) & (np.sqrt(bd2) < 3 * self.sigma)
        terms = ko.linear_system_terms(self.src, tgt, acc, self.sigma / 3)
        g_lo, g_cnt, _ = shard_slice(n, self.nranks, self.rank)
        part = np.zeros((NT + 1, ROOT_COLS))
        if g_cnt:
            rows = np.zeros((g_cnt * 32, NT + 1))
            lo, hi = g_lo * 32, min(n, (g_lo + g_cnt) * 32)
            if hi > lo:
                rows[:hi - lo, :NT] = terms[lo:hi]
                rows[:hi - lo, NT] = acc[lo:hi]
            part[:, self.rank] = _tree(rows)
        return torch.from_numpy(part)

    def solve(self, part, it, map_empty=False):
        if map_empty:
            self.pose = self.T.mul(self.guess_q).matrix()
            return True
        _, _, n_roots = shard_slice(len(self.src), self.nranks, self.rank)
        sums = _tree(part.numpy()[:, :n_roots].T.copy())
        self.iterations = it + 1
        done = False
        if int(sums[NT]) == 0:
            self.status, done = 1, True
        else:
            A, b = ko.unpack_system(sums[:NT])
            dx, ok = canon.ldlt_solve6(A, [-v for v in b])
            if not ok:
                self.status, done = 2, True
            else:
                Eq, self.E = canon.se3_exp_q(np.array(dx))
                self.T = Eq.mul(self.T)
                nrm = math.sqrt(((((dx[0] * dx[0] + dx[1] * dx[1]) + dx[2] * dx[2]) + dx[3] * dx[3]) + dx[4] * dx[4]) + dx[5] * dx[5])
                done = nrm < ko.EST_THRESHOLD or it + 1 >= ko.MAX_ITERS
        if done:
            self.pose = self.T.mul(self.guess_q).matrix()
        return done

    def end(self):
        k = self.w._kiss
        gain = canon.rigid_mul(canon.rigid_inv(self.guess), self.pose)
        k.adaptive_threshold.update_model_deviation(gain)
        x, y, z = canon.transform_points(self.pose, self.frame_ds[:, 0], self.frame_ds[:, 1], self.frame_ds[:, 2])
        pts = np.stack([x, y, z], axis=1)
        mine = shard_owner(ko.pack_keys(ko.voxel_keys(pts, k.local_map.voxel_size)), self.nranks) == self.rank
        k.local_map.add_points(pts[mine])
        k.local_map.remove_far_away_points(self.pose[:3, 3])
        k.poses.append(self.pose)
        return self.pose, {"iterations": self.iterations, "status": self.status, "sigma": self.sigma,
                           "n_src": len(self.src), "n_voxels": k.local_map.num_voxels()}
