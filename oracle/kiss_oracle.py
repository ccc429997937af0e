"""CPU oracle: NumPy float64 restatement of the kiss-icp 0.2.x odometry step that
ptudes-lab's KissICPWrapper drives.

TEST INFRASTRUCTURE ONLY - see oracle/canon.py for who may import this package.

PARITY UNPINNED: the reference holds no tests / golden vectors for this path and the
arithmetic lives in the absent third-party package kiss-icp (effective 0.2.9/0.2.10,
/root/reference/setup.py:22).  Orchestration follows /root/reference/src/ptudes/kiss.py:83-131
line by line; the kiss-icp internals follow SURVEY.md Appendix A, with the canonical
rules of Appendix B where upstream is implementation defined:

 B.1 downsample output in ascending input index (selection = first point per voxel)
 B.2 per-voxel insertion order = canonical order; slot = count + rank; cap 20
 B.3 query with no neighbour -> no correspondence
 B.4 prune erases every voxel whose first point fails the test
 B.5 zero correspondences -> ICP stops, pose = T_icp * guess accumulated so far
 B.6 NN ties -> first in (i,j,l) voxel order then stored order
 B.7 JtJ/Jtr reduced by a fixed adjacent-pairs binary tree over source index
 B.8 point_cloud order unspecified (compare as sets)

order="robin_map" (second mode, used only to MEASURE what rules B.1/B.2/B.4 cost against upstream):
emulates the iteration order of the tsl::robin_map upstream keeps its voxels in - power-of-two
bucket count, reserve(frame.size()) in VoxelDownsample, the 20-bit upstream voxel hash of
SURVEY A.5, robin-hood linear probing (whose layout is the stable sort of the entries by ideal
bucket) and the erase-while-iterating skip of RemovePointsFarFromLocation (backward-shift
deletion).  Still UPSTREAM-UNVERIFIED: it restates published behaviour of kiss-icp 0.2.10 and
tsl::robin_map 1.x from memory; nothing here was compared with the real packages.
"""
import math

import numpy as np

from . import canon

KEY_BIAS = 1 << 20          # voxel coordinates must satisfy |k| < 2^20
MAX_ITERS = 500             # upstream MAX_NUM_ITERATIONS_
EST_THRESHOLD = 1e-4        # upstream ESTIMATION_THRESHOLD_


# ---------------------------------------------------------------------------
# voxel keys
# ---------------------------------------------------------------------------
def voxel_keys(points, size):
    """(N,3) int32 keys: truncation toward zero of p/size (A.5, A.6, A.7)."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    return (pts / np.float64(size)).astype(np.int32)


def pack_keys(keys):
    k = keys.astype(np.int64)
    if k.size and (np.abs(k).max() >= KEY_BIAS):
        raise ValueError("voxel coordinate out of range (|k| >= 2^20)")
    return ((k[:, 0] + KEY_BIAS) << 42) | ((k[:, 1] + KEY_BIAS) << 21) | (k[:, 2] + KEY_BIAS)


def unpack_keys(packed):
    p = np.asarray(packed, dtype=np.int64)
    m = (1 << 21) - 1
    return np.stack([((p >> 42) & m) - KEY_BIAS, ((p >> 21) & m) - KEY_BIAS,
                     (p & m) - KEY_BIAS], axis=1).astype(np.int32)


# ---------------------------------------------------------------------------
# A.2 deskew, A.4 preprocess, A.5 downsample
# ---------------------------------------------------------------------------
def deskew_scan(frame, timestamps, start_pose, finish_pose):
    """kiss-icp DeSkewScan: p_i <- exp((t_i - 0.5) * log(start^-1 finish)) p_i.
    Call site: /root/reference/src/ptudes/kiss.py:90 (and :76-78)."""
    frame = np.asarray(frame, dtype=np.float64).reshape(-1, 3)
    ts = np.asarray(timestamps, dtype=np.float64).reshape(-1)
    delta = canon.se3_log(canon.rigid_mul(canon.rigid_inv(start_pose), finish_pose))
    return deskew_with_delta(frame, ts, delta)


def deskew_with_delta(frame, ts, delta):
    s = ts - 0.5
    tang = s[:, None] * delta[None, :]
    R, t = canon.se3_exp(tang)
    x, y, z = frame[:, 0], frame[:, 1], frame[:, 2]
    out = np.empty_like(frame)
    out[:, 0] = ((R[:, 0, 0] * x + R[:, 0, 1] * y) + R[:, 0, 2] * z) + t[:, 0]
    out[:, 1] = ((R[:, 1, 0] * x + R[:, 1, 1] * y) + R[:, 1, 2] * z) + t[:, 1]
    out[:, 2] = ((R[:, 2, 0] * x + R[:, 2, 1] * y) + R[:, 2, 2] * z) + t[:, 2]
    return out


def range_mask(frame, max_range, min_range):
    x, y, z = frame[:, 0], frame[:, 1], frame[:, 2]
    norm = np.sqrt((x * x + y * y) + z * z)
    return (norm < max_range) & (norm > min_range)


def preprocess(frame, max_range, min_range):
    """kiss-icp Preprocess: keep min < |p| < max (strict), order preserved
    (/root/reference/src/ptudes/kiss.py:93)."""
    frame = np.asarray(frame, dtype=np.float64).reshape(-1, 3)
    return frame[range_mask(frame, max_range, min_range)]


# Second multiplier of upstream's VoxelHash.  SURVEY A.5 gives the classic spatial-hash prime 19349663 (Teschner et
# al.); kiss-icp's own VoxelHashMap.hpp is remembered by some as carrying 19349669 instead (a long-standing typo of
# that prime).  Neither can be checked here.  It only matters to the order="robin_map" EMULATION (which bucket a voxel
# lands in); the measured size of the ordering effect is the same with either (profiles/r2_order_delta*.json).
UPSTREAM_HASH_Y = 19349663


def upstream_voxel_hash(keys, hash_y=None):
    """kiss-icp VoxelHash (A.5): ((1 << 20) - 1) & (x*73856093 ^ y*19349663 ^ z*83492791) on the int32
    lanes reinterpreted as uint32.  Known answers: (1,2,3) -> 363078, (-1,0,0) -> 592803."""
    hy = UPSTREAM_HASH_Y if hash_y is None else hash_y
    k = np.asarray(keys).reshape(-1, 3).astype(np.int64) & 0xFFFFFFFF
    h = ((k[:, 0] * 73856093) & 0xFFFFFFFF) ^ ((k[:, 1] * hy) & 0xFFFFFFFF) ^ ((k[:, 2] * 83492791) & 0xFFFFFFFF)
    return (h & ((1 << 20) - 1)).astype(np.int64)


ROBIN_DIST_LIMIT = 4096     # tsl::robin_map grows when an entry sits farther than this from its ideal bucket


def robin_bucket_count_reserve(n):
    """bucket_count() after tsl::robin_map::reserve(n): next power of two >= ceil(n / max_load_factor 0.5)."""
    want = 2 * int(n)
    b = 1
    while b < want:
        b <<= 1
    return b if n > 0 else 0


class RobinTable:
    """tsl::robin_map as far as ITERATION ORDER goes: power-of-two bucket count, max_load_factor 0.5, linear
    probing with robin-hood swaps (an entry being inserted or displaced takes a bucket only from a resident that
    is strictly closer to its own ideal bucket - so a displaced entry travels past its equals), growth by
    doubling with re-insertion in old bucket order, backward-shift deletion.  Values are opaque ids."""

    def __init__(self, reserve=0):
        self.bc = robin_bucket_count_reserve(reserve)
        self.slot = {}                  # bucket -> [id, hash, distance from ideal bucket]
        self.where = {}                 # id -> bucket
        self.grow_next = False

    def __len__(self):
        return len(self.slot)

    def _place(self, cur, h, b, d):
        mask = self.bc - 1
        slot, where = self.slot, self.where
        while True:
            r = slot.get(b)
            if r is None:
                slot[b] = [cur, h, d]
                where[cur] = b
                return
            if r[2] < d:                # resident is richer: it moves on, the traveller settles here
                if d > ROBIN_DIST_LIMIT:
                    self.grow_next = True
                slot[b] = [cur, h, d]
                where[cur] = b
                cur, h, d = r
            b = (b + 1) & mask
            d += 1

    def _grow(self):
        old = [self.slot[b] for b in sorted(self.slot)]
        self.bc = 2 if self.bc == 0 else self.bc * 2
        self.slot, self.where = {}, {}
        self.grow_next = False
        for cur, h, _ in old:           # rehash_impl: re-insert in the old table's bucket order
            self._place(cur, h, h & (self.bc - 1), 0)

    def insert_new(self, ident, h):
        """insert() of a key known to be absent."""
        h = int(h)
        if self.grow_next or len(self.slot) >= self.bc // 2:
            self._grow()
        self._place(ident, h, h & (self.bc - 1), 0)

    def erase(self, ident):
        """erase(key): backward-shift deletion.  Returns True if an entry slid into the erased bucket."""
        mask = self.bc - 1
        b = self.where.pop(ident)
        del self.slot[b]
        shifted = False
        nb = (b + 1) & mask
        while True:
            r = self.slot.get(nb)
            if r is None or r[2] == 0:
                return shifted
            del self.slot[nb]
            r[2] -= 1
            self.slot[b] = r
            self.where[r[0]] = b
            shifted = True
            b, nb = nb, (nb + 1) & mask

    def iteration(self):
        """ids in begin()..end() order (bucket order)."""
        return [self.slot[b][0] for b in sorted(self.slot)]


def robin_layout(hashes, bucket_count):
    """Entries (given in insertion order, distinct keys) of a robin_map with `bucket_count` buckets that never
    grows: (ids in iteration order, their buckets)."""
    t = RobinTable()
    t.bc = int(bucket_count)
    for e, h in enumerate(np.asarray(hashes).tolist()):
        t._place(e, h, h & (t.bc - 1), 0)
    bs = sorted(t.slot)
    return np.array([t.slot[b][0] for b in bs], dtype=np.int64), np.array(bs, dtype=np.int64)


def voxel_down_sample_idx(frame, voxel_size, order="index"):
    """Indices of the first point of every voxel (A.5): ascending (B.1, order="index") or in the iteration
    order of upstream's `tsl::robin_map grid; grid.reserve(frame.size())` (order="robin_map")."""
    frame = np.asarray(frame, dtype=np.float64).reshape(-1, 3)
    if frame.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    keys = voxel_keys(frame, voxel_size)
    packed = pack_keys(keys)
    _, first = np.unique(packed, return_index=True)
    first = np.sort(first)                 # insertion order of the voxels = order of first appearance
    if order == "index":
        return first
    assert order == "robin_map", order
    it, _ = robin_layout(upstream_voxel_hash(keys[first]), robin_bucket_count_reserve(frame.shape[0]))
    return first[it]


def voxel_down_sample(frame, voxel_size, order="index"):
    frame = np.asarray(frame, dtype=np.float64).reshape(-1, 3)
    return frame[voxel_down_sample_idx(frame, voxel_size, order)]


# ---------------------------------------------------------------------------
# A.6 / A.7 VoxelHashMap
# ---------------------------------------------------------------------------
_OFFSETS = np.array([(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)],
                    dtype=np.int64)  # i outermost, k innermost (A.7)


class VoxelHashMap:
    def __init__(self, voxel_size, max_distance, max_points_per_voxel=20, order="index"):
        self.voxel_size = float(voxel_size)
        self.max_distance = float(max_distance)
        self.max_points = int(max_points_per_voxel)
        self.order = order
        self.clear()

    def clear(self):
        self.keys = np.zeros(0, dtype=np.int64)                    # sorted packed keys
        self.pts = np.zeros((0, self.max_points, 3), dtype=np.float64)
        self.cnt = np.zeros(0, dtype=np.int32)
        self.seq = np.zeros(0, dtype=np.int64)                     # creation sequence number of every voxel
        self._next_seq = 0
        self.table = RobinTable() if self.order == "robin_map" else None   # upstream's map_, ids = packed keys
        self.skipped_last_prune = 0

    def empty(self):
        return self.keys.shape[0] == 0

    def num_voxels(self):
        return int(self.keys.shape[0])

    def _lookup(self, packed):
        pos = np.searchsorted(self.keys, packed)
        posc = np.minimum(pos, max(self.keys.shape[0] - 1, 0))
        found = (self.keys[posc] == packed) if self.keys.shape[0] else np.zeros(packed.shape, bool)
        return posc, found

    def add_points(self, points):
        points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
        n = points.shape[0]
        if n == 0:
            return
        packed = pack_keys(voxel_keys(points, self.voxel_size))
        order = np.argsort(packed, kind="stable")
        sp = packed[order]
        start = np.ones(n, dtype=bool)
        start[1:] = sp[1:] != sp[:-1]
        gstart = np.flatnonzero(start)
        gid = np.cumsum(start) - 1
        rank_sorted = np.arange(n) - gstart[gid]
        rank = np.empty(n, dtype=np.int64)
        rank[order] = rank_sorted
        uniq = sp[gstart]
        # existing / new voxels
        posc, found = self._lookup(uniq)
        new_keys = uniq[~found]
        if new_keys.shape[0]:
            # creation order = order in which the scan's points first touch the new voxels
            first_touch = order[gstart[~found]]
            new_seq = np.empty(new_keys.shape[0], dtype=np.int64)
            new_seq[np.argsort(first_touch, kind="stable")] = self._next_seq + np.arange(new_keys.shape[0])
            self._next_seq += new_keys.shape[0]
            keys = np.concatenate([self.keys, new_keys])
            pts = np.concatenate([self.pts, np.zeros((new_keys.shape[0], self.max_points, 3))])
            cnt = np.concatenate([self.cnt, np.zeros(new_keys.shape[0], dtype=np.int32)])
            seq = np.concatenate([self.seq, new_seq])
            o = np.argsort(keys, kind="stable")
            self.keys, self.pts, self.cnt, self.seq = keys[o], pts[o], cnt[o], seq[o]
            if self.table is not None:                               # map_.insert in the order the voxels are created
                created = new_keys[np.argsort(new_seq, kind="stable")]
                hs = upstream_voxel_hash(unpack_keys(created))
                for kk, hh in zip(created.tolist(), hs.tolist()):
                    self.table.insert_new(kk, hh)
        vid, f = self._lookup(packed)
        assert f.all()
        slot = self.cnt[vid].astype(np.int64) + rank
        keep = slot < self.max_points
        self.pts[vid[keep], slot[keep]] = points[keep]
        np.add.at(self.cnt, vid[keep], 1)

    def remove_far_away_points(self, origin):
        if self.empty():
            return
        o = np.asarray(origin, dtype=np.float64)
        dx = self.pts[:, 0, 0] - o[0]
        dy = self.pts[:, 0, 1] - o[1]
        dz = self.pts[:, 0, 2] - o[2]
        d2 = (dx * dx + dy * dy) + dz * dz
        far = d2 > self.max_distance * self.max_distance
        self.skipped_last_prune = 0
        if self.order == "robin_map" and far.any():
            far = self._robin_erase_while_iterating(far)
        keep = ~far
        self.keys, self.pts, self.cnt, self.seq = self.keys[keep], self.pts[keep], self.cnt[keep], self.seq[keep]

    def _robin_erase_while_iterating(self, far):
        """Which voxels upstream 0.2.x really erases: `for (auto& [voxel, block] : map_) if (far) map_.erase(voxel);`
        over a tsl::robin_map.  erase() shifts the rest of the cluster back by one bucket, so the entry that
        slides into the erased bucket is stepped over by the ++ of the range-for and survives this call (A.6)."""
        t = self.table
        far_of = dict(zip(self.keys.tolist(), far.tolist()))
        erased_keys = set()
        b, last = 0, t.bc
        while b < last:                                 # the iterator is a bucket pointer
            r = t.slot.get(b)
            if r is not None and far_of[r[0]]:
                kk = r[0]
                slid = t.erase(kk)
                erased_keys.add(kk)
                if slid:
                    nxt = t.slot.get(b)
                    if nxt is not None and far_of[nxt[0]]:
                        self.skipped_last_prune += 1
            b += 1
        return np.array([k in erased_keys for k in self.keys.tolist()], dtype=bool)

    def update(self, points, pose):
        """VoxelHashMap::Update(points, pose) (/root/reference/src/ptudes/kiss.py:129)."""
        points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
        pose = np.asarray(pose, dtype=np.float64)
        x, y, z = canon.transform_points(pose, points[:, 0], points[:, 1], points[:, 2])
        self.add_points(np.stack([x, y, z], axis=1))
        self.remove_far_away_points(pose[:3, 3])

    def point_cloud(self):
        if self.empty():
            return np.zeros((0, 3))
        mask = np.arange(self.max_points)[None, :] < self.cnt[:, None]
        return self.pts[mask]

    def voxel_table(self):
        """(keys (V,3) int32 sorted by packed key, counts (V,), points (V,maxp,3))."""
        return unpack_keys(self.keys), self.cnt.copy(), self.pts.copy()

    def nearest(self, points):
        """For every query: (found, nearest xyz, d2, order id = offset*maxp + slot)."""
        q = np.asarray(points, dtype=np.float64).reshape(-1, 3)
        m = q.shape[0]
        found_any = np.zeros(m, dtype=bool)
        best = np.zeros((m, 3))
        best_d2 = np.full(m, np.inf)
        best_ord = np.full(m, -1, dtype=np.int64)
        if m == 0 or self.empty():
            return found_any, best, best_d2, best_ord
        k = voxel_keys(q, self.voxel_size).astype(np.int64)
        CH = 4096
        for s in range(0, m, CH):
            e = min(m, s + CH)
            nk = k[s:e, None, :] + _OFFSETS[None, :, :]                     # (c,27,3)
            if np.abs(nk).max() >= KEY_BIAS:
                raise ValueError("voxel coordinate out of range")
            packed = ((nk[..., 0] + KEY_BIAS) << 42) | ((nk[..., 1] + KEY_BIAS) << 21) | (nk[..., 2] + KEY_BIAS)
            pos, fnd = self._lookup(packed.reshape(-1))
            pos = pos.reshape(e - s, 27)
            fnd = fnd.reshape(e - s, 27)
            cand = self.pts[pos]                                            # (c,27,P,3)
            cnt = np.where(fnd, self.cnt[pos], 0)
            dx = cand[..., 0] - q[s:e, None, None, 0]
            dy = cand[..., 1] - q[s:e, None, None, 1]
            dz = cand[..., 2] - q[s:e, None, None, 2]
            d2 = (dx * dx + dy * dy) + dz * dz
            valid = np.arange(self.max_points)[None, None, :] < cnt[..., None]
            d2 = np.where(valid, d2, np.inf).reshape(e - s, -1)
            arg = np.argmin(d2, axis=1)                                     # first minimum = B.6
            dmin = d2[np.arange(e - s), arg]
            ok = np.isfinite(dmin)
            found_any[s:e] = ok
            best_d2[s:e] = dmin
            best_ord[s:e] = np.where(ok, arg, -1)
            best[s:e] = cand.reshape(e - s, -1, 3)[np.arange(e - s), arg]
        return found_any, best, best_d2, best_ord

    def get_correspondences(self, points, max_correspondance_distance, return_index=False):
        """VoxelHashMap::GetCorrespondences (A.7): (source, target) in query order."""
        q = np.asarray(points, dtype=np.float64).reshape(-1, 3)
        found, best, d2, order = self.nearest(q)
        with np.errstate(invalid="ignore"):
            acc = found & (np.sqrt(d2) < max_correspondance_distance)
        if return_index:
            return acc, best, order
        return q[acc], best[acc]


# ---------------------------------------------------------------------------
# A.8 registration
# ---------------------------------------------------------------------------
def linear_system_terms(src, tgt, acc, kernel):
    """Per-point 27-vector: 21 upper-triangular JtJ entries (row-major) + 6 Jtr; rows of
    rejected points are zero.  J = [I | -hat(s)], w = k^2/(k+|r|^2)^2 (A.8)."""
    sx, sy, sz = src[:, 0], src[:, 1], src[:, 2]
    rx, ry, rz = sx - tgt[:, 0], sy - tgt[:, 1], sz - tgt[:, 2]
    r2 = (rx * rx + ry * ry) + rz * rz
    kk = kernel + r2
    w = (kernel * kernel) / (kk * kk)
    z = np.zeros_like(w)
    wsx, wsy, wsz = w * sx, w * sy, w * sz
    wrx, wry, wrz = w * rx, w * ry, w * rz
    cols = [
        w, z, z, z, wsz, -wsy,                  # row 0: (0,0) .. (0,5)
        w, z, -wsz, z, wsx,                     # row 1: (1,1) .. (1,5)
        w, wsy, -wsx, z,                        # row 2: (2,2) .. (2,5)
        w * (sy * sy + sz * sz), -(w * (sx * sy)), -(w * (sx * sz)),   # row 3
        w * (sx * sx + sz * sz), -(w * (sy * sz)),                     # row 4
        w * (sx * sx + sy * sy),                                       # row 5
        wrx, wry, wrz,
        sy * wrz - sz * wry, sz * wrx - sx * wrz, sx * wry - sy * wrx,
    ]
    terms = np.stack(cols, axis=1)
    terms[~acc] = 0.0
    return terms


def unpack_system(sums):
    A = np.zeros((6, 6))
    idx = 0
    for i in range(6):
        for j in range(i, 6):
            A[i, j] = sums[idx]
            A[j, i] = sums[idx]
            idx += 1
    return A, np.array(sums[21:27], dtype=np.float64)


def register_point_cloud(points, voxel_map, initial_guess, max_correspondance_distance, kernel,
                         trace=None, max_iters=MAX_ITERS):
    """kiss-icp RegisterFrame (A.8); call site /root/reference/src/ptudes/kiss.py:108-114.
    Returns (4x4 pose, stats dict)."""
    guess = np.array(initial_guess, dtype=np.float64)
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    stats = {"iterations": 0, "n_corr": 0, "dx_norm": 0.0, "status": 0}
    # the pybind boundary turns the 4x4 guess into a Sophus::SE3d (quaternion + translation);
    # whatever is returned goes back through .matrix(), i.e. is re-orthonormalised
    guess_q = canon.SE3q.from_matrix(guess)
    T_icp = canon.SE3q()
    if voxel_map.empty():
        return T_icp.mul(guess_q).matrix(), stats
    x, y, z = canon.transform_points(guess, pts[:, 0], pts[:, 1], pts[:, 2])
    src = np.stack([x, y, z], axis=1)
    for it in range(max_iters):
        acc, tgt, order = voxel_map.get_correspondences(src, max_correspondance_distance, return_index=True)
        n_corr = int(acc.sum())
        stats["iterations"] = it + 1
        stats["n_corr"] = n_corr
        if trace is not None:
            trace.append({"acc": acc.copy(), "order": np.where(acc, order, -1), "src": src.copy()})
        if n_corr == 0:                       # B.5
            stats["status"] = 1
            break
        sums = canon.pairwise_tree_sum(linear_system_terms(src, tgt, acc, kernel))
        A, b = unpack_system(sums)
        dx, ok = canon.ldlt_solve6(A, [-v for v in b])
        if not ok:
            stats["status"] = 2
            break
        Eq, E = canon.se3_exp_q(np.array(dx))
        x, y, z = canon.transform_points(E, src[:, 0], src[:, 1], src[:, 2])
        src = np.stack([x, y, z], axis=1)
        T_icp = Eq.mul(T_icp)
        nrm = math.sqrt(((((dx[0] * dx[0] + dx[1] * dx[1]) + dx[2] * dx[2]) + dx[3] * dx[3])
                         + dx[4] * dx[4]) + dx[5] * dx[5])
        stats["dx_norm"] = nrm
        if nrm < EST_THRESHOLD:
            break
    return T_icp.mul(guess_q).matrix(), stats


# ---------------------------------------------------------------------------
# A.9 adaptive threshold
# ---------------------------------------------------------------------------
class AdaptiveThreshold:
    def __init__(self, initial_threshold, min_motion_th, max_range):
        self.initial_threshold = float(initial_threshold)
        self.min_motion_th = float(min_motion_th)
        self.max_range = float(max_range)
        self.model_error_sse2 = 0.0
        self.num_samples = 0
        self.model_deviation = np.eye(4)

    def update_model_deviation(self, T):
        self.model_deviation = np.array(T, dtype=np.float64)

    def compute_threshold(self):
        d = self.model_deviation
        theta = canon.rot_angle(d[:3, :3])
        delta_rot = 2.0 * self.max_range * math.sin(theta / 2.0)
        tx, ty, tz = float(d[0, 3]), float(d[1, 3]), float(d[2, 3])
        delta_trans = math.sqrt((tx * tx + ty * ty) + tz * tz)
        err = delta_trans + delta_rot
        if err > self.min_motion_th:
            self.model_error_sse2 += err * err
            self.num_samples += 1
        if self.num_samples < 1:
            return self.initial_threshold
        return math.sqrt(self.model_error_sse2 / self.num_samples)


# ---------------------------------------------------------------------------
# config + KissICP object + the ptudes wrapper
# ---------------------------------------------------------------------------
class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def load_config(config_file=None, deskew=False, max_range=100.0):
    """kiss_icp.config.load_config defaults (A.1): voxel_size frozen = max_range/100."""
    cfg = _NS(
        data=_NS(preprocess=True, correct_scan=True, max_range=float(max_range), min_range=5.0,
                 deskew=bool(deskew)),
        mapping=_NS(voxel_size=float(max_range) / 100.0, max_points_per_voxel=20),
        adaptive_threshold=_NS(fixed_threshold=None, initial_threshold=2.0, min_motion_th=0.1),
    )
    return cfg


class _Compensator:
    def deskew_scan(self, frame, poses, timestamps):
        if len(poses) < 2:
            return frame
        return deskew_scan(frame, timestamps, poses[-2], poses[-1])


class KissICP:
    """State holder with the members /root/reference/src/ptudes/kiss.py reads off
    kiss_icp.kiss_icp.KissICP: poses, config, compensator, preprocess, voxelize,
    get_adaptive_threshold, get_prediction_model, adaptive_threshold, local_map."""

    def __init__(self, config, order="index"):
        self.poses = []
        self.order = order
        self.config = config
        self.compensator = _Compensator()
        self.adaptive_threshold = AdaptiveThreshold(config.adaptive_threshold.initial_threshold,
                                                    config.adaptive_threshold.min_motion_th,
                                                    config.data.max_range)
        self.local_map = VoxelHashMap(config.mapping.voxel_size, config.data.max_range,
                                      config.mapping.max_points_per_voxel, order=order)

    def preprocess(self, frame):
        return preprocess(frame, self.config.data.max_range, self.config.data.min_range)

    def voxelize(self, frame):
        v = self.config.mapping.voxel_size
        frame_downsample = voxel_down_sample(frame, v * 0.5, self.order)
        source = voxel_down_sample(frame_downsample, v * 1.5, self.order)
        return source, frame_downsample

    def has_moved(self):
        if len(self.poses) < 1:
            return False
        d = canon.rigid_mul(canon.rigid_inv(self.poses[0]), self.poses[-1])
        tx, ty, tz = float(d[0, 3]), float(d[1, 3]), float(d[2, 3])
        motion = math.sqrt((tx * tx + ty * ty) + tz * tz)
        return motion > 5.0 * self.config.adaptive_threshold.min_motion_th

    def get_adaptive_threshold(self):
        if not self.has_moved():
            return self.config.adaptive_threshold.initial_threshold
        return self.adaptive_threshold.compute_threshold()

    def get_prediction_model(self):
        if len(self.poses) < 2:
            return np.eye(4)
        return canon.rigid_mul(canon.rigid_inv(self.poses[-2]), self.poses[-1])


class OracleKissICPWrapper:
    """The ouster-free part of /root/reference/src/ptudes/kiss.py:18-166 on the oracle."""

    def __init__(self, *, _min_range=5, _max_range=100, order="index"):
        self._max_range = _max_range
        self._min_range = _min_range
        self._kiss_config = load_config(None, deskew=True, max_range=self._max_range)
        self._kiss_config.data.min_range = self._min_range          # kiss.py:43
        self._kiss = KissICP(config=self._kiss_config, order=order)
        self._poses_ts = []
        self._err_dt = []
        self._err_drot = []
        self._sigmas = []
        self.last_stats = None
        self.last_counts = None

    def deskew(self, frame, timestamps):
        return self._kiss.compensator.deskew_scan(frame, self._kiss.poses, timestamps)

    def _kiss_register_frame(self, frame, timestamps, ts, initial_guess=None, trace=None):
        kself = self._kiss
        n_in = len(frame)
        frame = kself.compensator.deskew_scan(frame, self.poses, timestamps)   # kiss.py:90
        frame = kself.preprocess(frame)                                       # kiss.py:93
        source, frame_downsample = kself.voxelize(frame)                      # kiss.py:96
        sigma = kself.get_adaptive_threshold()                                # kiss.py:99
        if initial_guess is None:                                             # kiss.py:102-105
            prediction = kself.get_prediction_model()
            last_pose = kself.poses[-1] if kself.poses else np.eye(4)
            initial_guess = canon.rigid_mul(last_pose, prediction)
        initial_guess = np.array(initial_guess, dtype=np.float64)
        new_pose, stats = register_point_cloud(source, kself.local_map, initial_guess,
                                               3 * sigma, sigma / 3, trace=trace)  # kiss.py:108-114
        pose_gain = canon.rigid_mul(canon.rigid_inv(initial_guess), new_pose)  # kiss.py:116
        tx, ty, tz = float(pose_gain[0, 3]), float(pose_gain[1, 3]), float(pose_gain[2, 3])
        dt = math.sqrt((tx * tx + ty * ty) + tz * tz)
        drot = abs(canon.so3_log(pose_gain[:3, :3])[1])
        self._err_dt.append(dt)
        self._err_drot.append(drot)
        self._sigmas.append(sigma)
        kself.adaptive_threshold.update_model_deviation(pose_gain)           # kiss.py:128
        kself.local_map.update(frame_downsample, new_pose)                    # kiss.py:129
        kself.poses.append(new_pose)                                          # kiss.py:130
        self.last_stats = stats
        self.last_counts = {"n": n_in, "n_range": len(frame), "n_ds": len(frame_downsample),
                            "n_src": len(source), "n_vox": kself.local_map.num_voxels()}
        return frame, source

    def register_points(self, frame, timestamps, ts, initial_guess=None, trace=None):
        """register_frame (kiss.py:54-74) minus the ouster LidarScan -> xyz projection."""
        self._kiss_register_frame(frame, timestamps, ts, initial_guess=initial_guess, trace=trace)
        self._poses_ts.append(ts)                                             # kiss.py:72
        return self.pose

    @property
    def velocity(self):
        if len(self.poses) < 2:
            return np.zeros(3)
        prediction = self._kiss.get_prediction_model()
        dt = self.poses_ts[-1] - self.poses_ts[-2]
        return prediction[:3, 3] / dt

    @property
    def pose(self):
        if not self.poses:
            return np.eye(4)
        return self.poses[-1]

    @property
    def poses(self):
        return self._kiss.poses

    @property
    def poses_ts(self):
        return self._poses_ts

    @property
    def local_map_points(self):
        return self._kiss.local_map.point_cloud()

    @property
    def _config(self):
        return self._kiss.config
