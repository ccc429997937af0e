"""Generate tests/golden/tiny_seq.npz - committed input/output vectors of the odometry step.

The reference cannot be imported or built in this environment (kiss-icp, ouster-sdk absent; the
arithmetic is not under /root/reference), so these are NOT outputs of the reference: they are
outputs of the NumPy oracle (oracle/kiss_oracle.py), written only after the independently
written C port (oracle/kiss_port.c) reproduced every one of them bit for bit.  Their job is to
pin both oracles and the CUDA path to one fixed set of numbers, so that a change in any of the
three shows up as a diff against a file in git (parity itself stays "unpinned" w.r.t. upstream
kiss-icp).

Inputs are stored too (beam directions + range images in integer millimetres), so nothing depends
on how a given NumPy build rounds sin/cos inside the synthetic generator.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import canon, kiss_oracle as ko, port  # noqa: E402
from ptudes_lab_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_seq.npz")
N_A, N_B, N_C = 10, 6, 8
TRACE_ITERS = 4


def table_checksum(pts):
    bits = np.ascontiguousarray(pts, dtype=np.float64).view(np.uint64).reshape(-1)
    w = (np.arange(bits.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return np.array([np.bitwise_xor.reduce(bits * w), (bits * w).sum(dtype=np.uint64)], dtype=np.uint64)


def guess_for(k, poses, rng):
    """injected guesses of run C: perturbed constant-velocity prediction from scan 2 on"""
    if k < 2:
        return None
    pred = canon.rigid_mul(canon.rigid_inv(poses[-2]), poses[-1])
    g = canon.rigid_mul(poses[-1], pred)
    return canon.rigid_mul(g, canon.se3_exp_mat(rng.normal(0, 0.01, 6)))


def run(dirs, ranges, min_r, max_r, n, with_guess=False):
    ref = ko.OracleKissICPWrapper(_min_range=min_r, _max_range=max_r)
    prt = port.PortKissICP(_min_range=min_r, _max_range=max_r, threads=2, trace_iterations=TRACE_ITERS)
    rng = np.random.default_rng(11)
    out = {k: [] for k in ("poses", "sigma", "err_dt", "err_drot", "iterations", "n_corr", "counts", "guesses",
                           "ds_idx", "src_idx", "trace")}
    for k in range(n):
        xyz, ts = synth.project_scan(ranges[k], dirs)
        g = guess_for(k, ref.poses, rng) if with_guess else None
        trace = []
        ref.register_points(xyz, ts, 0.1 * (k + 1), initial_guess=g, trace=trace)
        prt.register_points(xyz, ts, 0.1 * (k + 1), initial_guess=g)
        assert np.array_equal(ref.pose, prt.pose), ("port != oracle", k)
        c = ref.last_counts
        assert all(c[q] == prt.last_counts[q] for q in ("n", "n_range", "n_ds", "n_src", "n_vox"))
        assert ref.last_stats["iterations"] == prt.last_stats["iterations"]
        # downsample selections as indices into the input scan
        fr = ref._kiss.compensator.deskew_scan(xyz, ref.poses[:-1], ts)
        mask = ko.range_mask(fr, max_r, min_r)
        v = ref._config.mapping.voxel_size
        i1 = np.flatnonzero(mask)[ko.voxel_down_sample_idx(fr[mask], v * 0.5)]
        i2 = ko.voxel_down_sample_idx(fr[i1], v * 1.5)
        _, p1 = prt.get_points(0, with_index=True)
        _, p2 = prt.get_points(1, with_index=True)
        assert np.array_equal(np.flatnonzero(mask)[p1], i1) and np.array_equal(p2, i2)
        tr = np.full((TRACE_ITERS, len(i2)), -2, dtype=np.int32)
        for it in range(min(TRACE_ITERS, len(trace))):
            tr[it] = trace[it]["order"]
        pt = prt.get_trace()
        assert np.array_equal(pt, tr[:pt.shape[0]])
        out["poses"].append(ref.pose.copy())
        out["sigma"].append(ref._sigmas[-1])
        out["err_dt"].append(ref._err_dt[-1])
        out["err_drot"].append(ref._err_drot[-1])
        out["iterations"].append(ref.last_stats["iterations"])
        out["n_corr"].append(ref.last_stats["n_corr"])
        out["counts"].append([c["n"], c["n_range"], c["n_ds"], c["n_src"], c["n_vox"]])
        out["guesses"].append(np.full((4, 4), np.nan) if g is None else g)
        out["ds_idx"].append(i1.astype(np.int32))
        out["src_idx"].append(i2.astype(np.int32))
        out["trace"].append(tr)
    keys, cnt, pts = ref._kiss.local_map.voxel_table()
    k2, c2, p2 = prt.voxel_table()
    assert np.array_equal(keys, k2) and np.array_equal(cnt, c2) and np.array_equal(pts, p2)
    res = {
        "poses": np.stack(out["poses"]), "sigma": np.array(out["sigma"]), "err_dt": np.array(out["err_dt"]),
        "err_drot": np.array(out["err_drot"]), "iterations": np.array(out["iterations"], dtype=np.int32),
        "n_corr": np.array(out["n_corr"], dtype=np.int32), "counts": np.array(out["counts"], dtype=np.int32),
        "guesses": np.stack(out["guesses"]),
        "ds_idx": np.concatenate(out["ds_idx"]), "ds_off": np.cumsum([0] + [len(a) for a in out["ds_idx"]]).astype(np.int32),
        "src_idx": np.concatenate(out["src_idx"]), "src_off": np.cumsum([0] + [len(a) for a in out["src_idx"]]).astype(np.int32),
        "trace": np.concatenate(out["trace"], axis=1),
        "map_keys": keys, "map_counts": cnt,
        # position-weighted checksum of the bit patterns of the (V,20,3) table sorted by key:
        # pins every stored point AND its slot without storing 20 points per voxel
        "map_checksum": table_checksum(pts),
    }
    prt.close()
    return res


def main():
    seq = synth.make_sequence("tiny", 0)
    n = max(N_A, N_B, N_C)
    ranges = np.stack([seq.scan(k).range_mm for k in range(n)])
    data = {"dirs": seq.dirs, "ranges": ranges}
    for tag, (mn, mx, cnt, wg) in {"A": (5.0, 100.0, N_A, False), "B": (1.0, 70.0, N_B, False),
                                   "C": (5.0, 100.0, N_C, True)}.items():
        for k, v in run(seq.dirs, ranges, mn, mx, cnt, wg).items():
            data[f"{tag}_{k}"] = v
        data[f"{tag}_cfg"] = np.array([mn, mx, cnt])
    np.savez_compressed(OUT, **data)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
