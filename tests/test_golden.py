"""Committed golden vectors (tests/golden/tiny_seq.npz, made by tests/golden/make_golden.py) against
all three implementations: the NumPy oracle, the C port (both CPU, run everywhere) and the CUDA
path through the C ABI (gpu).  Everything is compared for exact equality."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import TRACE_ITERS, table_checksum  # noqa: E402

from ptudes_lab_b200 import synth  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_seq.npz")
RUNS = ["A", "B", "C"]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _inputs(gold, k):
    return synth.project_scan(gold["ranges"][k], gold["dirs"])


def _guess(gold, tag, k):
    g = gold[f"{tag}_guesses"][k]
    return None if np.isnan(g[0, 0]) else g


def _check_scan(gold, tag, k, pose, sigma, err_dt, err_drot, iterations, counts):
    assert np.array_equal(pose, gold[f"{tag}_poses"][k]), (tag, k)
    assert sigma == gold[f"{tag}_sigma"][k] and err_dt == gold[f"{tag}_err_dt"][k], (tag, k)
    assert err_drot == gold[f"{tag}_err_drot"][k], (tag, k)
    assert iterations == gold[f"{tag}_iterations"][k], (tag, k)
    assert list(counts) == list(gold[f"{tag}_counts"][k]), (tag, k)


def _slice(gold, tag, name, k):
    off = gold[f"{tag}_{name[:-4]}_off"] if name.endswith("_idx") else None
    return gold[f"{tag}_{name}"][off[k]:off[k + 1]]


def _check_map(gold, tag, keys, cnt, pts):
    assert np.array_equal(keys, gold[f"{tag}_map_keys"])
    assert np.array_equal(cnt, gold[f"{tag}_map_counts"])
    assert np.array_equal(table_checksum(pts), gold[f"{tag}_map_checksum"])


@pytest.mark.parametrize("tag", RUNS)
def test_numpy_oracle_matches_golden(gold, tag):
    from oracle import kiss_oracle as ko
    mn, mx, n = gold[f"{tag}_cfg"]
    ref = ko.OracleKissICPWrapper(_min_range=mn, _max_range=mx)
    for k in range(int(n)):
        xyz, ts = _inputs(gold, k)
        trace = []
        ref.register_points(xyz, ts, 0.1 * (k + 1), initial_guess=_guess(gold, tag, k), trace=trace)
        c = ref.last_counts
        _check_scan(gold, tag, k, ref.pose, ref._sigmas[-1], ref._err_dt[-1], ref._err_drot[-1],
                    ref.last_stats["iterations"], [c["n"], c["n_range"], c["n_ds"], c["n_src"], c["n_vox"]])
        so = gold[f"{tag}_src_off"]
        tr = gold[f"{tag}_trace"][:, so[k]:so[k + 1]]
        for it in range(min(TRACE_ITERS, len(trace))):
            assert np.array_equal(trace[it]["order"], tr[it])
    _check_map(gold, tag, *ref._kiss.local_map.voxel_table())


@pytest.mark.parametrize("tag", RUNS)
def test_c_port_matches_golden(gold, tag):
    from oracle import port
    mn, mx, n = gold[f"{tag}_cfg"]
    p = port.PortKissICP(_min_range=mn, _max_range=mx, threads=2, trace_iterations=TRACE_ITERS)
    for k in range(int(n)):
        xyz, ts = _inputs(gold, k)
        p.register_points(xyz, ts, 0.1 * (k + 1), initial_guess=_guess(gold, tag, k))
        c = p.last_counts
        _check_scan(gold, tag, k, p.pose, p._sigmas[-1], p._err_dt[-1], p._err_drot[-1], p.last_stats["iterations"],
                    [c["n"], c["n_range"], c["n_ds"], c["n_src"], c["n_vox"]])
        _, i2 = p.get_points(1, with_index=True)
        assert np.array_equal(i2, _slice(gold, tag, "src_idx", k))
        so = gold[f"{tag}_src_off"]
        tr = p.get_trace()
        assert np.array_equal(tr, gold[f"{tag}_trace"][:tr.shape[0], so[k]:so[k + 1]])
    _check_map(gold, tag, *p.voxel_table())
    p.close()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", RUNS)
def test_cuda_path_matches_golden(gold, tag):
    from ptudes_lab_b200 import odometry
    mn, mx, n = gold[f"{tag}_cfg"]
    cfg = odometry.load_config(None, deskew=True, max_range=float(mx))
    cfg.data.min_range = float(mn)
    o = odometry.Odometry(cfg, max_points=16384, map_capacity=32768, trace_iterations=TRACE_ITERS)
    try:
        for k in range(int(n)):
            xyz, ts = _inputs(gold, k)
            pose, st = o.register_frame(xyz, ts, initial_guess=_guess(gold, tag, k))
            _check_scan(gold, tag, k, pose, st["sigma"], st["err_dt"], st["err_drot"], st["iterations"],
                        [st["n_in"], st["n_range"], st["n_ds"], st["n_src"], st["n_voxels"]])
            _, i1 = o.get_points(0, with_index=True)
            _, i2 = o.get_points(1, with_index=True)
            assert np.array_equal(i1, _slice(gold, tag, "ds_idx", k))
            assert np.array_equal(i2, _slice(gold, tag, "src_idx", k))
            so = gold[f"{tag}_src_off"]
            tr = o.get_trace()
            assert np.array_equal(tr, gold[f"{tag}_trace"][:tr.shape[0], so[k]:so[k + 1]])
        _check_map(gold, tag, *odometry.VoxelHashMap(o, 0).dump())
    finally:
        o.close()
