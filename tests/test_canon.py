"""CPU tests of the canonical arithmetic (oracle/canon.py)."""
import math

import numpy as np

from oracle import canon


def _ulp_diff(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.spacing(np.maximum(np.abs(b), 1e-300))


def test_det_sincos_close_to_libm():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-4.0, 4.0, 200000), rng.uniform(-1e-3, 1e-3, 10000),
                        rng.uniform(-100, 100, 20000), [0.0, math.pi / 4, math.pi / 2, math.pi, -math.pi]])
    s, c = canon.det_sincos(x)
    rs = np.array([math.sin(v) for v in x])
    rc = np.array([math.cos(v) for v in x])
    # within 2 ulp of glibc wherever the result is not tiny (near zeros of sin/cos the absolute
    # error is what matters: < 2e-16)
    big_s = np.abs(rs) > 1e-3
    big_c = np.abs(rc) > 1e-3
    assert _ulp_diff(s[big_s], rs[big_s]).max() <= 2.0
    assert _ulp_diff(c[big_c], rc[big_c]).max() <= 2.0
    assert np.abs(s - rs).max() < 2.5e-16 and np.abs(c - rc).max() < 2.5e-16
    assert canon.det_sincos_scalar(0.0) == (0.0, 1.0)


def test_se3_exp_log_roundtrip():
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = np.concatenate([rng.uniform(-5, 5, 3), rng.uniform(-1.5, 1.5, 3)])
        T = canon.se3_exp_mat(a)
        assert np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3)).max() < 1e-14
        b = canon.se3_log(T)
        assert np.abs(a - b).max() < 1e-12
        assert np.abs(canon.se3_exp_mat(b) - T).max() < 1e-13
    # pure translation, and the small-angle branches
    T = canon.se3_exp_mat(np.array([1.0, 2.0, 3.0, 0, 0, 0]))
    assert np.array_equal(T[:3, :3], np.eye(3)) and T[:3, 3].tolist() == [1.0, 2.0, 3.0]
    T = canon.se3_exp_mat(np.array([1.0, 2.0, 3.0, 1e-12, 0, 0]))
    assert np.abs(T[:3, 3] - [1, 2, 3]).max() < 1e-11
    assert np.abs(canon.se3_log(np.eye(4))).max() == 0.0


def test_se3_exp_matches_rodrigues():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(2)
    w = rng.uniform(-2, 2, (100, 3))
    R, _ = canon.se3_exp(np.concatenate([np.zeros((100, 3)), w], axis=1))
    assert np.abs(R - Rotation.from_rotvec(w).as_matrix()).max() < 1e-14


def test_rigid_helpers_match_numpy():
    rng = np.random.default_rng(3)
    A = canon.se3_exp_mat(rng.uniform(-1, 1, 6))
    B = canon.se3_exp_mat(rng.uniform(-1, 1, 6))
    assert np.abs(canon.rigid_mul(A, B) - A @ B).max() < 1e-15
    assert np.abs(canon.rigid_inv(A) - np.linalg.inv(A)).max() < 1e-14
    assert abs(canon.rot_angle(A[:3, :3]) - np.linalg.norm(canon.se3_log(A)[3:])) < 1e-14


def test_pairwise_tree_sum():
    rng = np.random.default_rng(4)
    for n in (0, 1, 5, 32, 33, 100, 1000):
        a = rng.normal(size=(n, 27))
        s = canon.pairwise_tree_sum(a)
        assert np.abs(s - a.sum(axis=0)).max() < 1e-12 if n else np.all(s == 0)
    # explicit shape: ((a0+a1)+(a2+a3)) for 4 values padded to 32
    a = np.array([[1e16], [1.0], [-1e16], [1.0]])
    assert canon.pairwise_tree_sum(a)[0] == (1e16 + 1.0) + (-1e16 + 1.0)


def test_ldlt_solve6():
    rng = np.random.default_rng(5)
    for _ in range(50):
        J = rng.normal(size=(40, 6))
        A = J.T @ J
        b = rng.normal(size=6)
        x, ok = canon.ldlt_solve6(A, b)
        assert ok
        assert np.abs(np.array(x) - np.linalg.solve(A, b)).max() < 1e-9
    x, ok = canon.ldlt_solve6(np.zeros((6, 6)), np.ones(6))
    assert not ok
