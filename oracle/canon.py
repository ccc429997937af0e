"""Canonical float64 arithmetic shared by the oracle and (restated in CUDA) by the product.

TEST INFRASTRUCTURE ONLY.  Nothing outside tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.

PARITY UNPINNED: the arithmetic of the hot path lives in the third-party package
kiss-icp 0.2.x (pinned by /root/reference/setup.py:22; call sites
/root/reference/src/ptudes/kiss.py:90,93,96,99,108-114,128,129), which is not
under /root/reference and not installable here.  The formulas below restate the
published Sophus / Eigen / kiss-icp 0.2.10 algorithms; where
upstream leaves the order of floating-point operations to Eigen expression templates
or to libm, this file FIXES an order (one IEEE-754 double operation per written
operator, no fused multiply-add, left-to-right as parenthesised) so that three
independent implementations - this NumPy one, oracle/kiss_port.c and the CUDA kernels -
agree bit for bit.

sin/cos: the per-point deskew and the per-iteration SE3 exp need sin/cos on the
device.  CUDA's libdevice and glibc differ in the last ulp, so the canon uses its own
`det_sincos` (Cody-Waite reduction + Taylor kernels evaluated in plain double
arithmetic); it is within 2 ulp of libm (tests/test_canon.py) - a documented deviation
far below every tolerance of the path.  Once-per-scan host-side scalars (SE3 log for the
deskew twist, the adaptive-threshold angle) use libm through Python's `math`, the same
glibc the C port and the product's host code call.
"""
import math

import numpy as np

EPS = 1e-10  # Sophus Constants<double>::epsilon()

# ---------------------------------------------------------------------------
# det_sincos
# ---------------------------------------------------------------------------
TWO_OVER_PI = 6.36619772367581382433e-01
# pi/2 split in three parts, the first two with 33 significant bits so that k*PIO2_x is
# exact for |k| < 2^20 (the classic Cody-Waite split of the bits of pi/2).
PIO2_1 = 1.57079632673412561417e+00
PIO2_2 = 6.07710050630396597660e-11
PIO2_3 = 2.02226624879595063154e-21

_S = [-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0,
      1.0 / 6227020800.0, -1.0 / 1307674368000.0, 1.0 / 355687428096000.0]
_C = [1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0,
      -1.0 / 87178291200.0, 1.0 / 20922789888000.0, -1.0 / 6402373705728000.0]


def det_sincos(x):
    """(sin x, cos x) for a float or ndarray, |x| < ~1e5; canonical op order."""
    x = np.asarray(x, dtype=np.float64)
    k = np.rint(x * TWO_OVER_PI)
    r = ((x - k * PIO2_1) - k * PIO2_2) - k * PIO2_3
    z = r * r
    ps = _S[7]
    for c in (_S[6], _S[5], _S[4], _S[3], _S[2], _S[1], _S[0]):
        ps = ps * z + c
    s = r + (r * z) * ps
    pc = _C[7]
    for c in (_C[6], _C[5], _C[4], _C[3], _C[2], _C[1], _C[0]):
        pc = pc * z + c
    c_ = 1.0 - (0.5 * z - (z * z) * pc)
    q = k.astype(np.int64) & 3
    sin = np.where(q == 0, s, np.where(q == 1, c_, np.where(q == 2, -s, -c_)))
    cos = np.where(q == 0, c_, np.where(q == 1, -s, np.where(q == 2, -c_, s)))
    return sin, cos


def det_sincos_scalar(x):
    s, c = det_sincos(np.float64(x))
    return float(s), float(c)


# ---------------------------------------------------------------------------
# small rigid-transform helpers on 4x4 row-major matrices (python floats)
# ---------------------------------------------------------------------------
def mat_identity():
    return np.eye(4)


def rigid_mul(A, B):
    """A @ B for rigid 4x4 (last row assumed 0 0 0 1), canonical order:
    R = Ra Rb with ((a0*b0 + a1*b1) + a2*b2); t = ((a0*tb0 + a1*tb1) + a2*tb2) + ta."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    C = np.eye(4)
    for i in range(3):
        a0, a1, a2 = float(A[i, 0]), float(A[i, 1]), float(A[i, 2])
        for j in range(3):
            C[i, j] = (a0 * float(B[0, j]) + a1 * float(B[1, j])) + a2 * float(B[2, j])
        C[i, 3] = ((a0 * float(B[0, 3]) + a1 * float(B[1, 3])) + a2 * float(B[2, 3])) + float(A[i, 3])
    return C


def rigid_inv(T):
    """Inverse of a rigid 4x4: [R^T | -(R^T t)], canonical order."""
    T = np.asarray(T, dtype=np.float64)
    C = np.eye(4)
    tx, ty, tz = float(T[0, 3]), float(T[1, 3]), float(T[2, 3])
    for i in range(3):
        r0, r1, r2 = float(T[0, i]), float(T[1, i]), float(T[2, i])
        C[i, 0], C[i, 1], C[i, 2] = r0, r1, r2
        C[i, 3] = -((r0 * tx + r1 * ty) + r2 * tz)
    return C


def transform_points(T, x, y, z):
    """p' = R p + t on SoA arrays; ((r0*x + r1*y) + r2*z) + t."""
    T = np.asarray(T, dtype=np.float64)
    xo = ((T[0, 0] * x + T[0, 1] * y) + T[0, 2] * z) + T[0, 3]
    yo = ((T[1, 0] * x + T[1, 1] * y) + T[1, 2] * z) + T[1, 3]
    zo = ((T[2, 0] * x + T[2, 1] * y) + T[2, 2] * z) + T[2, 3]
    return xo, yo, zo


# ---------------------------------------------------------------------------
# SO3 / SE3 exp (Sophus conventions), vectorised over leading dim
# ---------------------------------------------------------------------------
def quat_to_rot(qw, qx, qy, qz):
    """Eigen QuaternionBase::toRotationMatrix, vectorised over leading dims."""
    qw, qx, qy, qz = (np.asarray(v, dtype=np.float64) for v in (qw, qx, qy, qz))
    tx, ty, tz = 2.0 * qx, 2.0 * qy, 2.0 * qz
    twx, twy, twz = tx * qw, ty * qw, tz * qw
    txx, txy, txz = tx * qx, ty * qx, tz * qx
    tyy, tyz, tzz = ty * qy, tz * qy, tz * qz
    R = np.empty(qw.shape + (3, 3))
    R[..., 0, 0] = 1.0 - (tyy + tzz)
    R[..., 0, 1] = txy - twz
    R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz
    R[..., 1, 1] = 1.0 - (txx + tzz)
    R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy
    R[..., 2, 1] = tyz + twx
    R[..., 2, 2] = 1.0 - (txx + tyy)
    return R


def quat_mul(a, b):
    """Eigen quaternion product a*b on (w, x, y, z) python floats, left-to-right sums."""
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return (((aw * bw - ax * bx) - ay * by) - az * bz,
            ((aw * bx + ax * bw) + ay * bz) - az * by,
            ((aw * by + ay * bw) + az * bx) - ax * bz,
            ((aw * bz + az * bw) + ax * by) - ay * bx)


def quat_normalize1(q):
    """Sophus SO3::operator* first-order renormalisation: q *= 2 / (1 + |q|^2) unless |q|^2 == 1.
    This is what keeps the rotation of every pose the registration returns orthonormal."""
    w, x, y, z = q
    n2 = ((w * w + x * x) + y * y) + z * z
    if n2 != 1.0:
        s = 2.0 / (1.0 + n2)
        return (w * s, x * s, y * s, z * s)
    return (w, x, y, z)


class SE3q:
    """Sophus::SE3d as kiss-icp's registration holds it: unit quaternion (w,x,y,z) + translation.
    Products renormalise the quaternion (Sophus), so chains of them never drift away from SO(3)."""

    def __init__(self, q=(1.0, 0.0, 0.0, 0.0), t=(0.0, 0.0, 0.0)):
        self.q = tuple(float(v) for v in q)
        self.t = tuple(float(v) for v in t)

    @staticmethod
    def from_matrix(T):
        T = np.asarray(T, dtype=np.float64)
        return SE3q(rot_to_quat(T[:3, :3]), (T[0, 3], T[1, 3], T[2, 3]))

    def rot(self):
        return quat_to_rot(*self.q)

    def mul(self, other):
        """self * other: q = normalize1(qa qb), t = Ra tb + ta."""
        q = quat_normalize1(quat_mul(self.q, other.q))
        R = self.rot()
        x, y, z = other.t
        t = tuple(float(((R[i, 0] * x + R[i, 1] * y) + R[i, 2] * z) + self.t[i]) for i in range(3))
        return SE3q(q, t)

    def matrix(self):
        T = np.eye(4)
        T[:3, :3] = self.rot()
        T[:3, 3] = self.t
        return T


def se3_exp(tangent, return_quat=False):
    """tangent (...,6) = [upsilon, omega] -> R (...,3,3), t (...,3).

    Follows Sophus SE3::exp / SO3::expAndTheta: unit quaternion from the half angle,
    Eigen's quaternion->matrix expansion, V = I + a*Om + b*Om^2 (V = R for theta<eps).
    Uses det_sincos instead of libm (see module docstring)."""
    tg = np.asarray(tangent, dtype=np.float64)
    ux, uy, uz = tg[..., 0], tg[..., 1], tg[..., 2]
    wx, wy, wz = tg[..., 3], tg[..., 4], tg[..., 5]
    theta_sq = (wx * wx + wy * wy) + wz * wz
    small = theta_sq < EPS * EPS
    theta = np.where(small, 0.0, np.sqrt(theta_sq))
    half = 0.5 * theta
    sh, ch = det_sincos(half)
    with np.errstate(divide="ignore", invalid="ignore"):
        po4 = theta_sq * theta_sq
        imag_s = (0.5 - (1.0 / 48.0) * theta_sq) + (1.0 / 3840.0) * po4
        real_s = (1.0 - (1.0 / 8.0) * theta_sq) + (1.0 / 384.0) * po4
        imag = np.where(small, imag_s, sh / theta)
        real = np.where(small, real_s, ch)
    qw = real
    qx, qy, qz = imag * wx, imag * wy, imag * wz
    R = quat_to_rot(qw, qx, qy, qz)
    # V
    st, ct = det_sincos(theta)
    with np.errstate(divide="ignore", invalid="ignore"):
        tsq = theta * theta
        a = (1.0 - ct) / tsq
        b = (theta - st) / (tsq * theta)
    o00 = -(wy * wy + wz * wz)
    o11 = -(wx * wx + wz * wz)
    o22 = -(wx * wx + wy * wy)
    o01, o02, o12 = wx * wy, wx * wz, wy * wz
    big = theta >= EPS  # Sophus: V = so3.matrix() if theta < eps
    V = np.empty_like(R)
    V[..., 0, 0] = np.where(big, 1.0 + b * o00, R[..., 0, 0])
    V[..., 0, 1] = np.where(big, a * (-wz) + b * o01, R[..., 0, 1])
    V[..., 0, 2] = np.where(big, a * wy + b * o02, R[..., 0, 2])
    V[..., 1, 0] = np.where(big, a * wz + b * o01, R[..., 1, 0])
    V[..., 1, 1] = np.where(big, 1.0 + b * o11, R[..., 1, 1])
    V[..., 1, 2] = np.where(big, a * (-wx) + b * o12, R[..., 1, 2])
    V[..., 2, 0] = np.where(big, a * (-wy) + b * o02, R[..., 2, 0])
    V[..., 2, 1] = np.where(big, a * wx + b * o12, R[..., 2, 1])
    V[..., 2, 2] = np.where(big, 1.0 + b * o22, R[..., 2, 2])
    t = np.empty(tg.shape[:-1] + (3,))
    t[..., 0] = (V[..., 0, 0] * ux + V[..., 0, 1] * uy) + V[..., 0, 2] * uz
    t[..., 1] = (V[..., 1, 0] * ux + V[..., 1, 1] * uy) + V[..., 1, 2] * uz
    t[..., 2] = (V[..., 2, 0] * ux + V[..., 2, 1] * uy) + V[..., 2, 2] * uz
    if return_quat:
        return R, t, (qw, qx, qy, qz)
    return R, t


def se3_exp_mat(tangent):
    R, t = se3_exp(np.asarray(tangent, dtype=np.float64))
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def se3_exp_q(tangent):
    """SE3::exp of one tangent as (SE3q, 4x4 matrix); the matrix is the one points are moved with."""
    R, t, q = se3_exp(np.asarray(tangent, dtype=np.float64), return_quat=True)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return SE3q(tuple(float(v) for v in q), tuple(float(v) for v in t)), T


# ---------------------------------------------------------------------------
# SO3 / SE3 log (host-side scalars, libm through `math`)
# ---------------------------------------------------------------------------
def rot_to_quat(R):
    """Eigen's rotation-matrix -> quaternion conversion; returns (w, x, y, z)."""
    m = [[float(R[i][j]) for j in range(3)] for i in range(3)]
    t = (m[0][0] + m[1][1]) + m[2][2]
    q = [0.0, 0.0, 0.0]
    if t > 0.0:
        t = math.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        q[0] = (m[2][1] - m[1][2]) * t
        q[1] = (m[0][2] - m[2][0]) * t
        q[2] = (m[1][0] - m[0][1]) * t
    else:
        i = 0
        if m[1][1] > m[0][0]:
            i = 1
        if m[2][2] > m[i][i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(((m[i][i] - m[j][j]) - m[k][k]) + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        w = (m[k][j] - m[j][k]) * t
        q[j] = (m[j][i] + m[i][j]) * t
        q[k] = (m[k][i] + m[i][k]) * t
    return w, q[0], q[1], q[2]


def so3_log(R):
    """Sophus SO3::logAndTheta on the quaternion of R: returns (omega[3], theta)."""
    w, x, y, z = rot_to_quat(R)
    sq_n = (x * x + y * y) + z * z
    if sq_n < EPS * EPS:
        sq_w = w * w
        two_atan = 2.0 / w - (2.0 / 3.0) * sq_n / (w * sq_w)
        theta = 2.0 * sq_n / w
    else:
        n = math.sqrt(sq_n)
        atan_nbyw = math.atan2(-n, -w) if w < 0.0 else math.atan2(n, w)
        two_atan = 2.0 * atan_nbyw / n
        theta = two_atan * n
    return [two_atan * x, two_atan * y, two_atan * z], theta


def se3_log(T):
    """Sophus SE3::log of a rigid 4x4 -> [upsilon(3), omega(3)] as ndarray(6)."""
    T = np.asarray(T, dtype=np.float64)
    om, theta = so3_log(T[:3, :3])
    wx, wy, wz = om
    o00 = -(wy * wy + wz * wz)
    o11 = -(wx * wx + wz * wz)
    o22 = -(wx * wx + wy * wy)
    o01, o02, o12 = wx * wy, wx * wz, wy * wz
    if abs(theta) < EPS:
        c = 1.0 / 12.0
    else:
        half = 0.5 * theta
        c = (1.0 - (theta * math.cos(half)) / (2.0 * math.sin(half))) / (theta * theta)
    v = [[1.0 + c * o00, 0.5 * wz + c * o01, -0.5 * wy + c * o02],
         [-0.5 * wz + c * o01, 1.0 + c * o11, 0.5 * wx + c * o12],
         [0.5 * wy + c * o02, -0.5 * wx + c * o12, 1.0 + c * o22]]
    tx, ty, tz = float(T[0, 3]), float(T[1, 3]), float(T[2, 3])
    up = [(v[i][0] * tx + v[i][1] * ty) + v[i][2] * tz for i in range(3)]
    return np.array(up + [wx, wy, wz], dtype=np.float64)


def rot_angle(R):
    """Eigen::AngleAxisd(R).angle(): 2*atan2(|q.vec|, |q.w|) in [0, pi]."""
    w, x, y, z = rot_to_quat(R)
    n = math.sqrt((x * x + y * y) + z * z)
    return 2.0 * math.atan2(n, abs(w))


# ---------------------------------------------------------------------------
# Canonical reduction and 6x6 solve
# ---------------------------------------------------------------------------
def pairwise_tree_sum(a):
    """Adjacent-pairs binary tree over axis 0, zero-padded to a power of two >= 32
    (level 0 adds a[2k]+a[2k+1]).  This is exactly what a warp xor-butterfly with
    strides 1,2,4,8,16 followed by a tree over warp partials computes."""
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    if n == 0:
        return np.zeros(a.shape[1:])
    p = 32
    while p < n:
        p *= 2
    if p != n:
        pad = np.zeros((p - n,) + a.shape[1:])
        a = np.concatenate([a, pad], axis=0)
    while a.shape[0] > 1:
        a = a[0::2] + a[1::2]
    return a[0]


def ldlt_solve6(A, b):
    """Solve A x = b for symmetric 6x6 A by LDL^T with diagonal pivoting (largest
    |diagonal| first, as Eigen's LDLT does); plain scalar ops in a fixed order.
    Returns (x, ok)."""
    n = 6
    a = [[float(A[i][j]) for j in range(n)] for i in range(n)]
    perm = list(range(n))
    for k in range(n):
        # pivot: largest |a[i][i]|, i >= k (first wins ties)
        p = k
        best = abs(a[k][k])
        for i in range(k + 1, n):
            v = abs(a[i][i])
            if v > best:
                best = v
                p = i
        if p != k:
            a[k], a[p] = a[p], a[k]
            for r in range(n):
                a[r][k], a[r][p] = a[r][p], a[r][k]
            perm[k], perm[p] = perm[p], perm[k]
        d = a[k][k]
        if d == 0.0 or d != d:
            return [0.0] * n, False
        for i in range(k + 1, n):
            a[i][k] = a[i][k] / d  # L(i,k)
        for j in range(k + 1, n):
            ljd = a[j][k] * d
            for i in range(j, n):
                a[i][j] = a[i][j] - a[i][k] * ljd
                a[j][i] = a[i][j]
    # solve: y = P b ; L z = y ; w = z / D ; L^T v = w ; x = P^T v
    y = [float(b[perm[i]]) for i in range(n)]
    for i in range(n):
        s = y[i]
        for j in range(i):
            s = s - a[i][j] * y[j]
        y[i] = s
    for i in range(n):
        y[i] = y[i] / a[i][i]
    for i in range(n - 1, -1, -1):
        s = y[i]
        for j in range(i + 1, n):
            s = s - a[j][i] * y[j]
        y[i] = s
    x = [0.0] * n
    for i in range(n):
        x[perm[i]] = y[i]
    ok = all(v == v and abs(v) != float("inf") for v in x)
    return x, ok
