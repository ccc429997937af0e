// Canonical float64 arithmetic of the odometry step, shared by host and device code.
//
// Every function here has a fixed order of IEEE-754 double operations (one rounding per
// written operator; the translation unit is compiled with -fmad=false and the host pass
// without -mfma), so results are bit-identical on the CPU, on the GPU, and in the test
// oracle that restates the same formulas.  The formulas follow Sophus SE3/SO3 exp/log,
// Eigen's quaternion<->matrix conversions and LDLT as used by kiss-icp 0.2.x, the
// third-party package that /root/reference/src/ptudes/kiss.py:7-10 calls into.
#pragma once
#include <math.h>

#define PTK_HD __host__ __device__ __forceinline__

namespace ptk {

constexpr double kSophusEps = 1e-10;

// ---- deterministic sin/cos (Cody-Waite reduction + Taylor kernels) -------------------
PTK_HD void det_sincos(double x, double* sn, double* cs) {
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_1 = 1.57079632673412561417e+00;
    const double PIO2_2 = 6.07710050630396597660e-11;
    const double PIO2_3 = 2.02226624879595063154e-21;
    double k = rint(x * TWO_OVER_PI);
    double r = ((x - k * PIO2_1) - k * PIO2_2) - k * PIO2_3;
    double z = r * r;
    double ps = 1.0 / 355687428096000.0;
    ps = ps * z + (-1.0 / 1307674368000.0);
    ps = ps * z + (1.0 / 6227020800.0);
    ps = ps * z + (-1.0 / 39916800.0);
    ps = ps * z + (1.0 / 362880.0);
    ps = ps * z + (-1.0 / 5040.0);
    ps = ps * z + (1.0 / 120.0);
    ps = ps * z + (-1.0 / 6.0);
    double s = r + (r * z) * ps;
    double pc = -1.0 / 6402373705728000.0;
    pc = pc * z + (1.0 / 20922789888000.0);
    pc = pc * z + (-1.0 / 87178291200.0);
    pc = pc * z + (1.0 / 479001600.0);
    pc = pc * z + (-1.0 / 3628800.0);
    pc = pc * z + (1.0 / 40320.0);
    pc = pc * z + (-1.0 / 720.0);
    pc = pc * z + (1.0 / 24.0);
    double c = 1.0 - (0.5 * z - (z * z) * pc);
    long long q = ((long long)k) & 3LL;
    if (q == 0)      { *sn = s;  *cs = c; }
    else if (q == 1) { *sn = c;  *cs = -s; }
    else if (q == 2) { *sn = -s; *cs = -c; }
    else             { *sn = -c; *cs = s; }
}

// ---- rigid transforms: row-major R[9], t[3] ---------------------------------------------
struct Rigid {
    double r[9];
    double t[3];
};

PTK_HD Rigid rigid_identity() {
    Rigid T;
    T.r[0] = 1; T.r[1] = 0; T.r[2] = 0;
    T.r[3] = 0; T.r[4] = 1; T.r[5] = 0;
    T.r[6] = 0; T.r[7] = 0; T.r[8] = 1;
    T.t[0] = 0; T.t[1] = 0; T.t[2] = 0;
    return T;
}

PTK_HD Rigid rigid_mul(const Rigid& A, const Rigid& B) {
    Rigid C;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double a0 = A.r[3 * i], a1 = A.r[3 * i + 1], a2 = A.r[3 * i + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) C.r[3 * i + j] = (a0 * B.r[j] + a1 * B.r[3 + j]) + a2 * B.r[6 + j];
        C.t[i] = ((a0 * B.t[0] + a1 * B.t[1]) + a2 * B.t[2]) + A.t[i];
    }
    return C;
}

PTK_HD Rigid rigid_inv(const Rigid& T) {
    Rigid C;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double r0 = T.r[i], r1 = T.r[3 + i], r2 = T.r[6 + i];
        C.r[3 * i] = r0; C.r[3 * i + 1] = r1; C.r[3 * i + 2] = r2;
        C.t[i] = -((r0 * T.t[0] + r1 * T.t[1]) + r2 * T.t[2]);
    }
    return C;
}

PTK_HD void rigid_apply(const Rigid& T, double x, double y, double z, double& xo, double& yo, double& zo) {
    xo = ((T.r[0] * x + T.r[1] * y) + T.r[2] * z) + T.t[0];
    yo = ((T.r[3] * x + T.r[4] * y) + T.r[5] * z) + T.t[1];
    zo = ((T.r[6] * x + T.r[7] * y) + T.r[8] * z) + T.t[2];
}

PTK_HD void rigid_from_mat16(const double* m, Rigid& T) {
    T.r[0] = m[0]; T.r[1] = m[1]; T.r[2] = m[2];  T.t[0] = m[3];
    T.r[3] = m[4]; T.r[4] = m[5]; T.r[5] = m[6];  T.t[1] = m[7];
    T.r[6] = m[8]; T.r[7] = m[9]; T.r[8] = m[10]; T.t[2] = m[11];
}

PTK_HD void rigid_to_mat16(const Rigid& T, double* m) {
    m[0] = T.r[0]; m[1] = T.r[1]; m[2] = T.r[2];  m[3] = T.t[0];
    m[4] = T.r[3]; m[5] = T.r[4]; m[6] = T.r[5];  m[7] = T.t[1];
    m[8] = T.r[6]; m[9] = T.r[7]; m[10] = T.r[8]; m[11] = T.t[2];
    m[12] = 0; m[13] = 0; m[14] = 0; m[15] = 1;
}

// ---- SE3 exp (Sophus SE3::exp / SO3::expAndTheta, Eigen quaternion->matrix) ----------
struct Quat { double w, x, y, z; };

// Eigen quaternion product a*b, sums left to right.
PTK_HD Quat quat_mul(const Quat& a, const Quat& b) {
    Quat c;
    c.w = ((a.w * b.w - a.x * b.x) - a.y * b.y) - a.z * b.z;
    c.x = ((a.w * b.x + a.x * b.w) + a.y * b.z) - a.z * b.y;
    c.y = ((a.w * b.y + a.y * b.w) + a.z * b.x) - a.x * b.z;
    c.z = ((a.w * b.z + a.z * b.w) + a.x * b.y) - a.y * b.x;
    return c;
}

// Sophus SO3::operator* first-order renormalisation.
PTK_HD Quat quat_normalize1(Quat q) {
    const double n2 = ((q.w * q.w + q.x * q.x) + q.y * q.y) + q.z * q.z;
    if (n2 != 1.0) {
        const double s = 2.0 / (1.0 + n2);
        q.w = q.w * s; q.x = q.x * s; q.y = q.y * s; q.z = q.z * s;
    }
    return q;
}

// Eigen QuaternionBase::toRotationMatrix
PTK_HD void quat_to_rot(const Quat& q, double* r) {
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    r[0] = 1.0 - (tyy + tzz); r[1] = txy - twz;         r[2] = txz + twy;
    r[3] = txy + twz;         r[4] = 1.0 - (txx + tzz); r[5] = tyz - twx;
    r[6] = txz - twy;         r[7] = tyz + twx;         r[8] = 1.0 - (txx + tyy);
}

PTK_HD Rigid se3_exp_q(const double* a, Quat* qout) {
    const double ux = a[0], uy = a[1], uz = a[2], wx = a[3], wy = a[4], wz = a[5];
    const double theta_sq = (wx * wx + wy * wy) + wz * wz;
    const bool small_ = theta_sq < kSophusEps * kSophusEps;
    const double theta = small_ ? 0.0 : sqrt(theta_sq);
    double imag, real;
    if (small_) {
        const double po4 = theta_sq * theta_sq;
        imag = (0.5 - (1.0 / 48.0) * theta_sq) + (1.0 / 3840.0) * po4;
        real = (1.0 - (1.0 / 8.0) * theta_sq) + (1.0 / 384.0) * po4;
    } else {
        double sh, ch;
        det_sincos(0.5 * theta, &sh, &ch);
        imag = sh / theta;
        real = ch;
    }
    Quat q;
    q.w = real; q.x = imag * wx; q.y = imag * wy; q.z = imag * wz;
    if (qout) *qout = q;
    Rigid T;
    quat_to_rot(q, T.r);
    double v[9];
    if (theta >= kSophusEps) {
        double st, ct;
        det_sincos(theta, &st, &ct);
        const double tsq = theta * theta;
        const double ca = (1.0 - ct) / tsq;
        const double cb = (theta - st) / (tsq * theta);
        const double o00 = -(wy * wy + wz * wz), o11 = -(wx * wx + wz * wz), o22 = -(wx * wx + wy * wy);
        const double o01 = wx * wy, o02 = wx * wz, o12 = wy * wz;
        v[0] = 1.0 + cb * o00;          v[1] = ca * (-wz) + cb * o01;   v[2] = ca * wy + cb * o02;
        v[3] = ca * wz + cb * o01;      v[4] = 1.0 + cb * o11;          v[5] = ca * (-wx) + cb * o12;
        v[6] = ca * (-wy) + cb * o02;   v[7] = ca * wx + cb * o12;      v[8] = 1.0 + cb * o22;
    } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = T.r[i];
    }
    T.t[0] = (v[0] * ux + v[1] * uy) + v[2] * uz;
    T.t[1] = (v[3] * ux + v[4] * uy) + v[5] * uz;
    T.t[2] = (v[6] * ux + v[7] * uy) + v[8] * uz;
    return T;
}

PTK_HD Rigid se3_exp(const double* a) { return se3_exp_q(a, nullptr); }

// Sophus::SE3d as kiss-icp's registration holds poses: unit quaternion + translation.  Products
// renormalise the quaternion, so the rotation of every returned pose stays orthonormal.
struct SE3q {
    Quat q;
    double t[3];
};

PTK_HD SE3q se3q_identity() {
    SE3q T;
    T.q.w = 1.0; T.q.x = 0.0; T.q.y = 0.0; T.q.z = 0.0;
    T.t[0] = 0.0; T.t[1] = 0.0; T.t[2] = 0.0;
    return T;
}

// a * b: q = normalize1(qa qb), t = R(qa) tb + ta
PTK_HD SE3q se3q_mul(const SE3q& a, const SE3q& b) {
    SE3q c;
    c.q = quat_normalize1(quat_mul(a.q, b.q));
    double r[9];
    quat_to_rot(a.q, r);
#pragma unroll
    for (int i = 0; i < 3; ++i) c.t[i] = ((r[3 * i] * b.t[0] + r[3 * i + 1] * b.t[1]) + r[3 * i + 2] * b.t[2]) + a.t[i];
    return c;
}

PTK_HD Rigid se3q_matrix(const SE3q& a) {
    Rigid T;
    quat_to_rot(a.q, T.r);
    T.t[0] = a.t[0]; T.t[1] = a.t[1]; T.t[2] = a.t[2];
    return T;
}

// ---- 6x6 LDL^T with diagonal pivoting, fixed op order --------------------------------
// A is overwritten.  Returns false for a zero/NaN pivot or a non-finite solution.
PTK_HD bool ldlt_solve6(double (*a)[6], const double* b, double* x) {
    int perm[6] = {0, 1, 2, 3, 4, 5};
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double best = fabs(a[k][k]);
        for (int i = k + 1; i < 6; ++i) {
            double v = fabs(a[i][i]);
            if (v > best) { best = v; p = i; }
        }
        if (p != k) {
            for (int c = 0; c < 6; ++c) { double tmp = a[k][c]; a[k][c] = a[p][c]; a[p][c] = tmp; }
            for (int r = 0; r < 6; ++r) { double tmp = a[r][k]; a[r][k] = a[r][p]; a[r][p] = tmp; }
            int tp = perm[k]; perm[k] = perm[p]; perm[p] = tp;
        }
        double d = a[k][k];
        if (d == 0.0 || d != d) return false;
        for (int i = k + 1; i < 6; ++i) a[i][k] = a[i][k] / d;
        for (int j = k + 1; j < 6; ++j) {
            double ljd = a[j][k] * d;
            for (int i = j; i < 6; ++i) {
                a[i][j] = a[i][j] - a[i][k] * ljd;
                a[j][i] = a[i][j];
            }
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < 6; ++i) {
        double s = y[i];
        for (int j = 0; j < i; ++j) s = s - a[i][j] * y[j];
        y[i] = s;
    }
    for (int i = 0; i < 6; ++i) y[i] = y[i] / a[i][i];
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int j = i + 1; j < 6; ++j) s = s - a[j][i] * y[j];
        y[i] = s;
    }
    bool ok = true;
    for (int i = 0; i < 6; ++i) {
        x[perm[i]] = y[i];
        if (!(fabs(y[i]) <= 1.7976931348623157e308)) ok = false;
    }
    return ok;
}

// Eigen rotation matrix -> quaternion (sqrt and divisions only: bit-exact on host and device)
PTK_HD void rot_to_quat(const double* r, double& w, double& x, double& y, double& z) {
    double m[3][3] = {{r[0], r[1], r[2]}, {r[3], r[4], r[5]}, {r[6], r[7], r[8]}};
    double t = (m[0][0] + m[1][1]) + m[2][2];
    double q[3] = {0, 0, 0};
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[2][1] - m[1][2]) * t;
        q[1] = (m[0][2] - m[2][0]) * t;
        q[2] = (m[1][0] - m[0][1]) * t;
    } else {
        int i = 0;
        if (m[1][1] > m[0][0]) i = 1;
        if (m[2][2] > m[i][i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(((m[i][i] - m[j][j]) - m[k][k]) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        w = (m[k][j] - m[j][k]) * t;
        q[j] = (m[j][i] + m[i][j]) * t;
        q[k] = (m[k][i] + m[i][k]) * t;
    }
    x = q[0]; y = q[1]; z = q[2];
}

PTK_HD SE3q se3q_from_rigid(const Rigid& T) {
    SE3q a;
    rot_to_quat(T.r, a.q.w, a.q.x, a.q.y, a.q.z);
    a.t[0] = T.t[0]; a.t[1] = T.t[1]; a.t[2] = T.t[2];
    return a;
}

// ---- host-only scalars (once per scan): libm atan2/sin/cos like the oracle -----------

inline void so3_log(const double* r, double* om, double& theta) {
    double w, x, y, z;
    rot_to_quat(r, w, x, y, z);
    double sq_n = (x * x + y * y) + z * z;
    double two_atan;
    if (sq_n < kSophusEps * kSophusEps) {
        double sq_w = w * w;
        two_atan = 2.0 / w - (2.0 / 3.0) * sq_n / (w * sq_w);
        theta = 2.0 * sq_n / w;
    } else {
        double n = sqrt(sq_n);
        double at = (w < 0.0) ? atan2(-n, -w) : atan2(n, w);
        two_atan = 2.0 * at / n;
        theta = two_atan * n;
    }
    om[0] = two_atan * x; om[1] = two_atan * y; om[2] = two_atan * z;
}

inline void se3_log(const Rigid& T, double* out) {
    double om[3], theta;
    so3_log(T.r, om, theta);
    const double wx = om[0], wy = om[1], wz = om[2];
    const double o00 = -(wy * wy + wz * wz), o11 = -(wx * wx + wz * wz), o22 = -(wx * wx + wy * wy);
    const double o01 = wx * wy, o02 = wx * wz, o12 = wy * wz;
    double c;
    if (fabs(theta) < kSophusEps) {
        c = 1.0 / 12.0;
    } else {
        double half = 0.5 * theta;
        c = (1.0 - (theta * cos(half)) / (2.0 * sin(half))) / (theta * theta);
    }
    double v[3][3] = {{1.0 + c * o00, 0.5 * wz + c * o01, -0.5 * wy + c * o02},
                      {-0.5 * wz + c * o01, 1.0 + c * o11, 0.5 * wx + c * o12},
                      {0.5 * wy + c * o02, -0.5 * wx + c * o12, 1.0 + c * o22}};
    for (int i = 0; i < 3; ++i) out[i] = (v[i][0] * T.t[0] + v[i][1] * T.t[1]) + v[i][2] * T.t[2];
    out[3] = wx; out[4] = wy; out[5] = wz;
}

inline double rot_angle(const double* r) {
    double w, x, y, z;
    rot_to_quat(r, w, x, y, z);
    double n = sqrt((x * x + y * y) + z * z);
    return 2.0 * atan2(n, fabs(w));
}

}  // namespace ptk
