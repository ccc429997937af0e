// Host-native 18-state error-state EKF (the consumer of the odometry poses), C ABI in include/ptk.h.
//
// Same filter as ptudes_lab_b200/ins/es_ekf.py, i.e. the reference's ESEKF
// (/root/reference/src/ptudes/ins/es_ekf.py:57-329): IMU mechanisation + covariance propagation
// (processImu :191-257) and the 6-D pose update (processPose :259-329), state order
// pos 0, vel 3, phi 6, gyro bias 9, accel bias 12, gravity 15 (:65-71), noise constants :115-118,
// initial covariance :99-135.  The reference's author flags the Python filter as a test bed that wants
// a C++ core (:60-62); at 100 Hz x tens of sequences the Python one is what the host spends its time on
// once the lidar step runs on the GPU.  Plain C++, no CUDA; the prediction exploits the block structure of F.
#include <math.h>
#include <string.h>

#include <new>

#include "../../include/ptk.h"

namespace {

constexpr int N = 18;
constexpr int POS = 0, VEL = 3, PHI = 6, BG = 9, BA = 12, GR = 15;
constexpr double GRAV = 9.782940329221166;            // ins/data.py:10
constexpr double ACC_BIAS_STD = 0.049, GYR_BIAS_STD = 0.38, ACC_VRW = 0.0043, GYR_ARW = 0.000466;

struct M3 { double a[9]; };

M3 m3_identity() { M3 r = {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; return r; }
M3 m3_mul(const M3& x, const M3& y) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[3 * i + j] = x.a[3 * i] * y.a[j] + x.a[3 * i + 1] * y.a[3 + j] + x.a[3 * i + 2] * y.a[6 + j];
    return r;
}
M3 m3_t(const M3& x) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.a[3 * i + j] = x.a[3 * j + i];
    return r;
}
void m3_vec(const M3& x, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = x.a[3 * i] * v[0] + x.a[3 * i + 1] * v[1] + x.a[3 * i + 2] * v[2];
}
M3 skew(const double* v) { M3 r = {{0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0}}; return r; }

M3 so3_exp(const double* w) {      // Rodrigues; series below 1e-8 rad
    double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    M3 K = skew(w), K2 = m3_mul(K, K), R = m3_identity();
    double a, b;
    if (th < 1e-8) { a = 1.0; b = 0.5; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / (th * th); }
    for (int i = 0; i < 9; ++i) R.a[i] += a * K.a[i] + b * K2.a[i];
    return R;
}

void so3_log(const M3& R, double* w) {
    double tr = R.a[0] + R.a[4] + R.a[8];
    double v[3] = {R.a[7] - R.a[5], R.a[2] - R.a[6], R.a[3] - R.a[1]};
    double s = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), c = tr - 1.0;
    double th = atan2(s, c);
    if (s > 1e-8) { for (int i = 0; i < 3; ++i) w[i] = v[i] * (th / s); return; }
    if (c > 0.0) { for (int i = 0; i < 3; ++i) w[i] = 0.5 * v[i]; return; }
    double A[9];
    for (int i = 0; i < 9; ++i) A[i] = 0.5 * (R.a[i] + (i % 4 == 0 ? 1.0 : 0.0));
    int k = 0;
    if (A[4] > A[0]) k = 1;
    if (A[8] > A[4 * k]) k = 2;
    double d = sqrt(A[4 * k]);
    for (int i = 0; i < 3; ++i) w[i] = A[3 * i + k] / d * th;
}

// inverse of a 6x6 by Gauss-Jordan with partial pivoting; false if singular
bool inv6(const double* S, double* Si) {
    double a[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { a[i][j] = S[6 * i + j]; a[i][6 + j] = i == j ? 1.0 : 0.0; }
    for (int c = 0; c < 6; ++c) {
        int p = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
        if (a[p][c] == 0.0) return false;
        if (p != c) for (int j = 0; j < 12; ++j) { double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
        double d = a[c][c];
        for (int j = 0; j < 12; ++j) a[c][j] /= d;
        for (int r = 0; r < 6; ++r) {
            if (r == c) continue;
            double f = a[r][c];
            if (f != 0.0) for (int j = 0; j < 12; ++j) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) Si[6 * i + j] = a[i][6 + j];
    return true;
}

}  // namespace

struct ptk_ekf {
    double pos[3], vel[3], bg[3], ba[3], grav[3];
    M3 att;
    double P[N * N];
    double ts, prev_ts;
    bool initialized;
    long long n_imu, n_pose;
};

extern "C" int ptk_ekf_create(ptk_ekf** out, const double* init_grav, const double* init_bacc, const double* init_bgyr) {
    if (!out) return PTK_E_ARG;
    ptk_ekf* f = new (std::nothrow) ptk_ekf();
    if (!f) return PTK_E_CAPACITY;
    memset(f, 0, sizeof(*f));
    f->att = m3_identity();
    for (int i = 0; i < 3; ++i) {
        f->grav[i] = init_grav ? init_grav[i] : (i == 2 ? -GRAV : 0.0);
        f->ba[i] = init_bacc ? init_bacc[i] : 0.0;
        f->bg[i] = init_bgyr ? init_bgyr[i] : 0.0;
    }
    // initial covariance (es_ekf.py:99-135): the attitude sigma is the rotation vector of the intrinsic
    // XYZ Euler rotation by (10, 10, 10) degrees
    const double a = 10.0 * M_PI / 180.0;
    double ex[3] = {a, 0, 0}, ey[3] = {0, a, 0}, ez[3] = {0, 0, a}, att_sigma[3];
    so3_log(m3_mul(m3_mul(so3_exp(ex), so3_exp(ey)), so3_exp(ez)), att_sigma);
    double sig[N];
    for (int i = 0; i < 3; ++i) { sig[POS + i] = 10.0; sig[VEL + i] = 5.0; sig[PHI + i] = att_sigma[i]; sig[BG + i] = 1.5; sig[BA + i] = 0.5; sig[GR + i] = 2.5; }
    for (int i = 0; i < N; ++i) f->P[i * N + i] = sig[i] * sig[i];
    *out = f;
    return PTK_OK;
}

extern "C" int ptk_ekf_destroy(ptk_ekf* f) { delete f; return PTK_OK; }

// predict (es_ekf.py:191-257)
extern "C" int ptk_ekf_process_imu(ptk_ekf* f, const double* lacc, const double* avel, double ts) {
    if (!f || !lacc || !avel) return PTK_E_ARG;
    f->prev_ts = f->ts;
    f->ts = ts;
    f->n_imu++;
    if (!f->initialized) { f->initialized = true; return PTK_OK; }      // the first sample only sets the clock
    const double dt = f->ts - f->prev_ts;
    const M3 Rp = f->att;
    double fb[3], wb[3], wdt[3], a_nav[3];
    for (int i = 0; i < 3; ++i) { fb[i] = lacc[i] - f->ba[i]; wb[i] = avel[i] - f->bg[i]; wdt[i] = wb[i] * dt; }
    const M3 dR = so3_exp(wdt);
    m3_vec(Rp, fb, a_nav);
    for (int i = 0; i < 3; ++i) {
        a_nav[i] += f->grav[i];
        f->pos[i] = f->pos[i] + f->vel[i] * dt + 0.5 * a_nav[i] * dt * dt;
        f->vel[i] = f->vel[i] + a_nav[i] * dt;
    }
    f->att = m3_mul(Rp, dR);
    // P <- F P F^T with F = I except the blocks (POS,VEL) = dt I, (VEL,PHI) = -dt R [f]x, (VEL,BA) = -dt R,
    // (PHI,PHI) = dR^T, (PHI,BG) = -dt I: only the POS, VEL and PHI block rows / columns change, so the two
    // products are done on those 3 x 18 / 18 x 3 panels instead of as dense 18^3 multiplications.
    const M3 RK = m3_mul(Rp, skew(fb)), dRt = m3_t(dR);
    static thread_local double T1[N * N];
    double* P = f->P;
    // T1 = F P  (row panels)
    memcpy(T1, P, sizeof(T1));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < N; ++j) {
            T1[(POS + i) * N + j] = P[(POS + i) * N + j] + dt * P[(VEL + i) * N + j];
            double v = P[(VEL + i) * N + j], ph = 0.0;
            for (int k = 0; k < 3; ++k) {
                v -= dt * (RK.a[3 * i + k] * P[(PHI + k) * N + j] + Rp.a[3 * i + k] * P[(BA + k) * N + j]);
                ph += dRt.a[3 * i + k] * P[(PHI + k) * N + j];
            }
            T1[(VEL + i) * N + j] = v;
            T1[(PHI + i) * N + j] = ph - dt * P[(BG + i) * N + j];
        }
    // P = T1 F^T  (column panels)
    memcpy(P, T1, sizeof(T1));
    for (int r = 0; r < N; ++r)
        for (int i = 0; i < 3; ++i) {
            P[r * N + POS + i] = T1[r * N + POS + i] + dt * T1[r * N + VEL + i];
            double v = T1[r * N + VEL + i], ph = 0.0;
            for (int k = 0; k < 3; ++k) {
                v -= dt * (RK.a[3 * i + k] * T1[r * N + PHI + k] + Rp.a[3 * i + k] * T1[r * N + BA + k]);
                ph += dRt.a[3 * i + k] * T1[r * N + PHI + k];
            }
            P[r * N + VEL + i] = v;
            P[r * N + PHI + i] = ph - dt * T1[r * N + BG + i];
        }
    for (int i = 0; i < 3; ++i) {
        f->P[(VEL + i) * N + VEL + i] += (dt * ACC_BIAS_STD) * (dt * ACC_BIAS_STD);
        f->P[(PHI + i) * N + PHI + i] += (dt * GYR_BIAS_STD) * (dt * GYR_BIAS_STD);
        f->P[(BA + i) * N + BA + i] += dt * ACC_VRW * ACC_VRW;
        f->P[(BG + i) * N + BG + i] += dt * GYR_ARW * GYR_ARW;
    }
    return PTK_OK;
}

extern "C" int ptk_ekf_process_imu_batch(ptk_ekf* f, const double* lacc, const double* avel, const double* ts, int n) {
    if (!f || n < 0 || (n > 0 && (!lacc || !avel || !ts))) return PTK_E_ARG;
    for (int k = 0; k < n; ++k) {
        int rc = ptk_ekf_process_imu(f, lacc + 3 * k, avel + 3 * k, ts[k]);
        if (rc) return rc;
    }
    return PTK_OK;
}

// update with a pose measurement (es_ekf.py:259-329); meas_cov 6x6 row-major or NULL (2 cm / 0.01 rad)
extern "C" int ptk_ekf_process_pose(ptk_ekf* f, const double* pose16, const double* meas_cov) {
    if (!f || !pose16) return PTK_E_ARG;
    f->n_pose++;
    M3 Rm;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rm.a[3 * i + j] = pose16[4 * i + j];
    double resid[6], w[3];
    for (int i = 0; i < 3; ++i) resid[i] = pose16[4 * i + 3] - f->pos[i];
    so3_log(m3_mul(m3_t(f->att), Rm), w);
    for (int i = 0; i < 3; ++i) resid[3 + i] = w[i];
    // H picks rows/columns POS and PHI: S = P[idx, idx] + R, K = P[:, idx] S^-1
    const int idx[6] = {POS, POS + 1, POS + 2, PHI, PHI + 1, PHI + 2};
    double S[36], Si[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double r = meas_cov ? meas_cov[6 * i + j] : (i == j ? (i < 3 ? 0.02 * 0.02 : 0.01 * 0.01) : 0.0);
            S[6 * i + j] = f->P[idx[i] * N + idx[j]] + r;
        }
    if (!inv6(S, Si)) return PTK_E_NUMERIC;
    double K[N * 6], dx[N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += f->P[i * N + idx[k]] * Si[6 * k + j];
            K[i * 6 + j] = s;
        }
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int j = 0; j < 6; ++j) s += K[i * 6 + j] * resid[j];
        dx[i] = s;
    }
    // P = (I - K H) P
    static thread_local double Pn[N * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = f->P[i * N + j];
            for (int k = 0; k < 6; ++k) s -= K[i * 6 + k] * f->P[idx[k] * N + j];
            Pn[i * N + j] = s;
        }
    memcpy(f->P, Pn, sizeof(Pn));
    for (int i = 0; i < 3; ++i) {
        f->pos[i] += dx[POS + i]; f->vel[i] += dx[VEL + i]; f->bg[i] += dx[BG + i]; f->ba[i] += dx[BA + i]; f->grav[i] += dx[GR + i];
    }
    f->att = m3_mul(f->att, so3_exp(dx + PHI));
    // reset: attitude block through G = I - [dphi / 2]x
    double half[3] = {0.5 * dx[PHI], 0.5 * dx[PHI + 1], 0.5 * dx[PHI + 2]};
    M3 G = m3_identity(), Kh = skew(half), B, Gt;
    for (int i = 0; i < 9; ++i) G.a[i] -= Kh.a[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) B.a[3 * i + j] = f->P[(PHI + i) * N + PHI + j];
    Gt = m3_t(G);
    B = m3_mul(m3_mul(G, B), Gt);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) f->P[(PHI + i) * N + PHI + j] = B.a[3 * i + j];
    return PTK_OK;
}

extern "C" int ptk_ekf_get_nav(const ptk_ekf* f, double* pos, double* att9, double* vel, double* bias_gyr, double* bias_acc,
                               double* grav) {
    if (!f) return PTK_E_ARG;
    if (pos) memcpy(pos, f->pos, 24);
    if (att9) memcpy(att9, f->att.a, 72);
    if (vel) memcpy(vel, f->vel, 24);
    if (bias_gyr) memcpy(bias_gyr, f->bg, 24);
    if (bias_acc) memcpy(bias_acc, f->ba, 24);
    if (grav) memcpy(grav, f->grav, 24);
    return PTK_OK;
}

extern "C" int ptk_ekf_get_pose(const ptk_ekf* f, double* pose16) {
    if (!f || !pose16) return PTK_E_ARG;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) pose16[4 * i + j] = f->att.a[3 * i + j];
        pose16[4 * i + 3] = f->pos[i];
    }
    pose16[12] = pose16[13] = pose16[14] = 0.0;
    pose16[15] = 1.0;
    return PTK_OK;
}

extern "C" int ptk_ekf_get_cov(const ptk_ekf* f, double* cov324) {
    if (!f || !cov324) return PTK_E_ARG;
    memcpy(cov324, f->P, sizeof(f->P));
    return PTK_OK;
}

extern "C" double ptk_ekf_ts(const ptk_ekf* f) { return f ? f->ts : 0.0; }
