/* kiss_port.c - plain C restatement of the kiss-icp 0.2.x odometry step that ptudes-lab's
 * KissICPWrapper drives.  TEST INFRASTRUCTURE ONLY: it is the second, independently written
 * CPU oracle (the first is oracle/kiss_oracle.py) and the CPU baseline bench.py times beside
 * the CUDA path.  Nothing in ptudes_lab_b200/ may load it.
 *
 * PARITY UNPINNED: the reference (/root/reference) holds no tests or golden vectors for this
 * path and the arithmetic lives in the absent third-party package kiss-icp (effective
 * 0.2.9/0.2.10, /root/reference/setup.py:22).  The order of operations follows
 * /root/reference/src/ptudes/kiss.py:83-131; the kiss-icp internals follow SURVEY.md
 * Appendix A with the canonical rules of Appendix B.  Floating point follows the canon of
 * oracle/canon.py (one IEEE double operation per written operator, no FMA: build with
 * -ffp-contract=off), so the NumPy oracle, this file and the CUDA kernels agree bit for bit.
 *
 * Structure mirrors upstream's C++ core: serial hash-map inserts for the two voxel grids and
 * for AddPoints (upstream: tsl::robin_map, serial), parallel-for over points for the deskew and
 * over source points for GetCorrespondences / BuildLinearSystem (upstream: TBB; here OpenMP).
 *
 * build: oracle/build_port.py  (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KP_MAXP 20
#define KEY_BIAS (1 << 20)
#define KP_EPS 1e-10 /* Sophus Constants<double>::epsilon() */

typedef uint64_t u64;

/* ------------------------------------------------------------------ det_sincos (canon) */
static const double TWO_OVER_PI = 6.36619772367581382433e-01;
static const double PIO2_1 = 1.57079632673412561417e+00;
static const double PIO2_2 = 6.07710050630396597660e-11;
static const double PIO2_3 = 2.02226624879595063154e-21;
static const double SC[8] = {-1.0 / 6.0, 1.0 / 120.0, -1.0 / 5040.0, 1.0 / 362880.0, -1.0 / 39916800.0,
                             1.0 / 6227020800.0, -1.0 / 1307674368000.0, 1.0 / 355687428096000.0};
static const double CC[8] = {1.0 / 24.0, -1.0 / 720.0, 1.0 / 40320.0, -1.0 / 3628800.0, 1.0 / 479001600.0,
                             -1.0 / 87178291200.0, 1.0 / 20922789888000.0, -1.0 / 6402373705728000.0};

static void det_sincos(double x, double* sn, double* cs) {
    double k = rint(x * TWO_OVER_PI);
    double r = ((x - k * PIO2_1) - k * PIO2_2) - k * PIO2_3;
    double z = r * r;
    double ps = SC[7], pc = CC[7];
    for (int i = 6; i >= 0; --i) ps = ps * z + SC[i];
    for (int i = 6; i >= 0; --i) pc = pc * z + CC[i];
    double s = r + (r * z) * ps;
    double c = 1.0 - (0.5 * z - (z * z) * pc);
    int q = (int)(((int64_t)k) & 3);
    switch (q) {
        case 0: *sn = s; *cs = c; break;
        case 1: *sn = c; *cs = -s; break;
        case 2: *sn = -s; *cs = -c; break;
        default: *sn = -c; *cs = s; break;
    }
}

/* ------------------------------------------------------------------ rigid 3x4 helpers */
typedef struct { double r[9]; double t[3]; } Rt;      /* rotation row-major + translation */
typedef struct { double w, x, y, z; double t[3]; } Sq; /* Sophus::SE3d: unit quaternion + t */

static Rt rt_identity(void) {
    Rt I; memset(&I, 0, sizeof I); I.r[0] = I.r[4] = I.r[8] = 1.0; return I;
}
static Rt rt_from16(const double* m) {
    Rt T;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T.r[3 * i + j] = m[4 * i + j]; T.t[i] = m[4 * i + 3]; }
    return T;
}
static void rt_to16(const Rt* T, double* m) {
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) m[4 * i + j] = T->r[3 * i + j]; m[4 * i + 3] = T->t[i]; }
    m[12] = m[13] = m[14] = 0.0; m[15] = 1.0;
}
static Rt rt_mul(const Rt* A, const Rt* B) {
    Rt C;
    for (int i = 0; i < 3; ++i) {
        double a0 = A->r[3 * i], a1 = A->r[3 * i + 1], a2 = A->r[3 * i + 2];
        for (int j = 0; j < 3; ++j) C.r[3 * i + j] = (a0 * B->r[j] + a1 * B->r[3 + j]) + a2 * B->r[6 + j];
        C.t[i] = ((a0 * B->t[0] + a1 * B->t[1]) + a2 * B->t[2]) + A->t[i];
    }
    return C;
}
static Rt rt_inv(const Rt* T) {
    Rt C;
    for (int i = 0; i < 3; ++i) {
        double r0 = T->r[i], r1 = T->r[3 + i], r2 = T->r[6 + i];
        C.r[3 * i] = r0; C.r[3 * i + 1] = r1; C.r[3 * i + 2] = r2;
        C.t[i] = -((r0 * T->t[0] + r1 * T->t[1]) + r2 * T->t[2]);
    }
    return C;
}
static inline void rt_apply(const Rt* T, double x, double y, double z, double* xo, double* yo, double* zo) {
    *xo = ((T->r[0] * x + T->r[1] * y) + T->r[2] * z) + T->t[0];
    *yo = ((T->r[3] * x + T->r[4] * y) + T->r[5] * z) + T->t[1];
    *zo = ((T->r[6] * x + T->r[7] * y) + T->r[8] * z) + T->t[2];
}

/* Eigen QuaternionBase::toRotationMatrix */
static void quat_to_rot(double qw, double qx, double qy, double qz, double* R) {
    double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

/* Eigen rotation matrix -> quaternion */
static void rot_to_quat(const double* m, double* w, double* q) {
    double t = (m[0] + m[4]) + m[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        *w = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t;
        q[1] = (m[2] - m[6]) * t;
        q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(((m[4 * i] - m[4 * j]) - m[4 * k]) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        *w = (m[3 * k + j] - m[3 * j + k]) * t;
        q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
        q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    }
}

static Sq sq_identity(void) { Sq s = {1.0, 0.0, 0.0, 0.0, {0.0, 0.0, 0.0}}; return s; }
static Sq sq_from_rt(const Rt* T) {
    Sq s; double q[3];
    rot_to_quat(T->r, &s.w, q);
    s.x = q[0]; s.y = q[1]; s.z = q[2];
    s.t[0] = T->t[0]; s.t[1] = T->t[1]; s.t[2] = T->t[2];
    return s;
}
/* Sophus SE3 product: quaternion product + first-order renormalisation; t = Ra tb + ta */
static Sq sq_mul(const Sq* a, const Sq* b) {
    Sq c;
    double w = ((a->w * b->w - a->x * b->x) - a->y * b->y) - a->z * b->z;
    double x = ((a->w * b->x + a->x * b->w) + a->y * b->z) - a->z * b->y;
    double y = ((a->w * b->y + a->y * b->w) + a->z * b->x) - a->x * b->z;
    double z = ((a->w * b->z + a->z * b->w) + a->x * b->y) - a->y * b->x;
    double n2 = ((w * w + x * x) + y * y) + z * z;
    if (n2 != 1.0) { double s = 2.0 / (1.0 + n2); w = w * s; x = x * s; y = y * s; z = z * s; }
    c.w = w; c.x = x; c.y = y; c.z = z;
    double R[9];
    quat_to_rot(a->w, a->x, a->y, a->z, R);
    for (int i = 0; i < 3; ++i) c.t[i] = ((R[3 * i] * b->t[0] + R[3 * i + 1] * b->t[1]) + R[3 * i + 2] * b->t[2]) + a->t[i];
    return c;
}
static Rt sq_matrix(const Sq* s) {
    Rt T;
    quat_to_rot(s->w, s->x, s->y, s->z, T.r);
    T.t[0] = s->t[0]; T.t[1] = s->t[1]; T.t[2] = s->t[2];
    return T;
}

/* Sophus SE3::exp, tangent = [upsilon, omega]; optionally returns the quaternion */
static Rt se3_exp(const double* tg, Sq* qout) {
    double ux = tg[0], uy = tg[1], uz = tg[2], wx = tg[3], wy = tg[4], wz = tg[5];
    double theta_sq = (wx * wx + wy * wy) + wz * wz;
    int small = theta_sq < KP_EPS * KP_EPS;
    double theta = small ? 0.0 : sqrt(theta_sq);
    double half = 0.5 * theta;
    double sh, ch;
    det_sincos(half, &sh, &ch);
    double imag, real;
    if (small) {
        double po4 = theta_sq * theta_sq;
        imag = (0.5 - (1.0 / 48.0) * theta_sq) + (1.0 / 3840.0) * po4;
        real = (1.0 - (1.0 / 8.0) * theta_sq) + (1.0 / 384.0) * po4;
    } else {
        imag = sh / theta;
        real = ch;
    }
    double qw = real, qx = imag * wx, qy = imag * wy, qz = imag * wz;
    Rt T;
    quat_to_rot(qw, qx, qy, qz, T.r);
    double V[9];
    if (theta >= KP_EPS) {
        double st, ct;
        det_sincos(theta, &st, &ct);
        double tsq = theta * theta;
        double a = (1.0 - ct) / tsq;
        double b = (theta - st) / (tsq * theta);
        double o00 = -(wy * wy + wz * wz), o11 = -(wx * wx + wz * wz), o22 = -(wx * wx + wy * wy);
        double o01 = wx * wy, o02 = wx * wz, o12 = wy * wz;
        V[0] = 1.0 + b * o00;        V[1] = a * (-wz) + b * o01;  V[2] = a * wy + b * o02;
        V[3] = a * wz + b * o01;     V[4] = 1.0 + b * o11;        V[5] = a * (-wx) + b * o12;
        V[6] = a * (-wy) + b * o02;  V[7] = a * wx + b * o12;     V[8] = 1.0 + b * o22;
    } else {
        memcpy(V, T.r, sizeof V);
    }
    for (int i = 0; i < 3; ++i) T.t[i] = (V[3 * i] * ux + V[3 * i + 1] * uy) + V[3 * i + 2] * uz;
    if (qout) { qout->w = qw; qout->x = qx; qout->y = qy; qout->z = qz; qout->t[0] = T.t[0]; qout->t[1] = T.t[1]; qout->t[2] = T.t[2]; }
    return T;
}

/* Sophus SO3::logAndTheta on the quaternion of R (host scalars: libm) */
static void so3_log(const double* R, double* om, double* theta_out) {
    double w, q[3];
    rot_to_quat(R, &w, q);
    double sq_n = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
    double two_atan, theta;
    if (sq_n < KP_EPS * KP_EPS) {
        double sq_w = w * w;
        two_atan = 2.0 / w - (2.0 / 3.0) * sq_n / (w * sq_w);
        theta = 2.0 * sq_n / w;
    } else {
        double n = sqrt(sq_n);
        double at = w < 0.0 ? atan2(-n, -w) : atan2(n, w);
        two_atan = 2.0 * at / n;
        theta = two_atan * n;
    }
    om[0] = two_atan * q[0]; om[1] = two_atan * q[1]; om[2] = two_atan * q[2];
    *theta_out = theta;
}

static void se3_log(const Rt* T, double* out6) {
    double om[3], theta;
    so3_log(T->r, om, &theta);
    double wx = om[0], wy = om[1], wz = om[2];
    double o00 = -(wy * wy + wz * wz), o11 = -(wx * wx + wz * wz), o22 = -(wx * wx + wy * wy);
    double o01 = wx * wy, o02 = wx * wz, o12 = wy * wz;
    double c;
    if (fabs(theta) < KP_EPS) c = 1.0 / 12.0;
    else { double half = 0.5 * theta; c = (1.0 - (theta * cos(half)) / (2.0 * sin(half))) / (theta * theta); }
    double v[9] = {1.0 + c * o00, 0.5 * wz + c * o01, -0.5 * wy + c * o02,
                   -0.5 * wz + c * o01, 1.0 + c * o11, 0.5 * wx + c * o12,
                   0.5 * wy + c * o02, -0.5 * wx + c * o12, 1.0 + c * o22};
    for (int i = 0; i < 3; ++i) out6[i] = (v[3 * i] * T->t[0] + v[3 * i + 1] * T->t[1]) + v[3 * i + 2] * T->t[2];
    out6[3] = wx; out6[4] = wy; out6[5] = wz;
}

static double rot_angle(const double* R) {   /* Eigen::AngleAxisd(R).angle() */
    double w, q[3];
    rot_to_quat(R, &w, q);
    double n = sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
    return 2.0 * atan2(n, fabs(w));
}

/* ------------------------------------------------------------------ voxel keys + hash set */
static inline int key_ok(int kx, int ky, int kz) {
    return abs(kx) < KEY_BIAS && abs(ky) < KEY_BIAS && abs(kz) < KEY_BIAS;
}
static inline u64 pack_key(int kx, int ky, int kz) {
    return ((u64)(uint32_t)(kx + KEY_BIAS) << 42) | ((u64)(uint32_t)(ky + KEY_BIAS) << 21) | (u64)(uint32_t)(kz + KEY_BIAS);
}
static inline u64 point_key(double x, double y, double z, double size) {
    /* (p / size).cast<int>(): truncation toward zero */
    return pack_key((int)(x / size), (int)(y / size), (int)(z / size));
}
static inline uint32_t mix(u64 k) {
    k ^= k >> 31; k *= 0x9E3779B97F4A7C15ull; k ^= k >> 29;
    return (uint32_t)k;
}

typedef struct { u64* keys; int* vals; uint32_t mask; int used; } Table;
#define T_EMPTY (~(u64)0)

static void table_init(Table* t, uint32_t cap_pow2) {
    t->keys = (u64*)malloc((size_t)cap_pow2 * sizeof(u64));
    t->vals = (int*)malloc((size_t)cap_pow2 * sizeof(int));
    memset(t->keys, 0xFF, (size_t)cap_pow2 * sizeof(u64));
    t->mask = cap_pow2 - 1; t->used = 0;
}
static void table_free(Table* t) { free(t->keys); free(t->vals); t->keys = NULL; t->vals = NULL; }
static void table_clear(Table* t) { memset(t->keys, 0xFF, ((size_t)t->mask + 1) * sizeof(u64)); t->used = 0; }
static void table_grow(Table* t) {
    Table n; table_init(&n, (t->mask + 1) * 2);
    for (uint32_t s = 0; s <= t->mask; ++s) {
        if (t->keys[s] == T_EMPTY) continue;
        uint32_t p = mix(t->keys[s]) & n.mask;
        while (n.keys[p] != T_EMPTY) p = (p + 1) & n.mask;
        n.keys[p] = t->keys[s]; n.vals[p] = t->vals[s];
    }
    n.used = t->used;
    table_free(t); *t = n;
}
/* find: value or -1 */
static inline int table_find(const Table* t, u64 key) {
    uint32_t p = mix(key) & t->mask;
    while (1) {
        u64 k = t->keys[p];
        if (k == key) return t->vals[p];
        if (k == T_EMPTY) return -1;
        p = (p + 1) & t->mask;
    }
}
/* insert if absent; returns existing value, or -1 after inserting val */
static inline int table_insert(Table* t, u64 key, int val) {
    if ((size_t)(t->used + 1) * 2 > (size_t)t->mask + 1) table_grow(t);
    uint32_t p = mix(key) & t->mask;
    while (1) {
        u64 k = t->keys[p];
        if (k == key) return t->vals[p];
        if (k == T_EMPTY) { t->keys[p] = key; t->vals[p] = val; t->used++; return -1; }
        p = (p + 1) & t->mask;
    }
}

/* ------------------------------------------------------------------ context */
typedef struct { u64 key; int count; double p[KP_MAXP][3]; } Voxel;

typedef struct kp_stats {
    int status, n_in, n_range, n_ds, n_src, n_voxels, iterations, n_corr;
    double dx_norm, sigma, err_dt, err_drot;
    int map_points, reserved;
} kp_stats;

typedef struct kp_ctx {
    double max_range, min_range, voxel_size, initial_threshold, min_motion_th;
    int maxp, deskew, max_iters, threads;
    double eps;
    /* KissICP state */
    Rt* poses; int n_poses, cap_poses;
    double sse2; int num_samples; Rt deviation;
    double last_sigma;
    /* local map */
    Table map; Voxel* vox; int n_vox, cap_vox;
    /* scratch */
    Table grid;
    double *frame, *ds, *src0, *src, *tgt, *terms;
    int *ds_idx, *src_idx, *order; unsigned char* acc;
    int cap_pts, n_frame, n_ds, n_src;
    /* trace of the last registration: per iteration order ids */
    int trace_iters; int* trace; int trace_n, trace_used;
} kp_ctx;

static void ensure_points(kp_ctx* c, int n) {
    if (n <= c->cap_pts) return;
    int cap = n + n / 4 + 1024;
    c->frame = (double*)realloc(c->frame, (size_t)cap * 3 * sizeof(double));
    c->ds = (double*)realloc(c->ds, (size_t)cap * 3 * sizeof(double));
    c->src0 = (double*)realloc(c->src0, (size_t)cap * 3 * sizeof(double));
    c->src = (double*)realloc(c->src, (size_t)cap * 3 * sizeof(double));
    c->tgt = (double*)realloc(c->tgt, (size_t)cap * 3 * sizeof(double));
    c->ds_idx = (int*)realloc(c->ds_idx, (size_t)cap * sizeof(int));
    c->src_idx = (int*)realloc(c->src_idx, (size_t)cap * sizeof(int));
    c->order = (int*)realloc(c->order, (size_t)cap * sizeof(int));
    c->acc = (unsigned char*)realloc(c->acc, (size_t)cap);
    c->cap_pts = cap;
}

kp_ctx* kp_create(double max_range, double min_range, double voxel_size, int max_points_per_voxel, int deskew,
                  double initial_threshold, double min_motion_th, int threads, int trace_iters) {
    kp_ctx* c = (kp_ctx*)calloc(1, sizeof(kp_ctx));
    c->max_range = max_range; c->min_range = min_range;
    c->voxel_size = voxel_size > 0.0 ? voxel_size : max_range / 100.0;
    c->maxp = max_points_per_voxel; c->deskew = deskew;
    c->initial_threshold = initial_threshold; c->min_motion_th = min_motion_th;
    c->max_iters = 500; c->eps = 1e-4;
    c->threads = threads > 0 ? threads : 1;
    c->deviation = rt_identity();
    table_init(&c->map, 1u << 16);
    table_init(&c->grid, 1u << 16);
    c->cap_vox = 1 << 14;
    c->vox = (Voxel*)malloc((size_t)c->cap_vox * sizeof(Voxel));
    c->trace_iters = trace_iters;
    return c;
}

void kp_destroy(kp_ctx* c) {
    if (!c) return;
    table_free(&c->map); table_free(&c->grid);
    free(c->vox); free(c->poses); free(c->frame); free(c->ds); free(c->src0); free(c->src); free(c->tgt);
    free(c->terms); free(c->ds_idx); free(c->src_idx); free(c->order); free(c->acc); free(c->trace);
    free(c);
}

void kp_set_iteration_limits(kp_ctx* c, int max_iters, double eps) { c->max_iters = max_iters; c->eps = eps; }
void kp_set_threads(kp_ctx* c, int threads) { c->threads = threads > 0 ? threads : 1; }
int kp_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ A.2 deskew, A.4 range, A.5 grid */
/* kiss-icp DeSkewScan (kiss.py:90): p_i <- exp((t_i - 0.5) * delta) p_i; parallel over points */
static void deskew_points(const double* xyz, const double* ts, int n, const double* delta, double* out, int threads) {
    (void)threads;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < n; ++i) {
        double s = ts[i] - 0.5;
        double tg[6];
        for (int k = 0; k < 6; ++k) tg[k] = s * delta[k];
        Rt M = se3_exp(tg, NULL);
        rt_apply(&M, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &out[3 * i], &out[3 * i + 1], &out[3 * i + 2]);
    }
}

int kp_deskew_scan(kp_ctx* c, const double* xyz, const double* ts, int n, const double* start16, const double* finish16,
                   double* out) {
    Rt a = rt_from16(start16), b = rt_from16(finish16), ai = rt_inv(&a);
    Rt rel = rt_mul(&ai, &b);
    double delta[6];
    se3_log(&rel, delta);
    deskew_points(xyz, ts, n, delta, out, c->threads);
    return 0;
}

/* kiss-icp Preprocess (kiss.py:93): keep min < |p| < max, order preserved; serial copy_if */
int kp_preprocess(const double* xyz, int n, double max_range, double min_range, double* out) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        double nrm = sqrt((x * x + y * y) + z * z);
        if (nrm < max_range && nrm > min_range) { out[3 * m] = x; out[3 * m + 1] = y; out[3 * m + 2] = z; ++m; }
    }
    return m;
}

/* kiss-icp VoxelDownsample (kiss.py:96): first point per voxel; output in input order (B.1).
 * returns count or -1 when a voxel coordinate is out of range */
static int downsample(Table* grid, const double* xyz, int n, double size, double* out, int* out_idx) {
    table_clear(grid);
    int m = 0;
    for (int i = 0; i < n; ++i) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        int kx = (int)(x / size), ky = (int)(y / size), kz = (int)(z / size);
        if (!key_ok(kx, ky, kz)) return -1;
        if (table_insert(grid, pack_key(kx, ky, kz), i) < 0) {
            out[3 * m] = x; out[3 * m + 1] = y; out[3 * m + 2] = z;
            if (out_idx) out_idx[m] = i;
            ++m;
        }
    }
    return m;
}

int kp_voxel_down_sample(kp_ctx* c, const double* xyz, int n, double voxel_size, double* out, int* out_idx) {
    return downsample(&c->grid, xyz, n, voxel_size, out, out_idx);
}

/* ------------------------------------------------------------------ A.6 VoxelHashMap */
static void map_add_points(kp_ctx* c, const double* pts, int n) {
    for (int i = 0; i < n; ++i) {
        double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        u64 key = point_key(x, y, z, c->voxel_size);
        int v = table_insert(&c->map, key, c->n_vox);
        if (v < 0) {
            if (c->n_vox == c->cap_vox) {
                c->cap_vox *= 2;
                c->vox = (Voxel*)realloc(c->vox, (size_t)c->cap_vox * sizeof(Voxel));
            }
            v = c->n_vox++;
            c->vox[v].key = key; c->vox[v].count = 0;
        }
        Voxel* V = &c->vox[v];
        if (V->count < c->maxp) { V->p[V->count][0] = x; V->p[V->count][1] = y; V->p[V->count][2] = z; V->count++; }
    }
}

/* RemovePointsFarFromLocation with rule B.4: erase EVERY voxel whose first point is too far */
static void map_remove_far(kp_ctx* c, const double* o) {
    double r2 = c->max_range * c->max_range;
    int w = 0, removed = 0;
    for (int v = 0; v < c->n_vox; ++v) {
        const Voxel* V = &c->vox[v];
        double dx = V->p[0][0] - o[0], dy = V->p[0][1] - o[1], dz = V->p[0][2] - o[2];
        double d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 > r2) { ++removed; continue; }
        if (w != v) c->vox[w] = c->vox[v];
        ++w;
    }
    if (!removed) return;
    c->n_vox = w;
    table_clear(&c->map);
    for (int v = 0; v < c->n_vox; ++v) table_insert(&c->map, c->vox[v].key, v);
}

void kp_map_clear(kp_ctx* c) { table_clear(&c->map); c->n_vox = 0; }
int kp_map_num_voxels(const kp_ctx* c) { return c->n_vox; }
int kp_map_num_points(const kp_ctx* c) {
    int m = 0;
    for (int v = 0; v < c->n_vox; ++v) m += c->vox[v].count;
    return m;
}
void kp_map_add_points(kp_ctx* c, const double* pts, int n) { map_add_points(c, pts, n); }
void kp_map_remove_far(kp_ctx* c, const double* origin3) { map_remove_far(c, origin3); }
void kp_map_update(kp_ctx* c, const double* pts, int n, const double* pose16) {
    Rt T = rt_from16(pose16);
    double* w = (double*)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(double));
    for (int i = 0; i < n; ++i) rt_apply(&T, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], &w[3 * i], &w[3 * i + 1], &w[3 * i + 2]);
    map_add_points(c, w, n);
    map_remove_far(c, T.t);
    free(w);
}

static int cmp_u64(const void* a, const void* b) {
    u64 x = ((const u64*)a)[0], y = ((const u64*)b)[0];
    return x < y ? -1 : (x > y ? 1 : 0);
}
/* dump sorted by packed key: keys (V,3) int32, counts (V), points (V,20,3) zero padded */
int kp_map_dump(const kp_ctx* c, int* keys, int* counts, double* points, int capacity) {
    int V = c->n_vox;
    if (V > capacity) return -1;
    u64* idx = (u64*)malloc((size_t)(V > 0 ? V : 1) * 2 * sizeof(u64));
    for (int v = 0; v < V; ++v) { idx[2 * v] = c->vox[v].key; idx[2 * v + 1] = (u64)v; }
    qsort(idx, (size_t)V, 2 * sizeof(u64), cmp_u64);
    for (int r = 0; r < V; ++r) {
        const Voxel* X = &c->vox[idx[2 * r + 1]];
        keys[3 * r] = (int)((X->key >> 42) & 0x1FFFFF) - KEY_BIAS;
        keys[3 * r + 1] = (int)((X->key >> 21) & 0x1FFFFF) - KEY_BIAS;
        keys[3 * r + 2] = (int)(X->key & 0x1FFFFF) - KEY_BIAS;
        counts[r] = X->count;
        for (int s = 0; s < KP_MAXP; ++s)
            for (int a = 0; a < 3; ++a) points[((size_t)r * KP_MAXP + s) * 3 + a] = s < X->count ? X->p[s][a] : 0.0;
    }
    free(idx);
    return V;
}

/* A.7 GetCorrespondences for one query: 27 voxels, i outermost / l innermost, stored order,
 * strict '<' so the first candidate wins ties (B.6).  Returns the order id or -1 (B.3). */
static inline int nearest_in_map(const kp_ctx* c, double qx, double qy, double qz, double* best, double* bd2) {
    double v = c->voxel_size;
    int kx = (int)(qx / v), ky = (int)(qy / v), kz = (int)(qz / v);
    double d2min = INFINITY;
    int ord = -1, o = 0;
    for (int i = -1; i <= 1; ++i)
        for (int j = -1; j <= 1; ++j)
            for (int l = -1; l <= 1; ++l, ++o) {
                int nx = kx + i, ny = ky + j, nz = kz + l;
                if (!key_ok(nx, ny, nz)) continue;
                int id = table_find(&c->map, pack_key(nx, ny, nz));
                if (id < 0) continue;
                const Voxel* V = &c->vox[id];
                for (int s = 0; s < V->count; ++s) {
                    double dx = V->p[s][0] - qx, dy = V->p[s][1] - qy, dz = V->p[s][2] - qz;
                    double d2 = (dx * dx + dy * dy) + dz * dz;
                    if (d2 < d2min) { d2min = d2; ord = o * c->maxp + s; best[0] = V->p[s][0]; best[1] = V->p[s][1]; best[2] = V->p[s][2]; }
                }
            }
    *bd2 = d2min;
    return ord;
}

static int correspondences(const kp_ctx* c, const double* q, int n, double max_dist, int* order, double* tgt,
                           unsigned char* acc) {
    int ncorr = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ncorr) num_threads(c->threads)
    for (int i = 0; i < n; ++i) {
        double b[3] = {0, 0, 0}, d2;
        int ord = c->n_vox ? nearest_in_map(c, q[3 * i], q[3 * i + 1], q[3 * i + 2], b, &d2) : -1;
        int ok = ord >= 0 && sqrt(d2) < max_dist;
        order[i] = ok ? ord : -1;
        acc[i] = (unsigned char)ok;
        tgt[3 * i] = ok ? b[0] : 0.0; tgt[3 * i + 1] = ok ? b[1] : 0.0; tgt[3 * i + 2] = ok ? b[2] : 0.0;
        ncorr += ok;
    }
    return ncorr;
}

int kp_map_get_correspondences(kp_ctx* c, const double* q, int n, double max_dist, int* order, double* tgt) {
    ensure_points(c, n);
    return correspondences(c, q, n, max_dist, order, tgt, c->acc);
}

/* ------------------------------------------------------------------ A.8 registration */
#define NT 27
/* the 27 per-point terms: 21 upper-triangular JtJ entries (row-major) + 6 Jtr */
static inline void point_terms(double sx, double sy, double sz, double tx, double ty, double tz, double kernel, double* c) {
    double rx = sx - tx, ry = sy - ty, rz = sz - tz;
    double r2 = (rx * rx + ry * ry) + rz * rz;
    double kk = kernel + r2;
    double w = (kernel * kernel) / (kk * kk);
    double wsx = w * sx, wsy = w * sy, wsz = w * sz;
    double wrx = w * rx, wry = w * ry, wrz = w * rz;
    c[0] = w; c[1] = 0.0; c[2] = 0.0; c[3] = 0.0; c[4] = wsz; c[5] = -wsy;
    c[6] = w; c[7] = 0.0; c[8] = -wsz; c[9] = 0.0; c[10] = wsx;
    c[11] = w; c[12] = wsy; c[13] = -wsx; c[14] = 0.0;
    c[15] = w * (sy * sy + sz * sz); c[16] = -(w * (sx * sy)); c[17] = -(w * (sx * sz));
    c[18] = w * (sx * sx + sz * sz); c[19] = -(w * (sy * sz));
    c[20] = w * (sx * sx + sy * sy);
    c[21] = wrx; c[22] = wry; c[23] = wrz;
    c[24] = sy * wrz - sz * wry; c[25] = sz * wrx - sx * wrz; c[26] = sx * wry - sy * wrx;
}

/* canonical reduction (B.7): adjacent-pairs binary tree over the source index, zero padded to a
 * power of two >= 32.  `terms` holds p rows of NT doubles and is consumed. */
static void tree_sum(double* terms, int p, double* out, int threads) {
    (void)threads;
    for (int stride = 1; stride < p; stride <<= 1) {
        int pairs = p / (2 * stride);
#pragma omp parallel for schedule(static) num_threads(threads) if (pairs >= 1024)
        for (int k = 0; k < pairs; ++k) {
            double* a = terms + (size_t)(2 * k) * stride * NT;
            const double* b = a + (size_t)stride * NT;
            for (int t = 0; t < NT; ++t) a[t] = a[t] + b[t];
        }
    }
    memcpy(out, terms, NT * sizeof(double));
}

/* LDL^T with diagonal pivoting (oracle/canon.py ldlt_solve6, op for op) */
static int ldlt_solve6(const double* A, const double* b, double* x) {
    double a[6][6];
    int perm[6];
    for (int i = 0; i < 6; ++i) { perm[i] = i; for (int j = 0; j < 6; ++j) a[i][j] = A[6 * i + j]; }
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double best = fabs(a[k][k]);
        for (int i = k + 1; i < 6; ++i) { double v = fabs(a[i][i]); if (v > best) { best = v; p = i; } }
        if (p != k) {
            for (int cidx = 0; cidx < 6; ++cidx) { double t = a[k][cidx]; a[k][cidx] = a[p][cidx]; a[p][cidx] = t; }
            for (int r = 0; r < 6; ++r) { double t = a[r][k]; a[r][k] = a[r][p]; a[r][p] = t; }
            int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        double d = a[k][k];
        if (d == 0.0 || d != d) return 0;
        for (int i = k + 1; i < 6; ++i) a[i][k] = a[i][k] / d;
        for (int j = k + 1; j < 6; ++j) {
            double ljd = a[j][k] * d;
            for (int i = j; i < 6; ++i) { a[i][j] = a[i][j] - a[i][k] * ljd; a[j][i] = a[i][j]; }
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < 6; ++i) { double s = y[i]; for (int j = 0; j < i; ++j) s = s - a[i][j] * y[j]; y[i] = s; }
    for (int i = 0; i < 6; ++i) y[i] = y[i] / a[i][i];
    for (int i = 5; i >= 0; --i) { double s = y[i]; for (int j = i + 1; j < 6; ++j) s = s - a[j][i] * y[j]; y[i] = s; }
    int ok = 1;
    for (int i = 0; i < 6; ++i) { x[perm[i]] = y[i]; if (!(fabs(y[i]) <= 1.7976931348623157e308)) ok = 0; }
    return ok;
}

/* kiss-icp RegisterFrame (kiss.py:108-114).  `pts` sensor-frame source. */
static Rt register_points(kp_ctx* c, const double* pts, int n, const Rt* guess, double max_dist, double kernel, kp_stats* st) {
    Sq gq = sq_from_rt(guess), Ticp = sq_identity();
    st->iterations = 0; st->n_corr = 0; st->dx_norm = 0.0; st->status = 0;
    c->trace_used = 0; c->trace_n = n;
    if (c->n_vox == 0) { Sq r = sq_mul(&Ticp, &gq); return sq_matrix(&r); }
    ensure_points(c, n);
    double* src = c->src;
    for (int i = 0; i < n; ++i) rt_apply(guess, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], &src[3 * i], &src[3 * i + 1], &src[3 * i + 2]);
    int p = 32;
    while (p < n) p *= 2;
    c->terms = (double*)realloc(c->terms, (size_t)p * NT * sizeof(double));
    if (c->trace_iters > 0) c->trace = (int*)realloc(c->trace, (size_t)c->trace_iters * (size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int it = 0; it < c->max_iters; ++it) {
        int ncorr = correspondences(c, src, n, max_dist, c->order, c->tgt, c->acc);
        st->iterations = it + 1; st->n_corr = ncorr;
        if (it < c->trace_iters) { memcpy(c->trace + (size_t)it * n, c->order, (size_t)n * sizeof(int)); c->trace_used = it + 1; }
        if (ncorr == 0) { st->status = 1; break; }          /* B.5 */
        double* terms = c->terms;
#pragma omp parallel for schedule(static) num_threads(c->threads)
        for (int i = 0; i < p; ++i) {
            double* t = terms + (size_t)i * NT;
            if (i < n && c->acc[i]) point_terms(src[3 * i], src[3 * i + 1], src[3 * i + 2], c->tgt[3 * i], c->tgt[3 * i + 1], c->tgt[3 * i + 2], kernel, t);
            else for (int k = 0; k < NT; ++k) t[k] = 0.0;
        }
        double sums[NT];
        tree_sum(terms, p, sums, c->threads);
        double A[36], b[6], dx[6];
        int idx = 0;
        for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { A[6 * i + j] = sums[idx]; A[6 * j + i] = sums[idx]; ++idx; }
        for (int i = 0; i < 6; ++i) b[i] = -sums[21 + i];
        if (!ldlt_solve6(A, b, dx)) { st->status = 2; break; }
        Sq Eq;
        Rt E = se3_exp(dx, &Eq);
#pragma omp parallel for schedule(static) num_threads(c->threads) if (n >= 4096)
        for (int i = 0; i < n; ++i) {
            double x, y, z;
            rt_apply(&E, src[3 * i], src[3 * i + 1], src[3 * i + 2], &x, &y, &z);
            src[3 * i] = x; src[3 * i + 1] = y; src[3 * i + 2] = z;
        }
        Ticp = sq_mul(&Eq, &Ticp);
        double nrm = sqrt(((((dx[0] * dx[0] + dx[1] * dx[1]) + dx[2] * dx[2]) + dx[3] * dx[3]) + dx[4] * dx[4]) + dx[5] * dx[5]);
        st->dx_norm = nrm;
        if (nrm < c->eps) break;
    }
    Sq r = sq_mul(&Ticp, &gq);
    return sq_matrix(&r);
}

int kp_register_point_cloud(kp_ctx* c, const double* pts, int n, const double* guess16, double max_dist, double kernel,
                            double* out16, kp_stats* st) {
    kp_stats tmp;
    if (!st) st = &tmp;
    memset(st, 0, sizeof *st);
    Rt g = rt_from16(guess16);
    Rt r = register_points(c, pts, n, &g, max_dist, kernel, st);
    rt_to16(&r, out16);
    st->n_in = n; st->n_src = n; st->n_voxels = c->n_vox;
    return st->status == 2 ? -5 : 0;
}

/* ------------------------------------------------------------------ A.9 threshold, prediction */
static int has_moved(const kp_ctx* c) {
    if (c->n_poses < 1) return 0;
    Rt i0 = rt_inv(&c->poses[0]);
    Rt d = rt_mul(&i0, &c->poses[c->n_poses - 1]);
    double motion = sqrt((d.t[0] * d.t[0] + d.t[1] * d.t[1]) + d.t[2] * d.t[2]);
    return motion > 5.0 * c->min_motion_th;
}
static double compute_threshold(kp_ctx* c) {
    double theta = rot_angle(c->deviation.r);
    double delta_rot = 2.0 * c->max_range * sin(theta / 2.0);
    const double* t = c->deviation.t;
    double delta_trans = sqrt((t[0] * t[0] + t[1] * t[1]) + t[2] * t[2]);
    double err = delta_trans + delta_rot;
    if (err > c->min_motion_th) { c->sse2 += err * err; c->num_samples += 1; }
    if (c->num_samples < 1) return c->initial_threshold;
    return sqrt(c->sse2 / c->num_samples);
}
static Rt prediction_model(const kp_ctx* c) {
    if (c->n_poses < 2) return rt_identity();
    Rt i2 = rt_inv(&c->poses[c->n_poses - 2]);
    return rt_mul(&i2, &c->poses[c->n_poses - 1]);
}

void kp_reset(kp_ctx* c) {
    c->n_poses = 0; c->sse2 = 0.0; c->num_samples = 0; c->deviation = rt_identity(); c->last_sigma = 0.0;
    kp_map_clear(c);
    c->n_frame = c->n_ds = c->n_src = 0;
}
int kp_num_poses(const kp_ctx* c) { return c->n_poses; }
int kp_get_pose(const kp_ctx* c, int index, double* out16) {
    if (index < 0) index += c->n_poses;
    if (index < 0 || index >= c->n_poses) return -1;
    rt_to16(&c->poses[index], out16);
    return 0;
}
void kp_get_prediction_model(const kp_ctx* c, double* out16) { Rt p = prediction_model(c); rt_to16(&p, out16); }

/* ------------------------------------------------------------------ the step: kiss.py:83-131 */
int kp_register_frame(kp_ctx* c, const double* xyz, const double* ts, int n, const double* guess16, double* out16,
                      kp_stats* st) {
    kp_stats tmp;
    if (!st) st = &tmp;
    memset(st, 0, sizeof *st);
    ensure_points(c, n);
    /* deskew (kiss.py:90): identity with < 2 poses */
    const double* frame_in = xyz;
    if (c->deskew && c->n_poses >= 2) {
        Rt rel = prediction_model(c);
        double delta[6];
        se3_log(&rel, delta);
        deskew_points(xyz, ts, n, delta, c->src0, c->threads);   /* src0 as scratch */
        frame_in = c->src0;
    }
    /* preprocess (kiss.py:93) */
    int nr = kp_preprocess(frame_in, n, c->max_range, c->min_range, c->frame);
    c->n_frame = nr;
    /* voxelize (kiss.py:96) */
    int nd = downsample(&c->grid, c->frame, nr, c->voxel_size * 0.5, c->ds, c->ds_idx);
    if (nd < 0) return -4;
    int ns = downsample(&c->grid, c->ds, nd, c->voxel_size * 1.5, c->src0, c->src_idx);
    if (ns < 0) return -4;
    c->n_ds = nd; c->n_src = ns;
    /* adaptive threshold (kiss.py:99) */
    double sigma = has_moved(c) ? compute_threshold(c) : c->initial_threshold;
    /* initial guess (kiss.py:102-105) */
    Rt guess;
    if (guess16) guess = rt_from16(guess16);
    else {
        Rt pred = prediction_model(c);
        Rt last = c->n_poses ? c->poses[c->n_poses - 1] : rt_identity();
        guess = rt_mul(&last, &pred);
    }
    /* register (kiss.py:108-114) */
    Rt pose = register_points(c, c->src0, ns, &guess, 3 * sigma, sigma / 3, st);
    /* pose gain metrics (kiss.py:116-124) + model deviation (kiss.py:128) */
    Rt gi = rt_inv(&guess);
    Rt gain = rt_mul(&gi, &pose);
    double om[3], theta;
    so3_log(gain.r, om, &theta);
    st->err_dt = sqrt((gain.t[0] * gain.t[0] + gain.t[1] * gain.t[1]) + gain.t[2] * gain.t[2]);
    st->err_drot = fabs(theta);
    st->sigma = sigma;
    c->deviation = gain;
    c->last_sigma = sigma;
    /* local_map.update(frame_downsample, new_pose) (kiss.py:129) */
    double* w = c->tgt;   /* scratch, nd <= cap */
    for (int i = 0; i < nd; ++i) rt_apply(&pose, c->ds[3 * i], c->ds[3 * i + 1], c->ds[3 * i + 2], &w[3 * i], &w[3 * i + 1], &w[3 * i + 2]);
    map_add_points(c, w, nd);
    map_remove_far(c, pose.t);
    /* poses.append (kiss.py:130) */
    if (c->n_poses == c->cap_poses) {
        c->cap_poses = c->cap_poses ? c->cap_poses * 2 : 256;
        c->poses = (Rt*)realloc(c->poses, (size_t)c->cap_poses * sizeof(Rt));
    }
    c->poses[c->n_poses++] = pose;
    rt_to16(&pose, out16);
    st->n_in = n; st->n_range = nr; st->n_ds = nd; st->n_src = ns; st->n_voxels = c->n_vox;
    st->map_points = kp_map_num_points(c);
    return st->status == 2 ? -5 : 0;
}

/* taps on the last step: which 0 = frame_downsample, 1 = source (sensor frame), 2 = preprocessed frame */
int kp_get_points(const kp_ctx* c, int which, double* out, int* out_idx, int capacity) {
    const double* p = which == 0 ? c->ds : (which == 1 ? c->src0 : c->frame);
    const int* ix = which == 0 ? c->ds_idx : (which == 1 ? c->src_idx : NULL);
    int m = which == 0 ? c->n_ds : (which == 1 ? c->n_src : c->n_frame);
    if (m > capacity) return -1;
    if (out) memcpy(out, p, (size_t)m * 3 * sizeof(double));
    if (out_idx && ix) memcpy(out_idx, ix, (size_t)m * sizeof(int));
    return m;
}
int kp_get_trace(const kp_ctx* c, int* out, int capacity_iters, int* n_src) {
    int iters = c->trace_used < capacity_iters ? c->trace_used : capacity_iters;
    if (n_src) *n_src = c->trace_n;
    if (out && iters > 0) memcpy(out, c->trace, (size_t)iters * c->trace_n * sizeof(int));
    return c->trace_used;
}

/* canon taps for the cross-checks in tests/ */
void kp_det_sincos(const double* x, int n, double* s, double* c) { for (int i = 0; i < n; ++i) det_sincos(x[i], &s[i], &c[i]); }
void kp_se3_exp(const double* tangent6, double* out16) { Rt T = se3_exp(tangent6, NULL); rt_to16(&T, out16); }
void kp_se3_log(const double* pose16, double* out6) { Rt T = rt_from16(pose16); se3_log(&T, out6); }
