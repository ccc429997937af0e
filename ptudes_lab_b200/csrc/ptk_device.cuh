// Device-side data layout and kernels of the odometry step (sm_100a).
//
// Reference path being replaced: the kiss-icp 0.2.x C++ core reached from
// /root/reference/src/ptudes/kiss.py:90 (DeSkewScan), :93 (Preprocess), :96 (VoxelDownsample x2),
// :108-114 (RegisterFrame: GetCorrespondences + BuildLinearSystem + LDLT + SE3 exp) and :129
// (VoxelHashMap::Update = AddPoints + RemovePointsFarFromLocation).
//
// Everything is float64 / int32 / packed-u64 keys; the work is gather/scatter and small
// reductions, so the kernels are plain CUDA-core kernels: no tensor cores on purpose.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptk_canon.cuh"

namespace ptk {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int MAXP = 20;                   // points per voxel block (mapping.max_points_per_voxel)
constexpr u64 KEY_EMPTY = ~0ull;
constexpr u64 KEY_TOMB = ~0ull - 1ull;
constexpr u32 NONE = 0xFFFFFFFFu;
constexpr int KEY_BIAS = 1 << 20;
constexpr int TILE = 1024;                 // items per compaction tile (256 threads x 4)
constexpr int PTK_MAX_PEERS = 8;           // ranks of the peer-memory (in-kernel) exchange of the hash-sharded mode
constexpr int XCH_FLAGS = 64;              // stamps per source rank: one per k_icp block (< 63) + the voxel count
constexpr int NSUM = 17;                   // distinct normal-equation sums (16) + correspondence count
constexpr int NRED = NSUM;
#ifndef PTK_ICP_THREADS
#define PTK_ICP_THREADS 512
#endif
#ifndef PTK_ICP_MINBLOCKS
#define PTK_ICP_MINBLOCKS 2
#endif
constexpr int ICP_THREADS = PTK_ICP_THREADS;
constexpr int ICP_WARPS = ICP_THREADS / 32;
constexpr int ICP_CHUNK = ICP_WARPS;        // 32-point groups a block handles at a time: one point per thread
#ifndef PTK_ICP_KX
#define PTK_ICP_KX 2
#endif
constexpr int ICP_KX = PTK_ICP_KX;          // runner-ups a correspondence cache entry keeps beside the winner
#ifndef PTK_ICP_SRC_CAP
#define PTK_ICP_SRC_CAP (ICP_KX <= 1 ? 768 : (ICP_KX == 2 ? 640 : 512))
#endif
constexpr int ICP_SRC_CAP = PTK_ICP_SRC_CAP;   // source points (+ cache entries) per block in smem
constexpr int ICP_SMEM = ICP_SRC_CAP * (3 * 8 + 4 * 8 + 8 + 3 * 8 + ICP_KX * (3 * 8 + 4) + 4) + 8;

enum StepFlags : int { F_DESKEW = 1, F_RANGE = 2, F_SECOND = 4, F_SELECT_RANGE = 8 };
enum ErrFlags : int { ERR_KEYRANGE = 1, ERR_POOL = 2, ERR_TABLE = 4 };

// One voxel of the local map: 20 points SoA, 512 B, 16 B aligned so rows of x/y/z can be read with vector
// loads.  Unused slots hold +inf (the searches need no count).
struct __align__(16) VoxelBlock {
    double x[MAXP];
    double y[MAXP];
    double z[MAXP];
    u64 key;
    u32 pad[6];
};
static_assert(sizeof(VoxelBlock) == 512, "VoxelBlock must be 512 B");

// What map maintenance needs of a voxel, 32 B beside the block pool: its first point (RemovePointsFarFromLocation
// tests that one), the number of valid points and the voxel's table slot.  The prune pass streams this array - one
// sector per voxel - instead of touching four sectors of every 512 B block.
struct __align__(32) VoxelMeta {
    double fx, fy, fz;
    u32 count;
    u32 slot;
};
static_assert(sizeof(VoxelMeta) == 32, "VoxelMeta must be 32 B");

struct __align__(16) MapSlot {
    u64 key;
    u32 id;     // voxel block index, NONE while being created
    u32 pad;
};

// Per-step parameters, written by the host before the launches of one step.
struct StepParams {
    const double* xyz;     // (n,3) row-major, device
    const double* ts;      // (n), device (may be null when !F_DESKEW)
    int n;
    int flags;
    double delta[6];       // deskew twist log(inv(T[-2]) T[-1])
    double ds1_size;       // first grid
    double ds2_size;       // second grid
    double ds1_inv, ds2_inv;   // 1 / size (quotient estimates, see trunc_div)
    double max_range, min_range;
    Rigid guess;
    double max_corr;
    double kernel;
    u32 epoch;             // step counter, >= 1
    u32 tbase1, tbase2;    // ticket bases of the two compaction kernels
    u32 release_base;      // (unused)
    u32 xch_epoch;         // hash-sharded mode: step counter the exchange stamps are built from (same on every rank)
    u32 near1, near2;      // masks of the front regions of the two scan tables (0: whole table), see table_insert_min
    int W;                 // range-image mode: columns per frame (n = H * W pixels)
    // range-image input (kiss.py:59-61 on the device): non-null `range` selects it
    const u32* range;      // (H*W) RANGE field in millimetres, 0 = no return
    double range_unit;     // metres per RANGE count the direction LUT expects (0.001, or 1 if pre-scaled)
    const double* lut_dir; // [3][H*W] XYZLut direction, one plane per coordinate
    const double* lut_off; // [3][H*W] XYZLut offset or null
    const double* col_ts;  // (W) normalised column timestamps (kiss.py:34-35)
    double* col_motion;    // [12][W] deskew motion of every column, written by k_col_motion
};

struct StepOut {
    Rigid pose;
    double dx_norm;
    int status, n_range, n_ds, n_src, n_vox, n_tomb, iterations, n_corr, map_points, err, bump, icp_searches;
    int n_valid, pad;
    long long icp_cyc[6];      // block 0's clock64 spent in: cache pass, searches, sums, barrier, tree, solve
};

// Per-sequence ("lane") device state.
struct LaneDev {
    // config
    double voxel_size, max_distance, voxel_inv;
    int maxp, max_iters;
    double eps;
    int shard_rank, shard_n;             // hash-sharded map: this context keeps voxels with shard_owner(key) == shard_rank
    int cap_points, pool_cap, trace_iters, ng_cap;
    u32 t_mask, m_mask;
    // scan-local tables (first-seen selection), self-cleaning
    u64* t1_keys; u32* t1_vals;
    u64* t2_keys; u32* t2_vals;
    u32* slot1;                          // [cap_points] table-1 slot of every input point
    u64* agg1; u64* agg2;                // tile aggregates of the two compactions
    // frame_downsample (grid 0.5 v), sensor frame
    double *ds_x, *ds_y, *ds_z;
    u32 *ds_idx, *ds_slot2, *ds_vid;
    // source (grid 1.5 v): sensor frame + current (transformed) positions
    double *s0_x, *s0_y, *s0_z, *s_x, *s_y, *s_z;
    u32* s_idx;
    // local map
    MapSlot* m_slots;
    VoxelMeta* vmeta;                    // [pool_cap] first point, count, table slot of every voxel block
    VoxelBlock* blocks;
    u32* vidx;                           // [pool_cap*MAXP] insertion scratch, NONE when idle
    u32* freelist;
    // icp
    double* part_a; double* part_b;      // [NRED][ng_cap] group partial sums (part_b: unused spare)
    double *c_tx, *c_ty, *c_tz, *c_slack; // correspondence cache: winner, bound on every other candidate
    double *c_px, *c_py, *c_pz;
    double* c_t2;                        // [3 * ICP_KX][cap_points] runner-up coordinates
    int* c_ord2;                         // [ICP_KX][cap_points]
    u64* c_key; int* c_ord;
    int* trace;                          // [trace_iters][cap_points]
    // hash-sharded mode, exchange through peer memory: entry r = rank r's buffer as mapped into this process
    double* xch_rec[PTK_MAX_PEERS];              // [source rank][2][5][cap_points]: d2, order id, target xyz per source point
    unsigned long long* xch_flag[PTK_MAX_PEERS]; // [source rank][XCH_FLAGS] stamps (epoch << 32 | exchange number)
    int* xch_nvox[PTK_MAX_PEERS];                // [source rank] voxels in that rank's shard
    // dynamic state
    int n_range, n_ds, n_src, n_valid;
    int free_top, bump, n_vox, n_tomb, map_points;
    u32 ticket1, ticket2;
    u32 icp_arrive; u32 icp_release;
    int icp_done, err;
    int icp_searches, icp_it;
    Rigid icp_E, icp_T;
    SE3q icp_Tq;                         // T_icp between the launches of the sharded (multi-GPU) loop
};

// ------------------------------------------------------------------------------------
__device__ __forceinline__ u32 hash_key(u64 k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (u32)k;
}

// Hash of the LOCAL MAP table: separable in the three voxel coordinates (the classic spatial hash, upstream's
// multipliers, plus one xor-shift), so the 27 neighbours of a query cost two XORs each instead of a 64-bit mix.
__device__ __forceinline__ u32 hash_mix(u32 h) { return h ^ (h >> 15); }
constexpr u32 HASH_A = 73856093u, HASH_B = 19349663u, HASH_C = 83492791u;
__device__ __forceinline__ u32 hash_map_key(u64 k) {
    const u32 x = (u32)(k >> 42) & 0x1FFFFFu, y = (u32)(k >> 21) & 0x1FFFFFu, z = (u32)k & 0x1FFFFFu;   // biased coordinates
    return hash_mix((x * HASH_A) ^ (y * HASH_B) ^ (z * HASH_C));
}

// Owner rank of a voxel in the hash-sharded multi-GPU mode (upper hash bits; the table slot uses the lower).
__device__ __forceinline__ u32 shard_owner(u64 k, int n) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (u32)(k >> 32) % (u32)n;
}

__device__ __forceinline__ bool key_in_range(int kx, int ky, int kz) {
    return (kx > -KEY_BIAS + 1 && kx < KEY_BIAS - 1 && ky > -KEY_BIAS + 1 && ky < KEY_BIAS - 1 &&
            kz > -KEY_BIAS + 1 && kz < KEY_BIAS - 1);
}

__device__ __forceinline__ u64 pack_key(int kx, int ky, int kz) {
    return ((u64)(u32)(kx + KEY_BIAS) << 42) | ((u64)(u32)(ky + KEY_BIAS) << 21) | (u64)(u32)(kz + KEY_BIAS);
}

// kiss-icp voxel key: (p / size).cast<int>() - truncation toward zero of the IEEE quotient.
// A double division is ~30 instructions, so the quotient is first estimated as x * (1/size)
// (within 2 ulp of x / size); unless that estimate sits within 1e-11 (relative) of an integer -
// where the rounding of the true quotient could decide the result - its truncation IS the
// truncation of x / size.  Only the rare borderline coordinate pays for the exact division.
__device__ __forceinline__ int trunc_div(double x, double size, double inv) {
    const double q = x * inv;
    const double t = trunc(q);
    const double f = fabs(q - t);
    const double eps = fabs(q) * 1e-11 + 1e-290;
    if (f > eps && f < 1.0 - eps && fabs(q) < 2.0e9) return (int)t;
    return (int)(x / size);
}

__device__ __forceinline__ void voxel_key(double x, double y, double z, double size, double inv, int& kx, int& ky, int& kz) {
    kx = trunc_div(x, size, inv);
    ky = trunc_div(y, size, inv);
    kz = trunc_div(z, size, inv);
}

// First-seen table: find-or-insert `key`, keep the minimum `val`; returns the slot.
// The table has room for every point of the largest scan (`mask` + 1 slots), but a scan only fills a few per cent
// of it, and with dozens of lanes those sparse tables would not stay in L2.  So the first TABLE_NEAR probes of a key
// stay inside a small region at the front (`near_mask` + 1 slots, sized by the host from the previous scan's
// counts); only a key that finds all of them taken moves on to the rest of the table (the slots with the bit
// `near_mask + 1` set).  The probe sequence of a key is fixed, so every thread with that key ends in the same slot
// whatever the region sizes; correctness never depends on the hint, only the cache footprint does.
constexpr u32 TABLE_NEAR = 32;
// `first`: the key word already read from the key's first slot (callers with several inserts to do issue those
// reads together, ahead of the dependent atomics).  `lowered`: the value became the slot's minimum at the time of the
// call - a value that did not can never be the final minimum.
__device__ __forceinline__ u32 table_insert_min_from(u64* keys, u32* vals, u32 mask, u32 near_mask, u64 key, u32 val,
                                                     u64 first, bool& lowered) {
    const u32 h = hash_key(key);
    u32 slot = h & near_mask;
    u32 j = 0;
    u64 k = first;
    while (true) {
        if (k == key) break;
        if (k == KEY_EMPTY) {
            u64 prev = atomicCAS(keys + slot, KEY_EMPTY, key);
            if (prev == KEY_EMPTY || prev == key) break;
        }
        ++j;
        if (near_mask == mask || j < TABLE_NEAR) slot = (h + j) & near_mask;
        else slot = ((h + j) & mask) | (near_mask + 1u);
        k = *((volatile u64*)(keys + slot));
    }
    lowered = atomicMin(vals + slot, val) > val;
    return slot;
}
__device__ __forceinline__ u64 table_first_probe(const u64* keys, u32 near_mask, u64 key) {
    return *((volatile const u64*)(keys + (hash_key(key) & near_mask)));
}
__device__ __forceinline__ u32 table_insert_min(u64* keys, u32* vals, u32 mask, u32 near_mask, u64 key, u32 val, bool& lowered) {
    return table_insert_min_from(keys, vals, mask, near_mask, key, val, table_first_probe(keys, near_mask, key), lowered);
}

// Load input point i and apply the per-point deskew (kiss-icp DeSkewScan).  Returns false for a
// pixel without return (range-image mode only; kiss.py:59 `scan.field(RANGE) != 0`).
// Range-image mode restates kiss.py:60 (XYZLut: range * direction [+ offset]) and looks the deskew
// motion of the pixel's column up in the table k_col_motion built with the very same se3_exp call,
// so the result is bit-identical to the per-point path on the projected cloud.
__device__ __forceinline__ bool load_point(const StepParams& P, int i, double& x, double& y, double& z) {
    if (P.range) {
        const u32 r = __ldg(P.range + i);
        if (r == 0) return false;
        const double rr = (double)r * P.range_unit;
        // direction / offset LUTs are stored as three planes of n pixels each: consecutive pixels, consecutive words
        const size_t np = (size_t)P.n;
        x = __ldg(P.lut_dir + i) * rr; y = __ldg(P.lut_dir + np + i) * rr; z = __ldg(P.lut_dir + 2 * np + i) * rr;
        if (P.lut_off) {
            x = x + __ldg(P.lut_off + i); y = y + __ldg(P.lut_off + np + i); z = z + __ldg(P.lut_off + 2 * np + i);
        }
        if (P.flags & F_DESKEW) {
            const double* m = P.col_motion + (i % P.W);
            const int W = P.W;
            const double xo = ((m[0] * x + m[W] * y) + m[2 * W] * z) + m[9 * W];
            const double yo = ((m[3 * W] * x + m[4 * W] * y) + m[5 * W] * z) + m[10 * W];
            const double zo = ((m[6 * W] * x + m[7 * W] * y) + m[8 * W] * z) + m[11 * W];
            x = xo; y = yo; z = zo;
        }
        return true;
    }
    const double* p = P.xyz + 3 * (size_t)i;
    x = p[0]; y = p[1]; z = p[2];
    if (P.flags & F_DESKEW) {
        double s = P.ts[i] - 0.5;
        double a[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) a[k] = s * P.delta[k];
        Rigid M = se3_exp(a);
        double xo, yo, zo;
        rigid_apply(M, x, y, z, xo, yo, zo);
        x = xo; y = yo; z = zo;
    }
    return true;
}

// Deskew motion exp((t_w - 0.5) * delta) of every column of the range image, [12][W] (9 rotation
// entries row-major, then the translation), so that threads of consecutive columns read
// consecutive words.
__global__ void k_col_motion(const StepParams* params) {
    const StepParams& P = params[blockIdx.y];
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (!P.range || !(P.flags & F_DESKEW) || w >= P.W) return;
    const double s = P.col_ts[w] - 0.5;
    double a[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = s * P.delta[k];
    const Rigid M = se3_exp(a);
#pragma unroll
    for (int k = 0; k < 9; ++k) P.col_motion[k * P.W + w] = M.r[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) P.col_motion[(9 + k) * P.W + w] = M.t[k];
}

// kiss-icp Preprocess: min_range < |p| < max_range on the rounded norm.  Away from the two
// thresholds (by 1e-12 relative, far more than the rounding of the square root) the comparison of
// the squared norm decides; only borderline points take the square root.
__device__ __forceinline__ bool range_pass(const StepParams& P, double x, double y, double z) {
    if (!(P.flags & F_RANGE)) return true;
    const double n2 = (x * x + y * y) + z * z;
    if (P.min_range >= 0.0 && P.max_range > 0.0) {
        const double mx2 = P.max_range * P.max_range, mn2 = P.min_range * P.min_range;
        if (n2 < mx2 * (1.0 - 1e-12) && n2 > mn2 * (1.0 + 1e-12)) return true;
        if (n2 > mx2 * (1.0 + 1e-12) || n2 < mn2 * (1.0 - 1e-12)) return false;
    }
    const double nrm = sqrt(n2);
    return nrm < P.max_range && nrm > P.min_range;
}

// ------------------------------------------------------------------------------------
// K1: deskew + range filter + voxel key (grid 1) + first-seen atomicMin.
#ifndef PTK_SI_TILES
#define PTK_SI_TILES 4
#endif
constexpr int SI_TILES = PTK_SI_TILES;      // 256-pixel tiles a block of k_scan_insert walks in wide batches
#ifndef PTK_CT_TILES
#define PTK_CT_TILES 1
#endif
constexpr int CT_TILES = PTK_CT_TILES;      // 1024-item tiles a block of k_compact1 / k_compact2 walks (1: walking
                                            // several lengthens the look-back chain and measured slower)

__global__ void __launch_bounds__(256, 4) k_scan_insert(LaneDev* lanes, const StepParams* params, int tiles) {
    LaneDev& L = lanes[blockIdx.y];
    const StepParams& P = params[blockIdx.y];
    int n_pass = 0, n_valid = 0;
#pragma unroll 1
    for (int t = 0; t < tiles; ++t) {
        const int i = (blockIdx.x * tiles + t) * blockDim.x + threadIdx.x;
        if (i - (int)threadIdx.x >= P.n) break;          // block-uniform
        bool pass = false, valid = false, ins = false;
        u64 key = KEY_EMPTY;
        if (i < P.n) {
            double x, y, z;
            valid = load_point(P, i, x, y, z);
            pass = valid && range_pass(P, x, y, z);
            if (pass) {
                int kx, ky, kz;
                voxel_key(x, y, z, P.ds1_size, P.ds1_inv, kx, ky, kz);
                if (key_in_range(kx, ky, kz)) { key = pack_key(kx, ky, kz); ins = true; }
                else atomicOr(&L.err, ERR_KEYRANGE);
            }
        }
        // neighbouring pixels of a beam mostly fall into the same voxel: one table operation per distinct
        // key of the warp, issued by the lowest lane (= lowest point index) of each group
        // (slot1 keeps the slot only for a point that may still win its voxel: the lowest index of its warp's
        // group, and only if it lowered the slot's minimum - every final winner is both)
        const u32 am = __ballot_sync(0xffffffffu, ins);
        u32 slot = NONE;
        if (ins) {
            const u32 peers = __match_any_sync(am, key);
            const int leader = __ffs(peers) - 1;
            if ((int)(threadIdx.x & 31) == leader) {
                bool lowered;
                slot = table_insert_min(L.t1_keys, L.t1_vals, L.t_mask, P.near1 ? P.near1 : L.t_mask, key, (u32)i, lowered);
                if (!lowered) slot = NONE;
            }
        }
        if (i < P.n) L.slot1[i] = slot;
        n_pass += __popc(__ballot_sync(0xffffffffu, pass));
        n_valid += __popc(__ballot_sync(0xffffffffu, valid));
    }
    // counters: one atomic per warp and kernel (len(frame) of kiss.py:60 = pixels with a return)
    if ((threadIdx.x & 31) == 0) {
        if (n_pass) atomicAdd(&L.n_range, n_pass);
        if (P.range && n_valid) atomicAdd(&L.n_valid, n_valid);
    }
}

// K1, range-image form.  A thread owns ONE COLUMN of the image for SI_ROWS consecutive rows: the column's deskew
// motion (12 doubles) is loaded once instead of once per pixel, and everything a pixel needs - its range, its three
// LUT words - is addressed by the pixel index alone, so all the loads of the thread's pixels are issued before the
// first use (one memory round trip instead of three per pixel), and so are the first table probes of its inserts.
// Lanes of a warp are consecutive columns of a row, i.e. ascending pixel indices: the warp-level grouping of equal
// keys works as in k_scan_insert.  Arithmetic is load_point's, operation for operation.
#ifndef PTK_SI_ROWS
#define PTK_SI_ROWS 2          // measured: 2 rows at 64 registers beat 4 rows at 80 and 8 at 128
#endif
constexpr int SI_ROWS = PTK_SI_ROWS;
#ifndef PTK_SI_MINBLOCKS
#define PTK_SI_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(256, PTK_SI_MINBLOCKS) k_scan_insert_range(LaneDev* lanes, const StepParams* params, int H, int ncb) {
    LaneDev& L = lanes[blockIdx.y];
    const StepParams& P = params[blockIdx.y];
    const int W = P.W;
    const int cb = blockIdx.x % ncb, rb = blockIdx.x / ncb;
    const int w = cb * 256 + threadIdx.x;
    const int h0 = rb * SI_ROWS;
    const bool colok = w < W;
    const int wc = colok ? w : W - 1;
    const size_t np = (size_t)P.n;
    const bool deskew = P.flags & F_DESKEW;
    u32 r[SI_ROWS];
    double x[SI_ROWS], y[SI_ROWS], z[SI_ROWS];
#pragma unroll
    for (int k = 0; k < SI_ROWS; ++k) {
        const int h = min(h0 + k, H - 1);
        const size_t i = (size_t)h * W + wc;
        r[k] = __ldg(P.range + i);
        x[k] = __ldg(P.lut_dir + i); y[k] = __ldg(P.lut_dir + np + i); z[k] = __ldg(P.lut_dir + 2 * np + i);
    }
    double m[12];
    if (deskew) {
#pragma unroll
        for (int q = 0; q < 12; ++q) m[q] = P.col_motion[(size_t)q * W + wc];
    }
    u64 key[SI_ROWS];
    bool lead[SI_ROWS];
    int n_pass = 0, n_valid = 0;
    const u32 near = P.near1 ? P.near1 : L.t_mask;
#pragma unroll
    for (int k = 0; k < SI_ROWS; ++k) {
        const bool inb = colok && h0 + k < H;
        const bool valid = inb && r[k] != 0;
        const double rr = (double)r[k] * P.range_unit;
        double px = x[k] * rr, py = y[k] * rr, pz = z[k] * rr;
        if (P.lut_off) {
            const size_t i = (size_t)min(h0 + k, H - 1) * W + wc;
            px = px + __ldg(P.lut_off + i); py = py + __ldg(P.lut_off + np + i); pz = pz + __ldg(P.lut_off + 2 * np + i);
        }
        if (deskew) {
            const double xo = ((m[0] * px + m[1] * py) + m[2] * pz) + m[9];
            const double yo = ((m[3] * px + m[4] * py) + m[5] * pz) + m[10];
            const double zo = ((m[6] * px + m[7] * py) + m[8] * pz) + m[11];
            px = xo; py = yo; pz = zo;
        }
        const bool pass = valid && range_pass(P, px, py, pz);
        bool ins = false;
        key[k] = KEY_EMPTY;
        if (pass) {
            int kx, ky, kz;
            voxel_key(px, py, pz, P.ds1_size, P.ds1_inv, kx, ky, kz);
            if (key_in_range(kx, ky, kz)) { key[k] = pack_key(kx, ky, kz); ins = true; }
            else atomicOr(&L.err, ERR_KEYRANGE);
        }
        const u32 am = __ballot_sync(0xffffffffu, ins);
        lead[k] = false;
        if (ins) {
            const u32 peers = __match_any_sync(am, key[k]);
            lead[k] = (int)(threadIdx.x & 31) == __ffs(peers) - 1;
        }
        n_pass += __popc(__ballot_sync(0xffffffffu, pass));
        n_valid += __popc(__ballot_sync(0xffffffffu, valid));
    }
    u64 first[SI_ROWS];
#pragma unroll
    for (int k = 0; k < SI_ROWS; ++k) first[k] = lead[k] ? table_first_probe(L.t1_keys, near, key[k]) : KEY_EMPTY;
    u32 slot[SI_ROWS];
    bool lowered[SI_ROWS];
#pragma unroll
    for (int k = 0; k < SI_ROWS; ++k) {      // the atomics of all rows go out before any of their results is looked at
        slot[k] = NONE;
        lowered[k] = false;
        if (lead[k]) slot[k] = table_insert_min_from(L.t1_keys, L.t1_vals, L.t_mask, near, key[k], (u32)((h0 + k) * W + w), first[k], lowered[k]);
    }
#pragma unroll
    for (int k = 0; k < SI_ROWS; ++k)
        if (colok && h0 + k < H) L.slot1[(h0 + k) * W + w] = lowered[k] ? slot[k] : NONE;
    if ((threadIdx.x & 31) == 0) {
        if (n_pass) atomicAdd(&L.n_range, n_pass);
        if (n_valid) atomicAdd(&L.n_valid, n_valid);
    }
}

// Block-wide exclusive scan of per-thread counts (256 threads); returns offset, total in `total`.
__device__ __forceinline__ int block_excl_scan(int v, int& total) {
    __shared__ int wsum[8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int s = wsum[k];
        if (k < w) base += s;
        tot += s;
    }
    total = tot;
    __syncthreads();
    return base + inc - v;
}

// Ticketed tile id + decoupled look-back over tile aggregates: returns the number of selected
// items in all earlier tiles.  agg word = (epoch << 32) | count.
__device__ __forceinline__ u32 take_ticket(u32* ticket, u32 tbase) {
    __shared__ u32 s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u) - tbase;
    __syncthreads();
    u32 t = s_tile;
    __syncthreads();
    return t;
}

__device__ __forceinline__ int lookback_prefix(u64* agg, u32 tile, int total, u32 epoch) {
    __shared__ int s_prefix;
    if (threadIdx.x == 0) {
        __threadfence();
        *((volatile u64*)(agg + tile)) = ((u64)epoch << 32) | (u32)total;
    }
    if (threadIdx.x < 32) {
        int sum = 0;
        for (u32 j = threadIdx.x; j < tile; j += 32) {
            u64 a;
            do { a = *((volatile u64*)(agg + j)); } while ((u32)(a >> 32) != epoch);
            sum += (int)(u32)a;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (threadIdx.x == 0) s_prefix = sum;
    }
    __syncthreads();
    int p = s_prefix;
    __syncthreads();
    return p;
}

// K2: stable compaction of the grid-1 winners (ascending input index) into frame_downsample,
// and (F_SECOND) first-seen insert of every kept point into the grid-2 table.
// With F_SELECT_RANGE the selection is "passes the range filter" instead (Preprocess).
__global__ void __launch_bounds__(256, 4) k_compact1(LaneDev* lanes, const StepParams* params) {
    LaneDev& L = lanes[blockIdx.y];
    const StepParams& P = params[blockIdx.y];
    __shared__ u32 s_win[TILE];             // the tile's winners (input indices), ascending
    // a block walks CT_TILES tiles (fewer, longer-lived blocks: block dispatch is not free); every block
    // takes exactly CT_TILES tickets whether or not they are in range, so the host knows the next base
#pragma unroll 1
  for (int rep = 0; rep < CT_TILES; ++rep) {
    u32 tile = take_ticket(&L.ticket1, P.tbase1);
    u32 ntiles = (u32)((P.n + TILE - 1) / TILE);
    if (ntiles == 0) {
        if (tile == 0 && threadIdx.x == 0) L.n_ds = 0;
        continue;
    }
    if (tile >= ntiles) continue;
    const int i0 = (int)tile * TILE + threadIdx.x * 4;
    bool win[4];
    int cnt = 0;
    if (P.flags & F_SELECT_RANGE) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k;
            win[k] = false;
            if (i < P.n) {
                double x, y, z;
                win[k] = load_point(P, i, x, y, z) && range_pass(P, x, y, z);
            }
        }
    } else {
        // slot1 holds a slot only for the points that could still win their voxel (k_scan_insert): one 16 B load
        // for the thread's four points, then the gathers of the few candidates, all in flight together
        u32 sl[4];
        if (i0 + 3 < P.n) {
            const uint4 v = *reinterpret_cast<const uint4*>(L.slot1 + i0);
            sl[0] = v.x; sl[1] = v.y; sl[2] = v.z; sl[3] = v.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) sl[k] = i0 + k < P.n ? L.slot1[i0 + k] : NONE;
        }
        u32 tv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) tv[k] = sl[k] != NONE ? L.t1_vals[sl[k]] : NONE;
#pragma unroll
        for (int k = 0; k < 4; ++k) win[k] = sl[k] != NONE && tv[k] == (u32)(i0 + k);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt += win[k] ? 1 : 0;
    int total;
    const int local = block_excl_scan(cnt, total);
    {
        int q = local;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (win[k]) s_win[q++] = (u32)(i0 + k);
    }
    const int prefix = lookback_prefix(L.agg1, tile, total, P.epoch);      // (its barriers also publish s_win)
    // one winner per thread, whole warps at a time: recompute the point (cheaper than having stored all of them),
    // write frame_downsample, and insert into the grid-2 table - lanes of a warp hold ascending output positions
    for (int q0 = 0; q0 < total; q0 += blockDim.x) {
        const int q = q0 + threadIdx.x;
        const bool act = q < total;
        const int pos = prefix + q;
        bool ins = false;
        u64 key = KEY_EMPTY;
        if (act) {
            const int i = (int)s_win[q];
            double x, y, z;
            load_point(P, i, x, y, z);
            L.ds_x[pos] = x; L.ds_y[pos] = y; L.ds_z[pos] = z;
            L.ds_idx[pos] = (u32)i;
            if (P.flags & F_SECOND) {
                int kx, ky, kz;
                voxel_key(x, y, z, P.ds2_size, P.ds2_inv, kx, ky, kz);
                if (key_in_range(kx, ky, kz)) { key = pack_key(kx, ky, kz); ins = true; }
                else atomicOr(&L.err, ERR_KEYRANGE);
            }
        }
        if (P.flags & F_SECOND) {       // one table operation per distinct key of the warp (lowest lane = lowest pos)
            const u32 am = __ballot_sync(0xffffffffu, ins);
            u32 s2 = NONE;
            if (ins) {
                const u32 peers = __match_any_sync(am, key);
                if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
                    bool lowered;
                    s2 = table_insert_min(L.t2_keys, L.t2_vals, L.t_mask, P.near2 ? P.near2 : L.t_mask, key, (u32)pos, lowered);
                    if (!lowered) s2 = NONE;
                }
            }
            if (act) L.ds_slot2[pos] = s2;
        }
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) L.n_ds = prefix + total;
    __syncthreads();                        // s_win is free again
  }
}

// K3: stable compaction of the grid-2 winners into `source` (sensor frame + guess-transformed),
// and self-cleaning of table 1.
__global__ void __launch_bounds__(256) k_compact2(LaneDev* lanes, const StepParams* params) {
    LaneDev& L = lanes[blockIdx.y];
    const StepParams& P = params[blockIdx.y];
#pragma unroll 1
  for (int rep = 0; rep < CT_TILES; ++rep) {
    u32 tile = take_ticket(&L.ticket2, P.tbase2);
    if (tile == 0 && threadIdx.x == 0) L.icp_arrive = 0;    // barrier counter of the k_icp that follows
    int n_ds = L.n_ds;
    u32 ntiles = (u32)((n_ds + TILE - 1) / TILE);
    if (ntiles == 0) {
        if (tile == 0 && threadIdx.x == 0) L.n_src = 0;
        continue;
    }
    if (tile >= ntiles) continue;
    int j0 = (int)tile * TILE + threadIdx.x * 4;
    bool win[4];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int j = j0 + k;
        win[k] = false;
        if (j < n_ds) {
            u32 s2 = L.ds_slot2[j];
            win[k] = (s2 != NONE) && (L.t2_vals[s2] == (u32)j);
            // table 1 is no longer needed: clear the slot this point won
            u32 s1 = L.slot1[L.ds_idx[j]];
            L.t1_keys[s1] = KEY_EMPTY;
            L.t1_vals[s1] = NONE;
        }
        cnt += win[k] ? 1 : 0;
    }
    int total;
    int local = block_excl_scan(cnt, total);
    int prefix = lookback_prefix(L.agg2, tile, total, P.epoch);
    int pos = prefix + local;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!win[k]) continue;
        int j = j0 + k;
        double x = L.ds_x[j], y = L.ds_y[j], z = L.ds_z[j];
        L.s0_x[pos] = x; L.s0_y[pos] = y; L.s0_z[pos] = z;
        double xo, yo, zo;
        rigid_apply(P.guess, x, y, z, xo, yo, zo);
        L.s_x[pos] = xo; L.s_y[pos] = yo; L.s_z[pos] = zo;
        L.s_idx[pos] = (u32)j;
        ++pos;
    }
    if (tile == ntiles - 1 && threadIdx.x == 0) L.n_src = prefix + total;
  }
}

// Clear the scan tables after a stand-alone downsample (the step cleans them on the way).
__global__ void k_clean_tables(LaneDev* lanes, int which) {
    LaneDev& L = lanes[blockIdx.y];
    int n_ds = L.n_ds;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_ds; j += gridDim.x * blockDim.x) {
        if (which & 1) {
            u32 s1 = L.slot1[L.ds_idx[j]];
            L.t1_keys[s1] = KEY_EMPTY; L.t1_vals[s1] = NONE;
        }
        if (which & 2) {
            u32 s2 = L.ds_slot2[j];
            if (s2 != NONE) { L.t2_keys[s2] = KEY_EMPTY; L.t2_vals[s2] = NONE; }
        }
    }
}

// ------------------------------------------------------------------------------------
// Nearest map point of (sx,sy,sz) among the 27 neighbouring voxels, one warp per query
// (kiss-icp VoxelHashMap::GetCorrespondences).
//
// Lanes 0..26 probe one voxel each (one 16 B slot load, usually a first-probe hit) and compute a
// conservative lower bound of the distance from the query to that voxel's box.  Voxels are scanned
// four per round (lanes = slots, the 12 row loads of a round are independent): first the query's
// own voxel together with the three nearest boxes, then whatever still qualifies, skipping every
// voxel whose box is farther than the best distance so far (or than `max_d2`, beyond which a match
// would be rejected anyway).  A skipped
// voxel holds only points STRICTLY farther than the current best, so the result - including
// ties, which go to the first candidate in upstream's (i,j,l)-then-stored order through the
// lexicographic (d2, order id) compare - is the same as scanning all 27.  Unused slots of a
// block hold +inf (set when the voxel is created), so no per-voxel count is needed.
// What a search needs of the lane, read once (the kernel keeps it in shared memory): going through
// `LaneDev&` would reload every field from global memory at each use, on the latency path of a search.
struct MapView {
    const MapSlot* slots;
    const VoxelBlock* blocks;
    u32 mask;
    double voxel, voxel_inv;
};
__device__ __forceinline__ MapView map_view(const LaneDev& L) {
    MapView m;
    m.slots = L.m_slots; m.blocks = L.blocks; m.mask = L.m_mask; m.voxel = L.voxel_size; m.voxel_inv = L.voxel_inv;
    return m;
}

__device__ __forceinline__ void nn_visit(const VoxelBlock* B, int v, int lane, int sl, double sx, double sy, double sz,
                                         double& best, double& sec, double& thr, int& ord, int& sord,
                                         double& bx, double& by, double& bz) {
    double x = __ldg(&B->x[sl]), y = __ldg(&B->y[sl]), z = __ldg(&B->z[sl]);
    double dx = x - sx, dy = y - sy, dz = z - sz;
    double d2 = (dx * dx + dy * dy) + dz * dz;
    int o = v * MAXP + lane;
    // this lane's nearest (with the tie rule), its runner-up, and a lower bound of everything after those
    if (d2 < best || (d2 == best && o < ord)) { thr = sec; sec = best; sord = ord; best = d2; ord = o; bx = x; by = y; bz = z; }
    else if (d2 < sec) { thr = sec; sec = d2; sord = o; }
    else thr = fmin(thr, d2);
}

__device__ __forceinline__ double warp_min_upper(double best) {
    // an upper bound of the warp-wide minimum of non-negative doubles with one redux
    u32 hi = (u32)((u64)__double_as_longlong(best) >> 32);
    u32 mhi = __reduce_min_sync(0xffffffffu, hi);
    return __longlong_as_double((long long)(((u64)mhi << 32) | 0xffffffffull));
}

// Beside the winner the search reports ICP_KX runner-ups (t2[3*j..], ord2[j]; ord2[j] < 0 if there is
// none) and `others`: a lower bound of the distance from the query to every candidate of the 27 voxels
// EXCEPT the ones reported (the nearest visited point not handed out, the box distance of every voxel
// the search skipped), rounded down; negative if nothing was found.  k_icp uses them to prove, for a query that has moved but stayed
// in its voxel, which map point a new search would return.
__device__ __forceinline__ bool warp_nearest(const MapView& L, double sx, double sy, double sz, int lane, double max_d2,
                                             double& bd2, int& bord, double& tx, double& ty, double& tz, double& others,
                                             u64* qkey = nullptr, double* t2 = nullptr, int* ord2 = nullptr) {
    const u32 FULL = 0xffffffffu;
    const double v = L.voxel;
    int kx, ky, kz;
    voxel_key(sx, sy, sz, v, L.voxel_inv, kx, ky, kz);
    if (qkey) *qkey = key_in_range(kx, ky, kz) ? pack_key(kx, ky, kz) : KEY_EMPTY;
    u32 id = NONE;
    double lb2 = INFINITY;
    if (lane < 27 && key_in_range(kx, ky, kz)) {
        int di = lane / 9 - 1, dj = (lane / 3) % 3 - 1, dk = lane % 3 - 1;
        int nx = kx + di, ny = ky + dj, nz = kz + dk;
        u64 key = pack_key(nx, ny, nz);
        u32 slot = hash_map_key(key) & L.mask;
        while (true) {
            const ulonglong2 raw = __ldg(reinterpret_cast<const ulonglong2*>(L.slots + slot));
            if (raw.x == key) { id = (u32)raw.y; break; }
            if (raw.x == KEY_EMPTY) break;
            slot = (slot + 1) & L.mask;
        }
        // voxel n covers [n v, (n+1) v) for n > 0, (-v, v) for n == 0 and ((n-1) v, n v] for n < 0
        // (keys truncate toward zero); 1e-7 m of slack per axis covers every rounding involved
        double lo, hi, ax, ay, az;
        lo = (double)(nx > 0 ? nx : nx - 1) * v; hi = (double)(nx < 0 ? nx : nx + 1) * v;
        ax = fmax(fmax(lo - sx, sx - hi) - 1e-7, 0.0);
        lo = (double)(ny > 0 ? ny : ny - 1) * v; hi = (double)(ny < 0 ? ny : ny + 1) * v;
        ay = fmax(fmax(lo - sy, sy - hi) - 1e-7, 0.0);
        lo = (double)(nz > 0 ? nz : nz - 1) * v; hi = (double)(nz < 0 ? nz : nz + 1) * v;
        az = fmax(fmax(lo - sz, sz - hi) - 1e-7, 0.0);
        lb2 = ((ax * ax + ay * ay) + az * az) * (1.0 - 1e-9);
    }
    double best = INFINITY, sec = INFINITY, thr = INFINITY, bx = 0, by = 0, bz = 0;
    int ord = 0x7fffffff, sord = 0x7fffffff;
    const int sl = lane < MAXP ? lane : 0;
    double bound = max_d2;
    u32 remaining = __ballot_sync(FULL, id != NONE);
    bool first = true;
    while (true) {
        u32 mask = __ballot_sync(FULL, lb2 <= bound) & remaining;
        if (!mask) break;
        int v0, v1, v2, v3;
        if (first) {
            // first round: the query's own voxel (box distance 0) and the three other voxels whose boxes are
            // nearest - they are the likeliest to be needed at all, and fetching them together with the own
            // voxel saves a dependent round trip; later rounds take what still qualifies, in index order
            first = false;
            const u32 kbits = ((u32)((u64)__double_as_longlong(lb2) >> 32) & ~31u) | (u32)lane;
            u32 m = mask;
            u32 k0 = __reduce_min_sync(FULL, (m >> lane) & 1u ? kbits : 0xffffffffu);
            v0 = (int)(k0 & 31u); m &= ~(1u << v0);
            v1 = v2 = v3 = v0;
            if (m) { u32 k1 = __reduce_min_sync(FULL, (m >> lane) & 1u ? kbits : 0xffffffffu); v1 = (int)(k1 & 31u); m &= ~(1u << v1); }
            if (m) { u32 k2 = __reduce_min_sync(FULL, (m >> lane) & 1u ? kbits : 0xffffffffu); v2 = (int)(k2 & 31u); m &= ~(1u << v2); }
            if (m) { u32 k3 = __reduce_min_sync(FULL, (m >> lane) & 1u ? kbits : 0xffffffffu); v3 = (int)(k3 & 31u); }
        } else {
            v0 = __ffs(mask) - 1; mask &= mask - 1;
            v1 = v0; v2 = v0; v3 = v0;
            if (mask) { v1 = __ffs(mask) - 1; mask &= mask - 1; }
            if (mask) { v2 = __ffs(mask) - 1; mask &= mask - 1; }
            if (mask) { v3 = __ffs(mask) - 1; }
        }
        remaining &= ~((1u << v0) | (1u << v1) | (1u << v2) | (1u << v3));
        const VoxelBlock* B0 = L.blocks + __shfl_sync(FULL, id, v0);
        const VoxelBlock* B1 = L.blocks + __shfl_sync(FULL, id, v1);
        const VoxelBlock* B2 = L.blocks + __shfl_sync(FULL, id, v2);
        const VoxelBlock* B3 = L.blocks + __shfl_sync(FULL, id, v3);
        if (lane < MAXP) {
            nn_visit(B0, v0, lane, sl, sx, sy, sz, best, sec, thr, ord, sord, bx, by, bz);
            if (v1 != v0) nn_visit(B1, v1, lane, sl, sx, sy, sz, best, sec, thr, ord, sord, bx, by, bz);
            if (v2 != v0) nn_visit(B2, v2, lane, sl, sx, sy, sz, best, sec, thr, ord, sord, bx, by, bz);
            if (v3 != v0) nn_visit(B3, v3, lane, sl, sx, sy, sz, best, sec, thr, ord, sord, bx, by, bz);
        }
        bound = fmin(bound, warp_min_upper(best));
    }
    // lexicographic (d2, ord) argmin: d2 >= 0, so its bit pattern orders like the value
    const u64 bits = (u64)__double_as_longlong(best);
    const u32 hi = (u32)(bits >> 32), lo = (u32)bits;
    const u32 mhi = __reduce_min_sync(FULL, hi);
    const u32 mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
    const bool cand = (hi == mhi) && (lo == mlo);
    const u32 mord = __reduce_min_sync(FULL, cand ? (u32)ord : 0x7fffffffu);
    const bool found = mord != 0x7fffffffu;
    const int owner = found ? (int)(mord % MAXP) : 0;
    tx = __shfl_sync(FULL, bx, owner);
    ty = __shfl_sync(FULL, by, owner);
    tz = __shfl_sync(FULL, bz, owner);
    bd2 = __longlong_as_double((long long)(((u64)mhi << 32) | (u64)mlo));
    bord = (int)mord;
    // runner-ups: ICP_KX more rounds of "nearest of what is left".  A lane knows its two nearest candidates
    // exactly and a lower bound (thr) of the rest; `cons` counts how many of them have been handed out.
    int cons = (found && lane == owner) ? 1 : 0;
    if (t2 != nullptr) {
#pragma unroll
        for (int j = 0; j < ICP_KX; ++j) {
            const double r = cons == 0 ? best : (cons == 1 ? sec : INFINITY);
            const int rord = cons == 0 ? ord : sord;
            const u64 rbits = (u64)__double_as_longlong(r);
            const u32 rhi = (u32)(rbits >> 32), rlo = (u32)rbits;
            const u32 mrhi = __reduce_min_sync(FULL, rhi);
            const u32 mrlo = __reduce_min_sync(FULL, rhi == mrhi ? rlo : 0xffffffffu);
            const bool has = found && mrhi < 0x7ff00000u;
            const int ownj = __ffs(__ballot_sync(FULL, rhi == mrhi && rlo == mrlo)) - 1;
            int oj = __shfl_sync(FULL, rord, ownj);
            if (!has) oj = -1;
            ord2[j] = oj;
            t2[3 * j] = t2[3 * j + 1] = t2[3 * j + 2] = 0.0;
            if (has) {      // its coordinates: one more (cache-hot) read of the voxel it sits in
                if (lane == ownj) ++cons;
                const VoxelBlock* Bj = L.blocks + __shfl_sync(FULL, id, oj / MAXP);
                const int sj = oj % MAXP;
                t2[3 * j] = __ldg(&Bj->x[sj]); t2[3 * j + 1] = __ldg(&Bj->y[sj]); t2[3 * j + 2] = __ldg(&Bj->z[sj]);
            }
        }
    }
    // lower bound of the squared distance to every candidate that was not handed out
    double other = cons == 0 ? best : (cons == 1 ? sec : thr);
    if ((remaining >> lane) & 1u) other = fmin(other, lb2);
    const u32 ohi = __reduce_min_sync(FULL, (u32)((u64)__double_as_longlong(other) >> 32));
    const double d2_other = __longlong_as_double((long long)((u64)ohi << 32));     // low word zero: rounds down
    others = found ? sqrt(d2_other) * (1.0 - 1e-12) : -1.0;
    return found;
}

// ---- TMA pieces (1-D bulk copy global -> shared, completion counted on an mbarrier) -----------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* mbar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* mbar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, u32 bytes, unsigned long long* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(mbar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, u32 phase) {
    u32 ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_addr(mbar)), "r"(phase) : "memory");
    } while (!ok);
}

// ---- the same search with ONE THREAD per query (the ICP kernel's form) -------------------------------------
// Every lane of a warp carries its own query (or none: active = false), so all missed points of a 32-point group
// are searched at the same time and the latency of a group's searches is that of one search.  Result semantics
// are those of warp_nearest: winner = lexicographic minimum of (d2, order id) over the 27 voxels (B.6), ICP_KX
// runner-ups, and `others` = a lower bound of the distance to every candidate not handed out.
//   1. 27 table probes, issued nine at a time; voxel ids and a rounded-down float of the box distance go to this
//      warp's shared-memory scratch (ids/lbs [27][32], one column per lane);
//   2. the query's own voxel, then each lane walks ITS OWN list of voxels whose box is not farther than the best
//      so far (a skipped voxel holds only strictly farther points).  The warp moves in rounds: every lane that
//      still has a voxel to look at fetches its 480 B of point rows with ONE TMA bulk copy (cp.async.bulk) into
//      its own slice of the warp's staging buffer, the copies of a round complete on the warp's mbarrier, and the
//      distances are computed from shared memory.  (Reading the rows with per-thread 16 B loads instead made the
//      kernel L1-bound: 32 lanes x 30 loads, every one a different cache line, per round.);
//   3. coordinates of the winner and the runner-ups are re-read through their order ids (cache-hot).
struct NearestOut {
    double tx, ty, tz, others;
    double t2[3 * ICP_KX];
    u64 qkey;
    int ord;
    int ord2[ICP_KX];
};

constexpr int STAGE_STRIDE = 62;            // doubles per lane in the staging buffer: 496 B (an odd number of 16 B units:
                                            // the lanes' LDS.128 fall into different banks), 480 B of them used
__device__ __forceinline__ void thread_nearest(const MapView& M, u32* ids, float* lbs, double* stage, unsigned long long* mbar,
                                               u32& phase, int lane, bool active, double sx, double sy, double sz, double max_d2,
                                               NearestOut& R) {
    int kx, ky, kz;
    voxel_key(sx, sy, sz, M.voxel, M.voxel_inv, kx, ky, kz);
    const bool inr = active && key_in_range(kx, ky, kz);
    R.qkey = inr ? pack_key(kx, ky, kz) : KEY_EMPTY;
    const double v = M.voxel;
    u32 present = 0;
    if (inr) {
        const u64 base = pack_key(kx - 1, ky - 1, kz - 1);
        const u32 bx = (u32)(kx - 1 + KEY_BIAS), by = (u32)(ky - 1 + KEY_BIAS), bz = (u32)(kz - 1 + KEY_BIAS);
        u32 hy[3], hz[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { hy[j] = (by + (u32)j) * HASH_B; hz[j] = (bz + (u32)j) * HASH_C; }
#pragma unroll 1
        for (int i = 0; i < 3; ++i) {
            const u32 hxi = (bx + (u32)i) * HASH_A;
            const u64 keyi = base + ((u64)i << 42);
            ulonglong2 raw[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const u32 slot = hash_mix(hxi ^ hy[q / 3] ^ hz[q % 3]) & M.mask;
                raw[q] = __ldg(reinterpret_cast<const ulonglong2*>(M.slots + slot));
            }
            // voxel n covers [n v, (n+1) v) for n > 0, (-v, v) for n == 0 and ((n-1) v, n v] for n < 0
            // (keys truncate toward zero); 1e-7 m of slack per axis covers every rounding involved
            const int nx = kx - 1 + i;
            double lo = (double)(nx > 0 ? nx : nx - 1) * v, hi = (double)(nx < 0 ? nx : nx + 1) * v;
            const double ax = fmax(fmax(lo - sx, sx - hi) - 1e-7, 0.0);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const u64 key = keyi + ((u64)(q / 3) << 21) + (u64)(q % 3);
                u32 id = NONE;
                if (raw[q].x == key) id = (u32)raw[q].y;
                else if (raw[q].x != KEY_EMPTY) {          // collision or tombstone: keep probing
                    u32 slot = (hash_mix(hxi ^ hy[q / 3] ^ hz[q % 3]) + 1u) & M.mask;
                    while (true) {
                        const ulonglong2 r2 = __ldg(reinterpret_cast<const ulonglong2*>(M.slots + slot));
                        if (r2.x == key) { id = (u32)r2.y; break; }
                        if (r2.x == KEY_EMPTY) break;
                        slot = (slot + 1) & M.mask;
                    }
                }
                const int c = i * 9 + q;
                ids[c * 32 + lane] = id;
                if (id != NONE) {
                    const int ny = ky - 1 + q / 3, nz = kz - 1 + q % 3;
                    lo = (double)(ny > 0 ? ny : ny - 1) * v; hi = (double)(ny < 0 ? ny : ny + 1) * v;
                    const double ay = fmax(fmax(lo - sy, sy - hi) - 1e-7, 0.0);
                    lo = (double)(nz > 0 ? nz : nz - 1) * v; hi = (double)(nz < 0 ? nz : nz + 1) * v;
                    const double az = fmax(fmax(lo - sz, sz - hi) - 1e-7, 0.0);
                    const double lb2 = ((ax * ax + ay * ay) + az * az) * (1.0 - 1e-9);
                    lbs[c * 32 + lane] = __double2float_rd(lb2);
                    present |= 1u << c;
                }
            }
        }
    }
    // top three candidates in the search's lexicographic (d2, order id) order + the nearest of all the others
    double d0 = INFINITY, d1 = INFINITY, d2 = INFINITY, rest = INFINITY;
    int o0 = -1, o1 = -1, o2 = -1;
    double bound = max_d2;
    u32 todo = present;
    int vv = (present >> 13) & 1u ? 13 : -1;     // the query's own voxel first
    todo &= ~(1u << 13);
    double* const mine = stage + lane * STAGE_STRIDE;
    while (true) {
        if (vv < 0) {
            while (todo) {
                const int c = __ffs(todo) - 1;
                todo &= todo - 1;
                const float lb = lbs[c * 32 + lane];
                if ((double)lb <= bound) { vv = c; break; }
                rest = fmin(rest, (double)lb);
            }
        }
        const u32 act = __ballot_sync(0xffffffffu, vv >= 0);
        if (!act) break;
        if (lane == 0) mbar_expect_tx(mbar, (u32)__popc(act) * (u32)(3 * MAXP * sizeof(double)));
        __syncwarp();
        if (vv >= 0) tma_load_1d(mine, M.blocks + ids[vv * 32 + lane], (u32)(3 * MAXP * sizeof(double)), mbar);
        mbar_wait(mbar, phase);
        phase ^= 1u;
        if (vv >= 0) {
            const double2* rows = reinterpret_cast<const double2*>(mine);
            const int obase = vv * MAXP;
#pragma unroll
            for (int h = 0; h < MAXP / 4; ++h) {
                const double2 xa = rows[2 * h], xb = rows[2 * h + 1];
                const double2 ya = rows[MAXP / 2 + 2 * h], yb = rows[MAXP / 2 + 2 * h + 1];
                const double2 za = rows[MAXP + 2 * h], zb = rows[MAXP + 2 * h + 1];
                const double xs[4] = {xa.x, xa.y, xb.x, xb.y}, ys[4] = {ya.x, ya.y, yb.x, yb.y}, zs[4] = {za.x, za.y, zb.x, zb.y};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double dx = xs[k] - sx, dy = ys[k] - sy, dz = zs[k] - sz;
                    const double d = (dx * dx + dy * dy) + dz * dz;
                    const int o = obase + 4 * h + k;
                    if (d < d2 || (d == d2 && o < o2)) {                 // beats the third (unused slots hold +inf: never)
                        rest = fmin(rest, d2);
                        if (d < d1 || (d == d1 && o < o1)) {
                            d2 = d1; o2 = o1;
                            if (d < d0 || (d == d0 && o < o0)) { d1 = d0; o1 = o0; d0 = d; o0 = o; }
                            else { d1 = d; o1 = o; }
                        } else { d2 = d; o2 = o; }
                    } else rest = fmin(rest, d);
                }
            }
            bound = fmin(bound, d0);
        }
        vv = -1;
        __syncwarp();             // everybody is done with the staging buffer before the next round's copies land
    }
    R.ord = o0;
    R.tx = R.ty = R.tz = 0.0;
    if (o0 >= 0) {
        const VoxelBlock* B = M.blocks + ids[(o0 / MAXP) * 32 + lane];
        const int sl = o0 % MAXP;
        R.tx = __ldg(&B->x[sl]); R.ty = __ldg(&B->y[sl]); R.tz = __ldg(&B->z[sl]);
    }
    const int oj[2] = {o1, o2};
    const double dj[2] = {d1, d2};
    double other = rest;
#pragma unroll
    for (int j = 0; j < ICP_KX; ++j) {
        const int o = j < 2 ? oj[j] : -1;
        R.ord2[j] = o;
        R.t2[3 * j] = R.t2[3 * j + 1] = R.t2[3 * j + 2] = 0.0;
        if (o >= 0) {
            const VoxelBlock* B = M.blocks + ids[(o / MAXP) * 32 + lane];
            const int sl = o % MAXP;
            R.t2[3 * j] = __ldg(&B->x[sl]); R.t2[3 * j + 1] = __ldg(&B->y[sl]); R.t2[3 * j + 2] = __ldg(&B->z[sl]);
        }
    }
#pragma unroll
    for (int j = ICP_KX; j < 2; ++j) other = fmin(other, dj[j]);      // runner-ups the cache does not keep count as "others"
    R.others = o0 >= 0 ? sqrt(other) * (1.0 - 1e-12) : -1.0;
}

// The distinct sums behind the 27 normal-equation terms: the 27 columns the oracle
// reduces contain 6 structural zeros, three copies of w and three +-pairs, so 16 sums (+ the
// correspondence count) determine all of them bit for bit (a sum of negated terms is the negated
// sum).  Order: w, w*sx, w*sy, w*sz, the 6 entries of the lower-right block, the 6 of Jtr, count.
__device__ __forceinline__ void lin_terms(double sx, double sy, double sz, double tx, double ty, double tz,
                                          double kernel, double* c) {
    double rx = sx - tx, ry = sy - ty, rz = sz - tz;
    double r2 = (rx * rx + ry * ry) + rz * rz;
    double kk = kernel + r2;
    double w = (kernel * kernel) / (kk * kk);
    double wrx = w * rx, wry = w * ry, wrz = w * rz;
    c[0] = w; c[1] = w * sx; c[2] = w * sy; c[3] = w * sz;
    c[4] = w * (sy * sy + sz * sz); c[5] = -(w * (sx * sy)); c[6] = -(w * (sx * sz));
    c[7] = w * (sx * sx + sz * sz); c[8] = -(w * (sy * sz));
    c[9] = w * (sx * sx + sy * sy);
    c[10] = wrx; c[11] = wry; c[12] = wrz;
    c[13] = sy * wrz - sz * wry; c[14] = sz * wrx - sx * wrz; c[15] = sx * wry - sy * wrx;
}

__device__ __forceinline__ double warp_butterfly(double x) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) x = x + __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// 16 values per lane -> 16 warp-wide sums with 16 shuffles instead of 80: at the stage with xor
// mask m a lane keeps one half of its values and adds the partner's copies of that half, so the
// working set halves every stage.  Every sum is still x[l] + x[l ^ m] over the same partial sums
// as warp_butterfly (addition commutes bit for bit), i.e. the canonical adjacent-pairs tree.
// Returns the total of value bitrev4(lane & 15).
__device__ __forceinline__ double warp_reduce16(const double* c, int lane) {
    const u32 FULL = 0xffffffffu;
    double a[8], b4[4], b2[2], b1;
    const bool u0 = lane & 1, u1 = lane & 2, u2 = lane & 4, u3 = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double keep = u0 ? c[i + 8] : c[i], send = u0 ? c[i] : c[i + 8];
        a[i] = keep + __shfl_xor_sync(FULL, send, 1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double keep = u1 ? a[i + 4] : a[i], send = u1 ? a[i] : a[i + 4];
        b4[i] = keep + __shfl_xor_sync(FULL, send, 2);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double keep = u2 ? b4[i + 2] : b4[i], send = u2 ? b4[i] : b4[i + 2];
        b2[i] = keep + __shfl_xor_sync(FULL, send, 4);
    }
    {
        const double keep = u3 ? b2[1] : b2[0], send = u3 ? b2[0] : b2[1];
        b1 = keep + __shfl_xor_sync(FULL, send, 8);
    }
    return b1 + __shfl_xor_sync(FULL, b1, 16);
}

// Adjacent-pairs binary tree over p[0..n), zero padded to a power of two: the canonical reduction
// over 32-point group partials (oracle/canon.py pairwise_tree_sum), computed by one warp with
// butterflies only.  Element idx sits in lane idx & 31 of butterfly idx >> 5.
__device__ __forceinline__ double warp_tree_sum(const double* p, int n, int lane) {
    int m = 1;
    while (m * 32 < n) m <<= 1;
    if (m == 1) return warp_butterfly(lane < n ? __ldcg(p + lane) : 0.0);
    if (m <= 32) {
        double reg = 0.0;
        for (int j = 0; j < m; j += 4) {      // m is 2 or a multiple of 4; loads of a quad go out together
            int i0 = j * 32 + lane;
            double x0 = i0 < n ? __ldcg(p + i0) : 0.0;
            double x1 = i0 + 32 < n ? __ldcg(p + i0 + 32) : 0.0;
            double x2 = (m > 2 && i0 + 64 < n) ? __ldcg(p + i0 + 64) : 0.0;
            double x3 = (m > 2 && i0 + 96 < n) ? __ldcg(p + i0 + 96) : 0.0;
            x0 = warp_butterfly(x0); x1 = warp_butterfly(x1);
            if (lane == j) reg = x0;
            if (lane == j + 1) reg = x1;
            if (m > 2) {
                x2 = warp_butterfly(x2); x3 = warp_butterfly(x3);
                if (lane == j + 2) reg = x2;
                if (lane == j + 3) reg = x3;
            }
        }
        return warp_butterfly(reg);
    }
    double U[8];                        // m in {64, 128, 256}: n_src up to 262144
    const int mc = m >> 5;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        U[c] = 0.0;
        if (c < mc) {
            double reg = 0.0;
            for (int j = 0; j < 32; ++j) {
                int idx = (c * 32 + j) * 32 + lane;
                double x = warp_butterfly(idx < n ? __ldcg(p + idx) : 0.0);
                if (lane == j) reg = x;
            }
            U[c] = warp_butterfly(reg);
        }
    }
    return ((U[0] + U[1]) + (U[2] + U[3])) + ((U[4] + U[5]) + (U[6] + U[7]));
}

// Where the 17 sums go in the symmetric 6x6 matrix (row-major, both triangles): +k+1 = +red[k],
// -(k+1) = -red[k], 0 = structural zero.  J = [I | -hat(s)].
__constant__ signed char kJtJMap[36] = {
    1, 0, 0, 0, 4, -3,
    0, 1, 0, -4, 0, 2,
    0, 0, 1, 3, -2, 0,
    0, -4, 3, 5, 6, 7,
    4, 0, -2, 6, 8, 9,
    -3, 2, 0, 7, 9, 10};
// LDL^T with diagonal pivoting of the 6x6 normal equations, oracle/canon.py ldlt_solve6 operation for operation
// (same pivot choice, same divisions, same (a[j][k] * d) then multiply-subtract, same subtraction order in the two
// substitutions), laid out for latency: ONE ENTRY of the lower triangle per lane (lane t = i (i + 1) / 2 + j holds
// a[i][j], i >= j; the canonical algorithm never reads the upper triangle again once a column is eliminated).  A
// pivot step is four dependent shuffle stages - diagonal fetch, the symmetric permutation as ONE indexed shuffle,
// pivot broadcast, the two L factors of an entry.  (A column-per-lane form with divergent updates took 7.6 k cycles
// per solve; this one about 6.4 k, most of it the seven dependent double-precision divisions.)
// Writes x to dx_out[6] (shared memory) and returns whether the solve is finite; warp-uniform.
__device__ __forceinline__ int tri_idx(int i, int j) { return (i * (i + 1)) / 2 + j; }
__device__ __forceinline__ bool ldlt_solve6_tri(const double* red, int lane, double* dx_out) {
    const u32 FULL = 0xffffffffu;
    // (mi, mj) of this lane; lanes >= 21 shadow lane 20
    int mi = 0, mj = 0;
    {
        const int t = lane < 21 ? lane : 20;
#pragma unroll
        for (int i = 1; i < 6; ++i) if (t >= (i * (i + 1)) / 2) mi = i;
        mj = t - (mi * (mi + 1)) / 2;
    }
    double a;
    {
        const int m = kJtJMap[mi * 6 + mj];
        a = m > 0 ? red[m - 1] : (m < 0 ? -red[-m - 1] : 0.0);
    }
    int perm[6] = {0, 1, 2, 3, 4, 5};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double best = fabs(__shfl_sync(FULL, a, tri_idx(k, k)));
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            const double v = fabs(__shfl_sync(FULL, a, tri_idx(i, i)));
            if (v > best) { best = v; p = i; }
        }
        if (p != k) {                                   // warp-uniform
            int si = mi == k ? p : (mi == p ? k : mi);
            int sj = mj == k ? p : (mj == p ? k : mj);
            if (si < sj) { const int t = si; si = sj; sj = t; }
            a = __shfl_sync(FULL, a, tri_idx(si, sj));
            const int tk = perm[k];
#pragma unroll
            for (int r = k + 1; r < 6; ++r)
                if (r == p) { perm[k] = perm[r]; perm[r] = tk; }
        }
        const double d = __shfl_sync(FULL, a, tri_idx(k, k));
        if (d == 0.0 || d != d) return false;
        if (mj == k && mi > k) a = a / d;                            // L(i, k)
        if (k < 5) {
            const double li = __shfl_sync(FULL, a, tri_idx(mi, k));  // (read by the trailing entries only)
            const double lj = __shfl_sync(FULL, a, tri_idx(mj, k));
            if (mj > k) a = a - li * (lj * d);                       // a[i][j] -= a[i][k] * (a[j][k] * d), i >= j > k
        }
    }
    // rows live in lanes 0..5 from here on: y = P b, L z = y, w = z / D, L^T v = w
    const int r = lane < 6 ? lane : 5;
    int pr = perm[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) if (r == i) pr = perm[i];
    double y = -red[10 + pr];
    double lrow[5], lcol[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        lrow[q] = __shfl_sync(FULL, a, r > q ? tri_idx(r, q) : 0);              // L(r, q), q < r
        lcol[q] = __shfl_sync(FULL, a, q + 1 > r ? tri_idx(q + 1, r) : 0);      // L(q + 1, r), q + 1 > r
    }
    const double dr = __shfl_sync(FULL, a, tri_idx(r, r));
#pragma unroll
    for (int q = 0; q < 5; ++q) {                   // column sweep = the canonical subtraction order of every row
        const double yq = __shfl_sync(FULL, y, q);
        if (r > q) y = y - lrow[q] * yq;
    }
    y = y / dr;
    double yv[6];
#pragma unroll
    for (int i = 4; i >= 0; --i) {                  // rows 4..0; a row subtracts its terms in ascending column order
        yv[i + 1] = __shfl_sync(FULL, y, i + 1);    // final since the previous round
        if (r == i) {
#pragma unroll
            for (int q = i + 1; q < 6; ++q) y = y - lcol[q - 1] * yv[q];
        }
    }
    const bool fin = fabs(y) <= 1.7976931348623157e308;
    const bool ok = (__ballot_sync(FULL, fin || lane >= 6) == FULL);
    if (lane < 6) dx_out[pr] = y;
    __syncwarp();
    return ok;
}

struct SolveSmem {
    double dx[6];
};

// Warp 0 of every block: expand the 17 sums into the 6x6 system, solve it (ldlt_solve6_tri), update
// T_icp, decide termination (kiss-icp RegisterFrame loop body after BuildLinearSystem).
__device__ __noinline__ void icp_solve_step(LaneDev& L, const StepParams& P, StepOut& O, const double* red, SolveSmem* S,
                                            Rigid* sE, SE3q* sT, int* s_done, int it, bool writer, int lane, double eps, int max_iters) {
#ifdef PTK_SOLVE_CLOCKS
    __shared__ long long s_sc[6];
    long long tl_ = clock64();
#define SOLVE_TICK(k) do { if (lane == 0) { long long t_ = clock64(); s_sc[k] = (it == 0 ? 0 : s_sc[k]) + (t_ - tl_); tl_ = t_; } } while (0)
#else
#define SOLVE_TICK(k) do { } while (0)
#endif
    const int n_corr = (int)red[16];
    int done = 0, status = 0;
    double nrm = 0.0;
    bool ok = false;
    if (n_corr == 0) {
        status = 1; done = 1;            // B.5
    } else {
        SOLVE_TICK(0);
        ok = ldlt_solve6_tri(red, lane, S->dx);
        SOLVE_TICK(1);
        SOLVE_TICK(2);
        if (!ok) { status = 2; done = 1; }
    }
    if (lane == 0) {
        Rigid Enew = rigid_identity();
        if (ok) {
            const double* dx = S->dx;
            SE3q Eq;
            Enew = se3_exp_q(dx, &Eq.q);
            Eq.t[0] = Enew.t[0]; Eq.t[1] = Enew.t[1]; Eq.t[2] = Enew.t[2];
            SOLVE_TICK(3);
            *sT = se3q_mul(Eq, *sT);
            SOLVE_TICK(4);
            nrm = sqrt(((((dx[0] * dx[0] + dx[1] * dx[1]) + dx[2] * dx[2]) + dx[3] * dx[3]) + dx[4] * dx[4]) + dx[5] * dx[5]);
            if (nrm < eps) done = 1;
        }
        if (it + 1 >= max_iters) done = 1;
        *sE = Enew;
        *s_done = done;
        if (done && writer) {
            O.pose = se3q_matrix(se3q_mul(*sT, se3q_from_rigid(P.guess)));
            O.iterations = it + 1; O.n_corr = n_corr; O.status = status; O.dx_norm = nrm;
        }
        SOLVE_TICK(5);
#ifdef PTK_SOLVE_CLOCKS
        if (done && writer) {     // this build reports the solve's own phases instead of the iteration's
#pragma unroll
            for (int k = 0; k < 6; ++k) O.icp_cyc[k] = s_sc[k];
        }
#endif
    }
#undef SOLVE_TICK
}

constexpr int SHARD_NO_ORD_K = 1 << 30;     // "no candidate" in an exchanged record

// K4a: the searches of ICP iteration 0.  The first iteration has to search for EVERY source point (there is no
// cache entry yet), which made it a third of the loop's time when the points were searched one per warp inside the
// cooperative kernel.  Here it is an ordinary wide launch: one THREAD per source point (thread_nearest), a warp per
// 32-point group, every lane of the batch at once.  The cache entries go to the lane's global cache arrays; k_icp
// starts from them.
constexpr int S0_THREADS = 64;
__global__ void __launch_bounds__(S0_THREADS) k_icp_search0(LaneDev* lanes, const StepParams* params) {
    LaneDev& L = lanes[blockIdx.y];
    const StepParams& P = params[blockIdx.y];
    __shared__ __align__(16) double s_stage[S0_THREADS / 32][32 * STAGE_STRIDE];      // TMA destination, one slice per lane
    __shared__ u32 s_ids[S0_THREADS / 32][27 * 32];
    __shared__ float s_lbs[S0_THREADS / 32][27 * 32];
    __shared__ unsigned long long s_mbar[S0_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        mbar_init(&s_mbar[warp], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    u32 phase = 0;
    const int n_src = L.n_src;
    if ((L.n_vox == 0 && L.shard_n <= 1) || n_src == 0) return;     // (a rank of a sharded map may own no voxel yet)
    const MapView M = map_view(L);
    const double max_corr = P.max_corr;
    const double max_d2 = (max_corr * max_corr) * (1.0 + 1e-9);   // farther candidates are rejected anyway
    const size_t cap = (size_t)L.cap_points;
    for (int g = blockIdx.x * (S0_THREADS / 32) + warp; g * 32 < n_src; g += gridDim.x * (S0_THREADS / 32)) {
        const int p = g * 32 + lane;
        const bool live = p < n_src;
        double sx = 0, sy = 0, sz = 0;
        if (live) { sx = L.s_x[p]; sy = L.s_y[p]; sz = L.s_z[p]; }
        NearestOut R;
        thread_nearest(M, s_ids[warp], s_lbs[warp], s_stage[warp], &s_mbar[warp], phase, lane, live, sx, sy, sz, max_d2, R);
        if (live) {
            L.c_tx[p] = R.tx; L.c_ty[p] = R.ty; L.c_tz[p] = R.tz;
#pragma unroll
            for (int j = 0; j < ICP_KX; ++j) {
                L.c_t2[(size_t)(3 * j) * cap + p] = R.t2[3 * j];
                L.c_t2[(size_t)(3 * j + 1) * cap + p] = R.t2[3 * j + 1];
                L.c_t2[(size_t)(3 * j + 2) * cap + p] = R.t2[3 * j + 2];
                L.c_ord2[(size_t)j * cap + p] = R.ord2[j];
            }
            L.c_px[p] = sx; L.c_py[p] = sy; L.c_pz[p] = sz;
            L.c_slack[p] = R.others;
            L.c_key[p] = R.qkey;
            L.c_ord[p] = R.ord;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&L.icp_searches, n_src);
}

// K4: the whole ICP loop (kiss-icp RegisterFrame) in one persistent cooperative kernel.
// grid = (blocks per lane, lanes), ICP_THREADS threads.  A block owns a contiguous range of
// 32-point groups and walks it in chunks of ICP_CHUNK groups (= one point per thread).  Per
// iteration and chunk:
//   1. thread per point: move the point by the last increment, then try the correspondence cache.
//      The last search of the point, at position p0, left its winner a, a runner-up b and a lower bound D
//      of the distance from p0 to every OTHER candidate (see warp_nearest).  If the point is still in the
//      same voxel (same 27 candidate voxels; the map does not change during the loop) and, with w the
//      lexicographically smaller of a and b in (distance^2, order id) - the search's own comparison -,
//      |p - w| + |p - p0| < D, then every other candidate c has |p - c| >= |p0 - c| - |p - p0| >= D - |p - p0|
//      > |p - w|: w is what a new search would return, so none is needed (a runner-up that has become the
//      nearest swaps places with a in the cache; ICP_KX runner-ups are kept); otherwise the point goes on
//      the block's work list;
//   2. warp per listed point: the pruned 27-voxel search, refreshing the cache entry;
//   3. thread per point: residual, Geman-McClure weight and the 16 distinct sums (+ count) in
//      registers; a warp IS a 32-point group, so xor-butterflies give the group partials directly.
// A counter barrier over the lane's blocks follows; after it EVERY block reduces the group partials
// with the same fixed tree, solves the 6x6 system and updates its copy of T_icp (identical code,
// identical bits), so one grid-wide hop per iteration is all the synchronisation there is.
// The cache changes which points are searched, never what a search would have returned, so the
// per-iteration correspondence sets stay bit-exact (tests compare them with the oracle's).
__global__ void __launch_bounds__(ICP_THREADS, PTK_ICP_MINBLOCKS) k_icp(LaneDev* lanes, const StepParams* params, StepOut* outs) {
    LaneDev& L = lanes[blockIdx.y];
    const StepParams& P = params[blockIdx.y];
    StepOut& O = outs[blockIdx.y];
    const int nblk = gridDim.x, b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_src = L.n_src;

    extern __shared__ double dyn_smem[];         // source points + cache entries of this block
    double* const ssx = dyn_smem;
    double* const ssy = ssx + ICP_SRC_CAP;
    double* const ssz = ssy + ICP_SRC_CAP;
    __shared__ double red[NSUM];
    __shared__ Rigid sE;
    __shared__ SE3q sT;
    __shared__ SolveSmem sS;
    __shared__ int s_done;
    __shared__ int s_nmiss;
    __shared__ int s_cnt;
    __shared__ MapView s_map;
    __shared__ unsigned short s_miss[ICP_CHUNK * 32];

    // ---- hash-sharded map: this context holds the voxels rank `sh_rank` of `sh_n` owns; the other
    // ranks run the same kernel on their GPUs and the blocks with the same index trade per-point records through
    // peer memory (NVLink): every rank writes its records straight into every other rank's buffer, then a stamp.
    const int sh_n = L.shard_n, sh_rank = L.shard_rank;
    const bool shard = sh_n > 1 && L.xch_rec[sh_rank] != nullptr;
    int n_vox_all = L.n_vox;
    if (shard) {
        __shared__ int s_nvox;
        if (threadIdx.x == 0) {
            const u64 st = ((u64)P.xch_epoch << 32) | 1ull;
            if (b == 0) {
                for (int r = 0; r < sh_n; ++r)
                    if (r != sh_rank) *((volatile int*)(L.xch_nvox[r] + sh_rank)) = L.n_vox;
                __threadfence_system();
                for (int r = 0; r < sh_n; ++r)
                    if (r != sh_rank) *((volatile u64*)(L.xch_flag[r] + sh_rank * XCH_FLAGS + (XCH_FLAGS - 1))) = st;
            }
            int tot = L.n_vox;
            for (int r = 0; r < sh_n; ++r) {
                if (r == sh_rank) continue;
                while (*((volatile u64*)(L.xch_flag[sh_rank] + r * XCH_FLAGS + (XCH_FLAGS - 1))) < st) { __nanosleep(100); }
                tot += *((volatile int*)(L.xch_nvox[sh_rank] + r));
            }
            s_nvox = tot;
        }
        __syncthreads();
        n_vox_all = s_nvox;
    }
    u32 xcnt = 0;             // exchanges this block has done in this launch
    if (n_vox_all == 0) {   // RegisterFrame: if (voxel_map.Empty()) return initial_guess;
        if (b == 0 && threadIdx.x == 0) {
            O.pose = se3q_matrix(se3q_mul(se3q_identity(), se3q_from_rigid(P.guess)));
            O.iterations = 0; O.n_corr = 0; O.status = 0; O.dx_norm = 0.0;
        }
        return;
    }
    const int n_groups = (n_src + 31) >> 5;
    // lane constants the loop would otherwise re-read from global memory after every barrier
    const double lane_eps = L.eps;
    const int lane_max_iters = L.max_iters, lane_trace_iters = L.trace_iters, lane_ng_cap = L.ng_cap, lane_cap_points = L.cap_points;
    double* const lane_part_a = L.part_a;
    double* const lane_part_b = L.part_b;
    // groups of this block: b, b + nblk, b + 2 nblk, ... (interleaved: the source is in beam order, and how far a
    // point moves per iteration - hence how often its cache entry misses - varies with the beam; dealing the
    // groups round-robin gives every block the same mix, so the blocks reach the barrier together)
    const int n_local = b < n_groups ? (n_groups - b + nblk - 1) / nblk : 0;
    const double max_corr = P.max_corr, kern = P.kernel;
    const double max_d2 = (max_corr * max_corr) * (1.0 + 1e-9);   // farther candidates are rejected anyway
    // the moving copy of this block's source points and their cache entries live in shared memory when they fit
    const bool in_smem = n_local * 32 <= ICP_SRC_CAP;
    // cache entry arrays: computed where used (one base + constants) instead of six live pointers
#define C_TX(i, g) (*(in_smem ? dyn_smem + 3 * ICP_SRC_CAP + (i) : L.c_tx + (g)))
#define C_TY(i, g) (*(in_smem ? dyn_smem + 4 * ICP_SRC_CAP + (i) : L.c_ty + (g)))
#define C_TZ(i, g) (*(in_smem ? dyn_smem + 5 * ICP_SRC_CAP + (i) : L.c_tz + (g)))
#define C_SLACK(i, g) (*(in_smem ? dyn_smem + 6 * ICP_SRC_CAP + (i) : L.c_slack + (g)))
#define C_KEY(i, g) (*(in_smem ? reinterpret_cast<u64*>(dyn_smem + 7 * ICP_SRC_CAP) + (i) : L.c_key + (g)))
    // runner-up j: coordinate c (0..2) at [11 + 3 j + c] * CAP; then the ints: ord, ord2[0..KX)
#define C_T2(j, c, i, g) (*(in_smem ? dyn_smem + (11 + 3 * (j) + (c)) * ICP_SRC_CAP + (i) \
                                    : L.c_t2 + ((size_t)(3 * (j) + (c)) * L.cap_points) + (g)))
#define C_ORD(i, g) (*(in_smem ? reinterpret_cast<int*>(dyn_smem + (11 + 3 * ICP_KX) * ICP_SRC_CAP) + (i) : L.c_ord + (g)))
#define C_ORD2(j, i, g) (*(in_smem ? reinterpret_cast<int*>(dyn_smem + (11 + 3 * ICP_KX) * ICP_SRC_CAP) + ((j) + 1) * ICP_SRC_CAP + (i) \
                                   : L.c_ord2 + (size_t)(j) * L.cap_points + (g)))
#define C_PX(i, g) (*(in_smem ? dyn_smem + 8 * ICP_SRC_CAP + (i) : L.c_px + (g)))
#define C_PY(i, g) (*(in_smem ? dyn_smem + 9 * ICP_SRC_CAP + (i) : L.c_py + (g)))
#define C_PZ(i, g) (*(in_smem ? dyn_smem + 10 * ICP_SRC_CAP + (i) : L.c_pz + (g)))
    if (threadIdx.x == 0) sT = se3q_identity();
    const double voxel = L.voxel_size, voxel_inv = L.voxel_inv;
    // phase clocks of block 0 (thread 0 only; six clock reads per iteration)
    const bool clk = b == 0 && threadIdx.x == 0;
    __shared__ long long s_cyc[6];
    __shared__ int s_searches;
    if (threadIdx.x < 6) s_cyc[threadIdx.x] = 0;
    if (threadIdx.x == 0) { s_searches = 0; s_cnt = 0; s_map = map_view(L); }
    long long tlast = clk ? clock64() : 0;
#define ICP_TICK(slot) do { if (clk) { const long long t_ = clock64(); s_cyc[slot] += t_ - tlast; tlast = t_; } } while (0)

    for (int it = 0;; ++it) {
        double* part = (it & 1) ? lane_part_b : lane_part_a;
        for (int k0 = 0; k0 < n_local; k0 += ICP_CHUNK) {
            const int gc = min(ICP_CHUNK, n_local - k0);   // local groups k0 .. k0 + gc, one per warp
            const int q = threadIdx.x;                     // point of this thread within the chunk
            const int grp = b + (k0 + warp) * nblk;        // this warp's group within the lane's source
            const int p = grp * 32 + lane;                 // this thread's point within the lane's source
            const int sp = k0 * 32 + q;                    // ... within the block
            const bool live = q < gc * 32 && p < n_src;
            if (threadIdx.x == 0) s_nmiss = 0;
            __syncthreads();                               // also: previous chunk / iteration fully consumed
            // ---- 1. move the point, consult the cache
            double sx = 0, sy = 0, sz = 0;
            bool miss = false;
            if (live) {
                if (it == 0 || !in_smem) { sx = __ldcg(L.s_x + p); sy = __ldcg(L.s_y + p); sz = __ldcg(L.s_z + p); }
                else { sx = ssx[sp]; sy = ssy[sp]; sz = ssz[sp]; }
                miss = true;
                if (it == 0) {
                    // every point's first search has been done by k_icp_search0: fetch its cache entry
                    if (in_smem) {
                        C_TX(sp, p) = L.c_tx[p]; C_TY(sp, p) = L.c_ty[p]; C_TZ(sp, p) = L.c_tz[p];
#pragma unroll
                        for (int j = 0; j < ICP_KX; ++j) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) C_T2(j, c, sp, p) = L.c_t2[(size_t)(3 * j + c) * L.cap_points + p];
                            C_ORD2(j, sp, p) = L.c_ord2[(size_t)j * L.cap_points + p];
                        }
                        C_PX(sp, p) = L.c_px[p]; C_PY(sp, p) = L.c_py[p]; C_PZ(sp, p) = L.c_pz[p];
                        C_SLACK(sp, p) = L.c_slack[p];
                        C_KEY(sp, p) = L.c_key[p];
                        C_ORD(sp, p) = L.c_ord[p];
                    }
                    miss = false;
                }
                if (it > 0) {
                    double xo, yo, zo;
                    rigid_apply(sE, sx, sy, sz, xo, yo, zo);
                    sx = xo; sy = yo; sz = zo;
                    const double others = C_SLACK(sp, p);
                    if (others > 0.0) {
                        const double mx = sx - C_PX(sp, p), my = sy - C_PY(sp, p), mz = sz - C_PZ(sp, p);
                        const double moved2 = (mx * mx + my * my) + mz * mz;
                        double ex = C_TX(sp, p) - sx, ey = C_TY(sp, p) - sy, ez = C_TZ(sp, p) - sz;
                        double da2 = (ex * ex + ey * ey) + ez * ez;              // as the search computes it
#pragma unroll
                        for (int j = 0; j < ICP_KX; ++j) {
                            const int oj = C_ORD2(j, sp, p);
                            if (oj < 0) continue;
                            const double bx = C_T2(j, 0, sp, p), by = C_T2(j, 1, sp, p), bz = C_T2(j, 2, sp, p);
                            ex = bx - sx; ey = by - sy; ez = bz - sz;
                            const double db2 = (ex * ex + ey * ey) + ez * ez;
                            const int o1 = C_ORD(sp, p);
                            if (db2 < da2 || (db2 == da2 && oj < o1)) {          // this runner-up has become the nearest
                                const double wx = C_TX(sp, p), wy = C_TY(sp, p), wz = C_TZ(sp, p);
                                C_TX(sp, p) = bx; C_TY(sp, p) = by; C_TZ(sp, p) = bz; C_ORD(sp, p) = oj;
                                C_T2(j, 0, sp, p) = wx; C_T2(j, 1, sp, p) = wy; C_T2(j, 2, sp, p) = wz; C_ORD2(j, sp, p) = o1;
                                da2 = db2;
                            }
                        }
                        // |p - w| + |p - p0| + 1e-9 < others, without the two square roots (the fp64 pipe is what this
                        // pass waits for): with o = others - 1e-9, a + m < o  <=>  o^2 - a^2 - m^2 > 0 and
                        // 4 a^2 m^2 < (o^2 - a^2 - m^2)^2; the factor keeps the test on the safe side of rounding
                        const double o1 = others - 1e-9;
                        const double rhs = (o1 * o1 - da2) - moved2;
                        if (o1 > 0.0 && rhs > 0.0 && 4.0 * (da2 * moved2) < (rhs * rhs) * (1.0 - 1e-12)) {
                            int kx, ky, kz;
                            voxel_key(sx, sy, sz, voxel, voxel_inv, kx, ky, kz);
                            miss = !(key_in_range(kx, ky, kz) && pack_key(kx, ky, kz) == C_KEY(sp, p));
                        }
                    }
                }
                if (in_smem) { ssx[sp] = sx; ssy[sp] = sy; ssz[sp] = sz; }
                else if (it > 0) { L.s_x[p] = sx; L.s_y[p] = sy; L.s_z[p] = sz; }
            }
            {
                const u32 mm = __ballot_sync(0xffffffffu, miss);
                if (mm) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_nmiss, __popc(mm));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (miss) s_miss[base + __popc(mm & ((1u << lane) - 1u))] = (unsigned short)q;
                }
            }
            __syncthreads();
            ICP_TICK(0);
            // ---- 2. full search of the listed points, one warp each (a HALF-warp per point - two queries per warp at
            // a time - measured slower: 1.24 against 1.02 ms per 48-lane launch; a search is bound by the length of
            // its own instruction chain, which the narrower form makes longer)
            const int nmiss = s_nmiss;
            if (threadIdx.x == 0) s_searches += nmiss;
            for (int i = warp; i < nmiss; i += ICP_WARPS) {
                const int mq = s_miss[i];
                const int msp = k0 * 32 + mq;
                const int mp = (b + (k0 + (mq >> 5)) * nblk) * 32 + (mq & 31);
                double qx, qy, qz;
                if (in_smem) { qx = ssx[msp]; qy = ssy[msp]; qz = ssz[msp]; }
                else { qx = __ldcg(L.s_x + mp); qy = __ldcg(L.s_y + mp); qz = __ldcg(L.s_z + mp); }
                double d2, tx, ty, tz, others;
                int ord;
                u64 qkey;
                double t2[3 * ICP_KX];
                int ord2[ICP_KX];
                const bool found = warp_nearest(s_map, qx, qy, qz, lane, max_d2, d2, ord, tx, ty, tz, others, &qkey, t2, ord2);
                if (lane == 0) {
                    C_TX(msp, mp) = tx; C_TY(msp, mp) = ty; C_TZ(msp, mp) = tz;
#pragma unroll
                    for (int j = 0; j < ICP_KX; ++j) {
                        C_T2(j, 0, msp, mp) = t2[3 * j]; C_T2(j, 1, msp, mp) = t2[3 * j + 1]; C_T2(j, 2, msp, mp) = t2[3 * j + 2];
                        C_ORD2(j, msp, mp) = ord2[j];
                    }
                    C_PX(msp, mp) = qx; C_PY(msp, mp) = qy; C_PZ(msp, mp) = qz;
                    C_SLACK(msp, mp) = others;
                    C_KEY(msp, mp) = qkey;
                    C_ORD(msp, mp) = found ? ord : -1;
                }
            }
            __syncthreads();
            ICP_TICK(1);
            // ---- 3. residual + weights, group partials by warp butterflies
            double wtx = 0, wty = 0, wtz = 0;
            int word = -1;
            if (live) {
                word = C_ORD(sp, p);
                if (word >= 0) { wtx = C_TX(sp, p); wty = C_TY(sp, p); wtz = C_TZ(sp, p); }
            }
            if (shard) {
                // what this rank found is the nearest point of ITS voxels; the correspondence is the lexicographic
                // minimum of (d2, order id) over the ranks (tie rule B.6: the order id encodes the voxel offset)
                const size_t cap = (size_t)L.cap_points;
                const int par = (int)(xcnt & 1u);
                double d2l = INFINITY, ordl = (double)SHARD_NO_ORD_K;
                if (live && word >= 0) {
                    const double dx = wtx - sx, dy = wty - sy, dz = wtz - sz;
                    d2l = (dx * dx + dy * dy) + dz * dz;
                    ordl = (double)word;
                }
                if (live) {
                    for (int r = 0; r < sh_n; ++r) {
                        if (r == sh_rank) continue;
                        double* rec = L.xch_rec[r] + (size_t)((sh_rank * 2 + par) * 5) * cap + p;
                        rec[0] = d2l; rec[cap] = ordl; rec[2 * cap] = wtx; rec[3 * cap] = wty; rec[4 * cap] = wtz;
                    }
                }
                __threadfence_system();
                __syncthreads();
                if (threadIdx.x == 0) {
                    const u64 st = ((u64)P.xch_epoch << 32) | (u64)(xcnt + 2u);
                    for (int r = 0; r < sh_n; ++r)
                        if (r != sh_rank) *((volatile u64*)(L.xch_flag[r] + sh_rank * XCH_FLAGS + b)) = st;
                    for (int r = 0; r < sh_n; ++r) {
                        if (r == sh_rank) continue;
                        while (*((volatile u64*)(L.xch_flag[sh_rank] + r * XCH_FLAGS + b)) < st) { }
                    }
                    __threadfence_system();
                }
                __syncthreads();
                if (live) {
                    for (int r = 0; r < sh_n; ++r) {
                        if (r == sh_rank) continue;
                        const double* rec = L.xch_rec[sh_rank] + (size_t)((r * 2 + par) * 5) * cap + p;
                        const double d2r = __ldcv(rec), ordr = __ldcv(rec + cap);
                        if (d2r < d2l || (d2r == d2l && ordr < ordl)) {
                            d2l = d2r; ordl = ordr;
                            wtx = __ldcv(rec + 2 * cap); wty = __ldcv(rec + 3 * cap); wtz = __ldcv(rec + 4 * cap);
                        }
                    }
                    word = ordl < (double)SHARD_NO_ORD_K ? (int)ordl : -1;
                }
                ++xcnt;
            }
            if (warp < gc) {
                double c[16];
                bool acc = false;
                int ord = -1;
                if (live) {
                    ord = word;
                    if (ord >= 0) {
                        const double tx = wtx, ty = wty, tz = wtz;
                        const double dx = tx - sx, dy = ty - sy, dz = tz - sz;
                        const double d2 = (dx * dx + dy * dy) + dz * dz;
                        // sqrt(d2) < max_corr: away from the threshold the squares decide (as in range_pass)
                        const double mc2 = max_corr * max_corr;
                        if (d2 < mc2 * (1.0 - 1e-12)) acc = true;
                        else if (d2 > mc2 * (1.0 + 1e-12)) acc = false;
                        else acc = sqrt(d2) < max_corr;
                        if (acc) lin_terms(sx, sy, sz, tx, ty, tz, kern, c);
                    }
                    if (it < lane_trace_iters) L.trace[(size_t)it * lane_cap_points + p] = acc ? ord : -1;
                }
                if (!acc) {
#pragma unroll
                    for (int v = 0; v < 16; ++v) c[v] = 0.0;
                }
                const double mine = warp_reduce16(c, lane);
                const u32 nacc = __popc(__ballot_sync(0xffffffffu, acc));   // a sum of 1.0s is exact in any order
                if (lane < 16) part[(size_t)(__brev((u32)lane) >> 28) * lane_ng_cap + grp] = mine;
                else if (lane == 16) part[(size_t)16 * lane_ng_cap + grp] = (double)nacc;
            }
        }
        // ---- one barrier over the lane's blocks (icp_arrive was zeroed by the previous kernel)
        __syncthreads();
        ICP_TICK(2);
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(&L.icp_arrive, 1u);
            const u32 target = (u32)nblk * (u32)(it + 1);
            while (*((volatile u32*)&L.icp_arrive) < target) { }
            __threadfence();
        }
        __syncthreads();
        ICP_TICK(3);
        // 16 sums, one warp each, with the canonical tree; the 17th (correspondence count) is a sum of
        // small integers - exact in any order - so all warps share it instead of one warp doing two trees
        for (int v = warp; v < 16; v += ICP_WARPS) {
            const double x = warp_tree_sum(part + (size_t)v * lane_ng_cap, n_groups, lane);
            if (lane == 0) red[v] = x;
        }
        {
            int cnt = 0;
            for (int g = threadIdx.x; g < n_groups; g += ICP_THREADS) cnt += (int)__ldcg(part + (size_t)16 * lane_ng_cap + g);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
        }
        __syncthreads();
        ICP_TICK(4);
        if (warp == 0) {
            if (lane == 0) { red[16] = (double)s_cnt; s_cnt = 0; }
            __syncwarp();
            icp_solve_step(L, P, O, red, &sS, &sE, &sT, &s_done, it, b == 0, lane, lane_eps, lane_max_iters);
        }
        __syncthreads();
        ICP_TICK(5);
        if (s_done) break;
    }
#undef ICP_TICK
    if (threadIdx.x == 0 && s_searches) atomicAdd(&L.icp_searches, s_searches);
#ifndef PTK_SOLVE_CLOCKS
    if (clk) {
#pragma unroll
        for (int k = 0; k < 6; ++k) O.icp_cyc[k] = s_cyc[k];
    }
#endif
#undef C_TX
#undef C_TY
#undef C_TZ
#undef C_SLACK
#undef C_KEY
#undef C_ORD
#undef C_ORD2
#undef C_T2
#undef C_PX
#undef C_PY
#undef C_PZ
}


// ------------------------------------------------------------------------------------
// K5: map insert, pass 1 (kiss-icp AddPoints): transform frame_downsample by the new pose,
// find-or-create the voxel, and run the atomicMin cascade that leaves, in vidx[voxel][count..],
// the smallest point indices of this scan in ascending order (= upstream's sequential insertion
// order under the canonical ordering).  Also clears table 2.
__device__ __forceinline__ u32 map_find_or_create(LaneDev& L, u64 key) {
    u32 slot = hash_map_key(key) & L.m_mask;
    u32 probes = 0;
    while (true) {
        volatile MapSlot* S = L.m_slots + slot;
        u64 k = S->key;
        if (k == KEY_EMPTY) {
            u64 prev = atomicCAS((u64*)&L.m_slots[slot].key, KEY_EMPTY, key);
            if (prev == KEY_EMPTY) {
                int t = atomicSub(&L.free_top, 1);
                u32 id;
                if (t > 0) id = L.freelist[t - 1];
                else id = (u32)atomicAdd(&L.bump, 1);
                if (id >= (u32)L.pool_cap) {
                    // no block left: flag the step, then hand the slot back as a tombstone so that no later
                    // launch finds a key without a voxel behind it (threads waiting for this slot's id leave
                    // their loop on ERR_POOL, which is set first)
                    atomicOr(&L.err, ERR_POOL);
                    __threadfence();
                    *((volatile u64*)&L.m_slots[slot].key) = KEY_TOMB;
                    atomicAdd(&L.n_tomb, 1);
                    return NONE;
                }
                VoxelBlock* B = L.blocks + id;
                double2* rows = reinterpret_cast<double2*>(B);   // unused slots read as +inf in the NN search
#pragma unroll
                for (int k = 0; k < (3 * MAXP) / 2; ++k) rows[k] = make_double2(INFINITY, INFINITY);
                B->key = key;
                L.vmeta[id].count = 0; L.vmeta[id].slot = slot;
                atomicAdd(&L.n_vox, 1);
                __threadfence();
                S->id = id;
                return id;
            }
            k = prev;
        }
        if (k == key) {
            u32 id;
            u32 spins = 0;
            do {
                id = S->id;
                if (id == NONE && ++spins > (1u << 24)) { atomicOr(&L.err, ERR_TABLE); break; }   // never wait unbounded
            } while (id == NONE && !(*((volatile int*)&L.err) & ERR_POOL));
            __threadfence();
            return id;
        }
        slot = (slot + 1) & L.m_mask;
        if (++probes > L.m_mask) { atomicOr(&L.err, ERR_TABLE); return NONE; }
    }
}

__global__ void __launch_bounds__(256) k_map_insert(LaneDev* lanes, const StepParams* params, const StepOut* outs, int use_pose) {
    LaneDev& L = lanes[blockIdx.y];
    const int n_ds = L.n_ds;
    Rigid T = use_pose ? outs[blockIdx.y].pose : rigid_identity();
    const int stride = gridDim.x * blockDim.x;
    const int n_round = ((n_ds + 31) / 32) * 32;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_round; j += stride) {
        bool act = j < n_ds;
        u64 key = KEY_EMPTY;
        if (act) {
            double x = L.ds_x[j], y = L.ds_y[j], z = L.ds_z[j];
            if (use_pose) { double xo, yo, zo; rigid_apply(T, x, y, z, xo, yo, zo); x = xo; y = yo; z = zo; }
            int kx, ky, kz;
            voxel_key(x, y, z, L.voxel_size, L.voxel_inv, kx, ky, kz);
            if (key_in_range(kx, ky, kz)) {
                key = pack_key(kx, ky, kz);
                if (L.shard_n > 1 && shard_owner(key, L.shard_n) != (u32)L.shard_rank) act = false;   // another rank's voxel
            } else { atomicOr(&L.err, ERR_KEYRANGE); act = false; }
            u32 s2 = L.ds_slot2[j];
            if (s2 != NONE) { L.t2_keys[s2] = KEY_EMPTY; L.t2_vals[s2] = NONE; L.ds_slot2[j] = NONE; }
        }
        u32 am = __ballot_sync(0xffffffffu, act);
        u32 vid = NONE;
        if (act) {
            u32 peers = __match_any_sync(am, key);
            int leader = __ffs(peers) - 1;
            if ((threadIdx.x & 31) == leader) vid = map_find_or_create(L, key);
            vid = __shfl_sync(peers, vid, leader);
        }
        if (j < n_ds) L.ds_vid[j] = vid;
        if (act && vid != NONE) {
            int c0 = (int)*((volatile u32*)&L.vmeta[vid].count);
            u32 cur = (u32)j;
            u32* slots = L.vidx + (size_t)vid * MAXP;
            for (int s = c0; s < L.maxp; ++s) {
                u32 old = atomicMin(slots + s, cur);
                if (old == NONE) break;
                if (old > cur) cur = old;
            }
        }
    }
}

// K6: map insert, pass 2: every point that survived the cascade writes itself into its slot.
__global__ void __launch_bounds__(256) k_map_commit(LaneDev* lanes, const StepOut* outs, int use_pose) {
    LaneDev& L = lanes[blockIdx.y];
    const int n_ds = L.n_ds;
    Rigid T = use_pose ? outs[blockIdx.y].pose : rigid_identity();
    const int stride = gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x == 0 && L.free_top < 0) L.free_top = 0;
    int added = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_ds; j += stride) {
        u32 vid = L.ds_vid[j];
        if (vid == NONE) continue;
        u32* slots = L.vidx + (size_t)vid * MAXP;
        int s = -1;
        for (int k = 0; k < L.maxp; ++k)
            if (slots[k] == (u32)j) { s = k; break; }
        if (s < 0) continue;
        double x = L.ds_x[j], y = L.ds_y[j], z = L.ds_z[j];
        if (use_pose) { double xo, yo, zo; rigid_apply(T, x, y, z, xo, yo, zo); x = xo; y = yo; z = zo; }
        VoxelBlock* B = L.blocks + vid;
        B->x[s] = x; B->y[s] = y; B->z[s] = z;
        if (s == 0) { L.vmeta[vid].fx = x; L.vmeta[vid].fy = y; L.vmeta[vid].fz = z; }
        slots[s] = NONE;
        atomicAdd(&L.vmeta[vid].count, 1u);
        ++added;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) added += __shfl_xor_sync(0xffffffffu, added, o);
    if ((threadIdx.x & 31) == 0 && added) atomicAdd(&L.map_points, added);
}

// K7: prune (kiss-icp RemovePointsFarFromLocation): erase every voxel whose FIRST point is
// farther than max_distance from `origin`.
__global__ void __launch_bounds__(256) k_map_prune(LaneDev* lanes, const StepOut* outs, const double* origin_override) {
    LaneDev& L = lanes[blockIdx.y];
    double ox, oy, oz;
    if (origin_override) { ox = origin_override[0]; oy = origin_override[1]; oz = origin_override[2]; }
    else { const Rigid& T = outs[blockIdx.y].pose; ox = T.t[0]; oy = T.t[1]; oz = T.t[2]; }
    const int bump = min(L.bump, L.pool_cap);    // failed allocations push the counter past the pool (ERR_POOL)
    const double r2max = L.max_distance * L.max_distance;
    const int stride = gridDim.x * blockDim.x;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < bump; u += stride) {
        VoxelMeta& V = L.vmeta[u];
        const u32 c = V.count;
        if (c == 0) continue;
        double dx = V.fx - ox, dy = V.fy - oy, dz = V.fz - oz;
        double d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 > r2max) {
            L.m_slots[V.slot].key = KEY_TOMB;
            V.count = 0;
            int t = atomicAdd(&L.free_top, 1);
            L.freelist[t] = (u32)u;
            atomicSub(&L.n_vox, 1);
            atomicAdd(&L.n_tomb, 1);
            atomicSub(&L.map_points, (int)c);
        }
    }
}

// K8: gather per-step counters into the output record and reset the per-step ones.
__global__ void k_finish(LaneDev* lanes, StepOut* outs, int n_lanes) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lanes) return;
    LaneDev& L = lanes[l];
    StepOut& O = outs[l];
    O.n_range = L.n_range; O.n_ds = L.n_ds; O.n_src = L.n_src;
    O.n_vox = L.n_vox; O.n_tomb = L.n_tomb; O.map_points = L.map_points;
    O.err = L.err; O.bump = L.bump; O.icp_searches = L.icp_searches; O.n_valid = L.n_valid;
    L.n_range = 0;
    L.n_valid = 0;
    L.err = 0;
    L.icp_searches = 0;
}

// Rebuild the map table without tombstones.
__global__ void k_map_rebuild(LaneDev* lanes) {
    LaneDev& L = lanes[blockIdx.y];
    const int bump = min(L.bump, L.pool_cap);    // failed allocations push the counter past the pool (ERR_POOL)
    const int stride = gridDim.x * blockDim.x;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < bump; u += stride) {
        VoxelBlock* B = L.blocks + u;
        if (L.vmeta[u].count == 0) continue;
        u64 key = B->key;
        u32 slot = hash_map_key(key) & L.m_mask;
        while (true) {
            u64 prev = atomicCAS((u64*)&L.m_slots[slot].key, KEY_EMPTY, key);
            if (prev == KEY_EMPTY) break;
            slot = (slot + 1) & L.m_mask;
        }
        L.m_slots[slot].id = (u32)u;
        L.vmeta[u].slot = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) L.n_tomb = 0;
}

// ---- hash-sharded map over several GPUs ---------------------------------
// One ICP iteration is split over three launches with two collectives between them (the host drives
// the loop, ptudes_lab_b200/sharded.py):
//   k_shard_search      nearest LOCAL map point of every source point -> record (d2, order id, xyz)
//   [all-gather of the records]
//   k_shard_system      global nearest = lexicographic min of (d2, order id) over the ranks' records;
//                       residual terms of this rank's slice of the source; k_shard_slice_root reduces
//                       the slice with the canonical tree (the slice is an aligned subtree)
//   [all-reduce of the [17][32] partial table, every rank filling only its own column: x + 0 is exact]
//   k_shard_solve       canonical tree over the slice roots, 6x6 solve, SE3 exp, T_icp update
// Every rank ends up with bit-identical sums, hence bit-identical poses - and identical to one GPU.
constexpr int SHARD_NO_ORD = 1 << 30;

__global__ void k_shard_search(LaneDev* lanes, int lane_id, const StepParams* params, int it, double* rec, int n) {
    LaneDev& L = lanes[lane_id];
    const StepParams& P = params[lane_id];
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n) return;
    double sx = __ldcg(L.s_x + p), sy = __ldcg(L.s_y + p), sz = __ldcg(L.s_z + p);
    if (it > 0) {
        double xo, yo, zo;
        rigid_apply(L.icp_E, sx, sy, sz, xo, yo, zo);
        sx = xo; sy = yo; sz = zo;
    }
    const double max_d2 = (P.max_corr * P.max_corr) * (1.0 + 1e-9);
    double d2 = INFINITY, tx = 0, ty = 0, tz = 0, others;
    int ord = SHARD_NO_ORD;
    const bool found = L.n_vox > 0 && warp_nearest(map_view(L), sx, sy, sz, lane, max_d2, d2, ord, tx, ty, tz, others);
    if (lane == 0) {
        if (it > 0) { L.s_x[p] = sx; L.s_y[p] = sy; L.s_z[p] = sz; }
        rec[p] = found ? d2 : INFINITY;
        rec[(size_t)n + p] = found ? (double)ord : (double)SHARD_NO_ORD;
        rec[2 * (size_t)n + p] = tx; rec[3 * (size_t)n + p] = ty; rec[4 * (size_t)n + p] = tz;
    }
}

__global__ void __launch_bounds__(ICP_THREADS) k_shard_system(LaneDev* lanes, int lane_id, const StepParams* params,
                                                              const double* gathered, int G, int n, int it, int g_lo, int g_cnt) {
    LaneDev& L = lanes[lane_id];
    const StepParams& P = params[lane_id];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gi = blockIdx.x * ICP_WARPS + warp;
    const int n_groups = (n + 31) >> 5;
    const int g = g_lo + gi;
    if (gi >= g_cnt || g >= n_groups) return;
    const int p = g * 32 + lane;
    double c[16];
    bool acc = false;
    if (p < n) {
        double bd2 = INFINITY, bord = (double)SHARD_NO_ORD;
        int br = 0;
        for (int r = 0; r < G; ++r) {
            const double* R = gathered + (size_t)r * 5 * n;
            const double d2 = R[p], ord = R[(size_t)n + p];
            if (d2 < bd2 || (d2 == bd2 && ord < bord)) { bd2 = d2; bord = ord; br = r; }
        }
        const double sx = __ldcg(L.s_x + p), sy = __ldcg(L.s_y + p), sz = __ldcg(L.s_z + p);
        if (bord < (double)SHARD_NO_ORD) {
            const double* R = gathered + (size_t)br * 5 * n;
            const double tx = R[2 * (size_t)n + p], ty = R[3 * (size_t)n + p], tz = R[4 * (size_t)n + p];
            const double dx = tx - sx, dy = ty - sy, dz = tz - sz;
            acc = sqrt((dx * dx + dy * dy) + dz * dz) < P.max_corr;
            if (acc) lin_terms(sx, sy, sz, tx, ty, tz, P.kernel, c);
        }
        if (it < L.trace_iters) L.trace[(size_t)it * L.cap_points + p] = acc ? (int)bord : -1;
    }
    if (!acc) {
#pragma unroll
        for (int v = 0; v < 16; ++v) c[v] = 0.0;
    }
    const double mine = warp_reduce16(c, lane);
    const u32 nacc = __popc(__ballot_sync(0xffffffffu, acc));
    if (lane < 16) L.part_a[(size_t)(__brev((u32)lane) >> 28) * L.ng_cap + g] = mine;
    else if (lane == 16) L.part_a[(size_t)16 * L.ng_cap + g] = (double)nacc;
}

// root of this rank's slice [g_lo, g_lo + g_cnt) of the group partials -> column `col` of out[17][32]
__global__ void __launch_bounds__(NSUM * 32) k_shard_slice_root(LaneDev* lanes, int lane_id, int n, int g_lo, int g_cnt, double* out, int col) {
    LaneDev& L = lanes[lane_id];
    const int v = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_groups = (n + 31) >> 5;
    const int n_valid = max(0, min(g_cnt, n_groups - g_lo));
    double x = 0.0;
    if (n_valid > 0) x = warp_tree_sum(L.part_a + (size_t)v * L.ng_cap + g_lo, n_valid, lane);
    if (lane == 0) out[v * 32 + col] = x;
}

__global__ void __launch_bounds__(NSUM * 32) k_shard_solve(LaneDev* lanes, int lane_id, const StepParams* params, StepOut* outs, const double* partials,
                              int n_roots, int it, int map_empty) {
    LaneDev& L = lanes[lane_id];
    const StepParams& P = params[lane_id];
    StepOut& O = outs[lane_id];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ double red[NSUM];
    __shared__ Rigid sE;
    __shared__ SE3q sT;
    __shared__ SolveSmem sS;
    __shared__ int s_done;
    if (map_empty) {          // RegisterFrame: if (voxel_map.Empty()) return initial_guess;
        if (threadIdx.x == 0) {
            O.pose = se3q_matrix(se3q_mul(se3q_identity(), se3q_from_rigid(P.guess)));
            O.iterations = 0; O.n_corr = 0; O.status = 0; O.dx_norm = 0.0;
            L.icp_done = 1;
        }
        return;
    }
    if (warp < NSUM) {
        const double x = warp_tree_sum(partials + warp * 32, n_roots, lane);
        if (lane == 0) red[warp] = x;
    }
    if (threadIdx.x == 0) sT = it == 0 ? se3q_identity() : L.icp_Tq;
    __syncthreads();
    if (warp == 0) icp_solve_step(L, P, O, red, &sS, &sE, &sT, &s_done, it, true, lane, L.eps, L.max_iters);
    __syncthreads();
    if (threadIdx.x == 0) { L.icp_E = sE; L.icp_Tq = sT; L.icp_done = s_done; }
}

// ---- stand-alone pieces -------------------------------------------------------------
__global__ void k_deskew(const double* xyz, const double* ts, int n, StepParams P, double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    P.xyz = xyz; P.ts = ts;
    double x, y, z;
    load_point(P, i, x, y, z);
    out[3 * (size_t)i] = x; out[3 * (size_t)i + 1] = y; out[3 * (size_t)i + 2] = z;
}

// gather SoA (x,y,z) -> AoS rows, optional index array
__global__ void k_gather_aos(const double* x, const double* y, const double* z, int n, double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[3 * (size_t)i] = x[i]; out[3 * (size_t)i + 1] = y[i]; out[3 * (size_t)i + 2] = z[i];
}

__global__ void k_scatter_soa(const double* in, int n, double* x, double* y, double* z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = in[3 * (size_t)i]; y[i] = in[3 * (size_t)i + 1]; z[i] = in[3 * (size_t)i + 2];
}

// load an external cloud as "frame_downsample" (for map_add_points / map_update taps)
__global__ void k_load_ds(LaneDev* lanes, int lane, const double* in, int n) {
    LaneDev& L = lanes[lane];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) L.n_ds = n;
    if (i >= n) return;
    L.ds_x[i] = in[3 * (size_t)i]; L.ds_y[i] = in[3 * (size_t)i + 1]; L.ds_z[i] = in[3 * (size_t)i + 2];
    L.ds_slot2[i] = NONE;
}

// load an external cloud as "source" transformed by guess (register_point_cloud tap)
__global__ void k_load_src(LaneDev* lanes, int lane, const double* in, int n, Rigid guess) {
    LaneDev& L = lanes[lane];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { L.n_src = n; L.icp_arrive = 0; }
    if (i >= n) return;
    double x = in[3 * (size_t)i], y = in[3 * (size_t)i + 1], z = in[3 * (size_t)i + 2];
    L.s0_x[i] = x; L.s0_y[i] = y; L.s0_z[i] = z;
    double xo, yo, zo;
    rigid_apply(guess, x, y, z, xo, yo, zo);
    L.s_x[i] = xo; L.s_y[i] = yo; L.s_z[i] = zo;
    L.s_idx[i] = (u32)i;
}

// _get_correspondences tap: one warp per query
__global__ void k_correspondences(LaneDev* lanes, int lane_id, const double* q, int n, double max_dist,
                                  int* out_order, double* out_target, int* n_corr) {
    LaneDev& L = lanes[lane_id];
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    double sx = q[3 * (size_t)w], sy = q[3 * (size_t)w + 1], sz = q[3 * (size_t)w + 2];
    double d2, tx, ty, tz, others;
    int ord;
    bool found = L.n_vox > 0 && warp_nearest(map_view(L), sx, sy, sz, lane, (max_dist * max_dist) * (1.0 + 1e-9), d2, ord, tx, ty, tz, others);
    bool acc = found && (sqrt(d2) < max_dist);
    if (lane == 0) {
        out_order[w] = acc ? ord : -1;
        out_target[3 * (size_t)w] = acc ? tx : 0.0;
        out_target[3 * (size_t)w + 1] = acc ? ty : 0.0;
        out_target[3 * (size_t)w + 2] = acc ? tz : 0.0;
        if (acc) atomicAdd(n_corr, 1);
    }
}

// map dump: compact live voxels (order unspecified)
__global__ void k_map_dump(LaneDev* lanes, int lane_id, int* keys, int* counts, double* points, double* cloud,
                           int capacity, int* n_out, int* n_pts_out) {
    LaneDev& L = lanes[lane_id];
    const int stride = gridDim.x * blockDim.x;
    const int bump = min(L.bump, L.pool_cap);
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < bump; u += stride) {
        const VoxelBlock* B = L.blocks + u;
        int c = (int)L.vmeta[u].count;
        if (c == 0) continue;
        if (cloud) {
            int at = atomicAdd(n_pts_out, c);
            for (int s = 0; s < c; ++s) {
                if (at + s < capacity) {
                    cloud[3 * (size_t)(at + s)] = B->x[s]; cloud[3 * (size_t)(at + s) + 1] = B->y[s];
                    cloud[3 * (size_t)(at + s) + 2] = B->z[s];
                }
            }
        } else {
            int at = atomicAdd(n_out, 1);
            if (at < capacity) {
                u64 k = B->key;
                keys[3 * at] = (int)((k >> 42) & 0x1FFFFF) - KEY_BIAS;
                keys[3 * at + 1] = (int)((k >> 21) & 0x1FFFFF) - KEY_BIAS;
                keys[3 * at + 2] = (int)(k & 0x1FFFFF) - KEY_BIAS;
                counts[at] = c;
                for (int s = 0; s < MAXP; ++s) {
                    points[((size_t)at * MAXP + s) * 3] = s < c ? B->x[s] : 0.0;
                    points[((size_t)at * MAXP + s) * 3 + 1] = s < c ? B->y[s] : 0.0;
                    points[((size_t)at * MAXP + s) * 3 + 2] = s < c ? B->z[s] : 0.0;
                }
            }
        }
    }
}

}  // namespace ptk
