// libptk: host side of the odometry step + the C ABI declared in include/ptk.h.
//
// The host keeps what kiss_icp.kiss_icp.KissICP keeps (pose list, adaptive threshold) and what
// /root/reference/src/ptudes/kiss.py:83-131 computes around the kiss-icp calls (initial guess,
// pose gain metrics); all per-point work runs in the kernels of ptk_device.cuh.
#include <cuda_runtime.h>
#include <sched.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <string>
#include <vector>

#include "../../include/ptk.h"
#include "ptk_device.cuh"

using namespace ptk;

namespace {

struct Threshold {   // kiss-icp AdaptiveThreshold
    double sse2 = 0.0;
    int num = 0;
    Rigid deviation = rigid_identity();
};

struct LaneHost {
    LaneDev d;                        // host mirror of the device struct (pointers + config)
    std::vector<Rigid> poses;
    Threshold thr;
    double last_sigma = 0.0;
    double pending_sigma = 0.0;       // threshold chosen by step_prepare for the step in flight
    int pending_nmax = 0, shard_nsrc = 0; // sharded loop state between ptk_shard_begin and ptk_shard_end
    StepParams last_params;
    StepOut last_out;
    bool have_last = false;
    bool failed = false;              // a capacity error left the local map incomplete: the lane refuses steps until ptk_reset
    int last_reg_iters = 0, last_reg_nsrc = 0;
    u32 epoch = 0, tbase1 = 0, tbase2 = 0, release_base = 0;
    double* in_xyz = nullptr;         // staging for host inputs
    double* in_ts = nullptr;
    u32* in_range = nullptr;          // staging for host range images (buffer 0)
    u32* pf_range[2] = {nullptr, nullptr};   // double-buffered staging: one feeds the step, one receives the prefetch
    const void* pf_src = nullptr;     // host image whose copy into pf_range[pf_buf] is in flight / done
    int pf_buf = 0, cur_buf = 0;
    double* col_motion = nullptr;     // [12][W] per-column deskew motion (range-image mode)
    std::vector<void*> allocs;
};

}  // namespace

enum ProfSlot { PS_SCAN_INSERT = 0, PS_COMPACT1, PS_COMPACT2, PS_ICP, PS_MAP_INSERT, PS_MAP_COMMIT, PS_MAP_PRUNE,
                PS_FINISH, PS_REBUILD, PS_OTHER, PS_COL_MOTION, PS_SEARCH0, PS_COUNT };
static_assert(PS_COUNT <= PTK_PROF_SLOTS, "profile slots");
static const char* const kProfNames[PS_COUNT] = {"k_scan_insert", "k_compact1", "k_compact2", "k_icp", "k_map_insert",
                                                 "k_map_commit", "k_map_prune", "k_finish", "k_map_rebuild", "other", "k_col_motion",
                                                 "k_icp_search0"};

struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> ev;      // event pool, two per bracketed launch
    int used = 0;
    struct Rec { int slot, e0; };
    std::vector<Rec> recs;
    double ms[PTK_PROF_SLOTS] = {0};
    long long n[PTK_PROF_SLOTS] = {0};
};

struct ptk_ctx {
    Prof prof;
    long long launches = 0;
    int device = 0;
    ptk_config cfg;
    int B = 1;                        // lanes; lane index B is the scratch lane of the stand-alone calls
    std::vector<LaneHost> lanes;
    LaneDev* d_lanes = nullptr;
    StepParams* d_params = nullptr;
    StepOut* d_outs = nullptr;
    StepParams* h_params = nullptr;   // pinned
    StepOut* h_outs = nullptr;        // pinned
    double* d_tmp = nullptr;          // scratch for taps (cap_points*3 doubles)
    int* d_tmp_i = nullptr;           // scratch ints (cap_points + 16)
    size_t big_tmp_bytes = 0;
    void* d_big = nullptr;            // lazily allocated scratch for map dumps
    // sensor model of the range-image entry points (ptk_set_sensor)
    int sen_H = 0, sen_W = 0;
    double sen_unit = 0.001;
    double* d_lut_dir = nullptr;
    double* d_lut_off = nullptr;
    double* d_col_ts = nullptr;
    std::vector<void*> sensor_allocs;
    cudaStream_t copy_stream = nullptr;      // H2D prefetch of the next scans (ptk_prefetch_scan_batch)
    cudaEvent_t pf_event = nullptr;
    std::vector<const unsigned int*> pf_pending;   // host images to copy during the next step
    int num_sms = 148;
    int icp_blocks_total = 148;
    int icp_max_blocks_per_lane = 1 << 20;   // PTK_ICP_MAX_BLOCKS_PER_LANE: fewer blocks = cheaper barrier, slower searches
    int icp_cluster = 0;              // blocks per lane of the cluster launch of wide batches (0: not available)
    int icp_cluster_min_lanes = 56;   // batch width from which the cluster launch is used
    bool blocking_sync = false;       // wait for a step without monopolising a core (yielding poll or blocking-sync event):
    cudaEvent_t sync_event = nullptr; // set by ptk_fleet_replay, whose worker threads may outnumber the host cores
    // hash-sharded mode with in-kernel exchange: this rank's buffer and every rank's buffer as mapped here
    void* xch_local = nullptr;
    size_t xch_bytes = 0;
    void* xch_peer[PTK_MAX_PEERS] = {nullptr};
    bool xch_ipc[PTK_MAX_PEERS] = {false};
    std::string err;
    std::vector<void*> allocs;
};

static thread_local std::string g_create_err;

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return PTK_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

// ---- launch bookkeeping: count every kernel; with profiling on bracket it with events ----
static int prof_begin(ptk_ctx* ctx, int slot, cudaStream_t st) {
    ctx->launches++;
    Prof& p = ctx->prof;
    if (!p.on) return -1;
    if (p.used + 2 > (int)p.ev.size()) {
        for (int k = 0; k < 2; ++k) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return -1;
            p.ev.push_back(e);
        }
    }
    int e0 = p.used;
    p.used += 2;
    cudaEventRecord(p.ev[e0], st);
    p.recs.push_back({slot, e0});
    return e0;
}
static void prof_end(ptk_ctx* ctx, int e0, cudaStream_t st) {
    if (e0 >= 0) cudaEventRecord(ctx->prof.ev[e0 + 1], st);
}
// call after the stream has been synchronised
static void prof_collect(ptk_ctx* ctx) {
    Prof& p = ctx->prof;
    for (auto& r : p.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.ev[r.e0], p.ev[r.e0 + 1]) == cudaSuccess) { p.ms[r.slot] += ms; p.n[r.slot]++; }
        else cudaGetLastError();
    }
    p.recs.clear();
    p.used = 0;
}
#define LAUNCH(slot, st, ...)                          \
    do {                                               \
        int pe_ = prof_begin(ctx, slot, st);           \
        __VA_ARGS__;                                   \
        prof_end(ctx, pe_, st);                        \
    } while (0)

static int fail(ptk_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

static u32 next_pow2(u32 v) {
    u32 p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename T>
static cudaError_t dalloc(std::vector<void*>& owner, T** p, size_t count, int fill = -1) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    owner.push_back(q);
    *p = (T*)q;
    if (fill >= 0) e = cudaMemset(q, fill, std::max<size_t>(count, 1) * sizeof(T));
    return e;
}

static bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// -------------------------------------------------------------------------------------
extern "C" void ptk_default_config(ptk_config* c) {
    if (!c) return;
    c->max_range = 100.0;
    c->min_range = 5.0;
    c->voxel_size = 0.0;
    c->max_points_per_voxel = 20;
    c->deskew = 1;
    c->initial_threshold = 2.0;
    c->min_motion_th = 0.1;
    c->max_iterations = 500;
    c->convergence_eps = 1e-4;
    c->max_points = 262144;
    c->map_capacity = 262144;
    c->batch = 1;
    c->trace_iterations = 0;
}

extern "C" int ptk_version(void) { return PTK_VERSION; }

static int lane_alloc(ptk_ctx* ctx, LaneHost& LH, bool scratch) {
    const ptk_config& c = ctx->cfg;
    LaneDev& d = LH.d;
    memset(&d, 0, sizeof(d));
    d.voxel_size = c.voxel_size;
    d.voxel_inv = 1.0 / c.voxel_size;
    d.max_distance = c.max_range;
    d.maxp = c.max_points_per_voxel;
    d.max_iters = c.max_iterations;
    d.eps = c.convergence_eps;
    d.cap_points = c.max_points;
    d.pool_cap = scratch ? 1 : c.map_capacity;
    d.trace_iters = scratch ? 0 : c.trace_iterations;
    d.ng_cap = (int)next_pow2((u32)((c.max_points + 31) / 32));
    u32 tcap = next_pow2((u32)(2 * (size_t)c.max_points));
    u32 mcap = next_pow2((u32)(4 * (size_t)d.pool_cap));
    d.t_mask = tcap - 1;
    d.m_mask = mcap - 1;
    size_t N = (size_t)c.max_points;
    auto& A = LH.allocs;
    CK(dalloc(A, &d.t1_keys, tcap, 0xFF));
    CK(dalloc(A, &d.t1_vals, tcap, 0xFF));
    CK(dalloc(A, &d.t2_keys, tcap, 0xFF));
    CK(dalloc(A, &d.t2_vals, tcap, 0xFF));
    CK(dalloc(A, &d.slot1, N, 0xFF));
    size_t ntile = (N + TILE - 1) / TILE + 1;
    CK(dalloc(A, &d.agg1, ntile, 0));
    CK(dalloc(A, &d.agg2, ntile, 0));
    CK(dalloc(A, &d.ds_x, N)); CK(dalloc(A, &d.ds_y, N)); CK(dalloc(A, &d.ds_z, N));
    CK(dalloc(A, &d.ds_idx, N, 0)); CK(dalloc(A, &d.ds_slot2, N, 0xFF)); CK(dalloc(A, &d.ds_vid, N, 0xFF));
    CK(dalloc(A, &d.s0_x, N)); CK(dalloc(A, &d.s0_y, N)); CK(dalloc(A, &d.s0_z, N));
    CK(dalloc(A, &d.s_x, N)); CK(dalloc(A, &d.s_y, N)); CK(dalloc(A, &d.s_z, N));
    CK(dalloc(A, &d.s_idx, N, 0));
    CK(dalloc(A, &d.m_slots, mcap, 0xFF));
    CK(dalloc(A, &d.blocks, (size_t)d.pool_cap, 0));
    CK(dalloc(A, &d.vmeta, (size_t)d.pool_cap, 0));
    CK(dalloc(A, &d.vidx, (size_t)d.pool_cap * MAXP, 0xFF));
    CK(dalloc(A, &d.freelist, (size_t)d.pool_cap, 0));
    CK(dalloc(A, &d.part_a, (size_t)NRED * d.ng_cap, 0));
    CK(dalloc(A, &d.part_b, (size_t)NRED * d.ng_cap, 0));
    CK(dalloc(A, &d.c_tx, N)); CK(dalloc(A, &d.c_ty, N)); CK(dalloc(A, &d.c_tz, N)); CK(dalloc(A, &d.c_slack, N));
    CK(dalloc(A, &d.c_key, N)); CK(dalloc(A, &d.c_ord, N));
    CK(dalloc(A, &d.c_px, N)); CK(dalloc(A, &d.c_py, N)); CK(dalloc(A, &d.c_pz, N));
    CK(dalloc(A, &d.c_t2, N * 3 * ICP_KX)); CK(dalloc(A, &d.c_ord2, N * ICP_KX));
    CK(dalloc(A, &d.trace, (size_t)std::max(d.trace_iters, 1) * N, 0xFF));
    CK(dalloc(A, &LH.in_xyz, N * 3));
    CK(dalloc(A, &LH.in_ts, N));
    CK(dalloc(A, &LH.in_range, N));
    LH.pf_range[0] = LH.in_range;
    CK(dalloc(A, &LH.pf_range[1], N));
    d.icp_E = rigid_identity();
    d.icp_T = rigid_identity();
    return PTK_OK;
}

extern "C" int ptk_ctx_create(ptk_ctx** out, int device, const ptk_config* cfg_in) {
    if (!out) return PTK_E_ARG;
    *out = nullptr;
    ptk_config cfg;
    if (cfg_in) cfg = *cfg_in; else ptk_default_config(&cfg);
    if (cfg.voxel_size <= 0.0) cfg.voxel_size = cfg.max_range / 100.0;
    if (cfg.max_points_per_voxel < 1 || cfg.max_points_per_voxel > MAXP || cfg.batch < 1 ||
        cfg.max_points < 1 || cfg.map_capacity < 1 || cfg.max_iterations < 1 || cfg.trace_iterations < 0 ||
        cfg.max_points > (1 << 24)) {
        g_create_err = "ptk_ctx_create: bad config";
        return PTK_E_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        g_create_err = std::string("ptk_ctx_create: no usable CUDA device (") +
                       (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") + ")";
        return PTK_E_CUDA;
    }
    ptk_ctx* ctx = new ptk_ctx();
    ctx->device = device;
    ctx->cfg = cfg;
    ctx->B = cfg.batch;
    auto bail = [&](int code) {
        g_create_err = ctx->err;
        ptk_ctx_destroy(ctx);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(PTK_E_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { ctx->err = "cudaGetDeviceProperties failed"; return bail(PTK_E_CUDA); }
    ctx->num_sms = prop.multiProcessorCount;
    int occ = 0;
    if (cudaFuncSetAttribute(k_icp, cudaFuncAttributeMaxDynamicSharedMemorySize, ICP_SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_icp, ICP_THREADS, ICP_SMEM) != cudaSuccess || occ < 1) {
        ctx->err = std::string("k_icp not launchable on this device: ") + cudaGetErrorString(cudaGetLastError());
        return bail(PTK_E_CUDA);
    }
    ctx->icp_blocks_total = occ * ctx->num_sms;
    {   // cluster size for wide batches: PTK_ICP_CLUSTER (1 disables), default 8 = the portable maximum
        if (const char* mb = getenv("PTK_ICP_MAX_BLOCKS_PER_LANE")) ctx->icp_max_blocks_per_lane = std::max(1, atoi(mb));
        const char* e = getenv("PTK_ICP_CLUSTER");
        int want = e ? atoi(e) : 8;
        if (const char* m = getenv("PTK_ICP_CLUSTER_MIN_LANES")) ctx->icp_cluster_min_lanes = atoi(m);
        ctx->icp_cluster = 0;
        if (want > 8 && cudaFuncSetAttribute(k_icp, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) want = 8;
        if (want > 1) {
            cudaLaunchConfig_t qc;
            memset(&qc, 0, sizeof(qc));
            qc.gridDim = dim3(want, 1); qc.blockDim = dim3(ICP_THREADS); qc.dynamicSmemBytes = ICP_SMEM;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = want; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k_icp, &qc) == cudaSuccess && ncl >= 1) ctx->icp_cluster = want;
        }
        cudaGetLastError();
    }
    ctx->lanes.resize(ctx->B + 1);
    for (int l = 0; l <= ctx->B; ++l) {
        int rc = lane_alloc(ctx, ctx->lanes[l], l == ctx->B);
        if (rc != PTK_OK) return bail(rc);
    }
    auto ck = [&](cudaError_t ee, const char* what) {
        if (ee != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(ee); return false; }
        return true;
    };
    int nl = ctx->B + 1;
    if (!ck(dalloc(ctx->allocs, &ctx->d_lanes, nl), "alloc lanes")) return bail(PTK_E_CUDA);
    if (!ck(dalloc(ctx->allocs, &ctx->d_params, nl, 0), "alloc params")) return bail(PTK_E_CUDA);
    if (!ck(dalloc(ctx->allocs, &ctx->d_outs, nl, 0), "alloc outs")) return bail(PTK_E_CUDA);
    if (!ck(dalloc(ctx->allocs, &ctx->d_tmp, (size_t)cfg.max_points * 3 + 64), "alloc tmp")) return bail(PTK_E_CUDA);
    if (!ck(dalloc(ctx->allocs, &ctx->d_tmp_i, (size_t)cfg.max_points + 64, 0), "alloc tmp_i")) return bail(PTK_E_CUDA);
    if (!ck(cudaMallocHost((void**)&ctx->h_params, sizeof(StepParams) * nl), "pinned params")) return bail(PTK_E_CUDA);
    if (!ck(cudaMallocHost((void**)&ctx->h_outs, sizeof(StepOut) * nl), "pinned outs")) return bail(PTK_E_CUDA);
    memset(ctx->h_params, 0, sizeof(StepParams) * nl);
    memset(ctx->h_outs, 0, sizeof(StepOut) * nl);
    for (int l = 0; l < nl; ++l)
        if (!ck(cudaMemcpy(ctx->d_lanes + l, &ctx->lanes[l].d, sizeof(LaneDev), cudaMemcpyHostToDevice), "upload lane")) return bail(PTK_E_CUDA);
    if (!ck(cudaDeviceSynchronize(), "create sync")) return bail(PTK_E_CUDA);
    *out = ctx;
    return PTK_OK;
}

extern "C" int ptk_ctx_destroy(ptk_ctx* ctx) {
    if (!ctx) return PTK_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& L : ctx->lanes)
        for (void* p : L.allocs) cudaFree(p);
    for (void* p : ctx->allocs) cudaFree(p);
    for (void* p : ctx->sensor_allocs) cudaFree(p);
    for (int r = 0; r < PTK_MAX_PEERS; ++r)
        if (ctx->xch_ipc[r] && ctx->xch_peer[r]) cudaIpcCloseMemHandle(ctx->xch_peer[r]);
    if (ctx->xch_local) cudaFree(ctx->xch_local);
    if (ctx->d_big) cudaFree(ctx->d_big);
    for (cudaEvent_t e : ctx->prof.ev) cudaEventDestroy(e);
    if (ctx->pf_event) cudaEventDestroy(ctx->pf_event);
    if (ctx->sync_event) cudaEventDestroy(ctx->sync_event);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->h_params) cudaFreeHost(ctx->h_params);
    if (ctx->h_outs) cudaFreeHost(ctx->h_outs);
    delete ctx;
    return PTK_OK;
}

extern "C" const char* ptk_last_error(const ptk_ctx* ctx) {
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

// reset the device-side dynamic state of one lane's map (and scan tables)
static int lane_reset_device(ptk_ctx* ctx, int l, cudaStream_t st, bool tables) {
    LaneHost& LH = ctx->lanes[l];
    LaneDev& d = LH.d;
    CK(cudaMemsetAsync(d.m_slots, 0xFF, ((size_t)d.m_mask + 1) * sizeof(MapSlot), st));
    CK(cudaMemsetAsync(d.blocks, 0, (size_t)d.pool_cap * sizeof(VoxelBlock), st));
    CK(cudaMemsetAsync(d.vmeta, 0, (size_t)d.pool_cap * sizeof(VoxelMeta), st));
    CK(cudaMemsetAsync(d.vidx, 0xFF, (size_t)d.pool_cap * MAXP * sizeof(u32), st));
    if (tables) {
        size_t tcap = (size_t)d.t_mask + 1;
        CK(cudaMemsetAsync(d.t1_keys, 0xFF, tcap * sizeof(u64), st));
        CK(cudaMemsetAsync(d.t1_vals, 0xFF, tcap * sizeof(u32), st));
        CK(cudaMemsetAsync(d.t2_keys, 0xFF, tcap * sizeof(u64), st));
        CK(cudaMemsetAsync(d.t2_vals, 0xFF, tcap * sizeof(u32), st));
        CK(cudaMemsetAsync(d.ds_slot2, 0xFF, (size_t)d.cap_points * sizeof(u32), st));
    }
    // dynamic counters live at the tail of LaneDev: re-upload the pristine host mirror
    LaneDev fresh = d;
    fresh.n_range = fresh.n_ds = fresh.n_src = fresh.n_valid = 0;
    fresh.free_top = fresh.bump = fresh.n_vox = fresh.n_tomb = fresh.map_points = 0;
    fresh.icp_arrive = 0; fresh.icp_done = 0; fresh.err = 0; fresh.icp_searches = 0;
    // tickets / release epochs keep counting on the host side; mirror them
    fresh.ticket1 = LH.tbase1; fresh.ticket2 = LH.tbase2; fresh.icp_release = LH.release_base;
    CK(cudaMemcpyAsync(ctx->d_lanes + l, &fresh, sizeof(LaneDev), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return PTK_OK;
}

extern "C" int ptk_reset(ptk_ctx* ctx, int lane) {
    if (!ctx) return PTK_E_ARG;
    if (lane >= ctx->B) return fail(ctx, PTK_E_ARG, "ptk_reset: lane out of range");
    CK(cudaSetDevice(ctx->device));
    int l0 = lane < 0 ? 0 : lane, l1 = lane < 0 ? ctx->B : lane + 1;
    for (int l = l0; l < l1; ++l) {
        LaneHost& LH = ctx->lanes[l];
        LH.poses.clear();
        LH.thr = Threshold();
        LH.last_sigma = 0.0;
        LH.have_last = false;
        LH.failed = false;
        LH.pf_src = nullptr;
        int rc = lane_reset_device(ctx, l, 0, true);
        if (rc) return rc;
    }
    return PTK_OK;
}

// stage an input array on the device if it lives on the host
static int stage_in(ptk_ctx* ctx, const double* p, size_t count, double* staging, const double** dev, cudaStream_t st) {
    if (!p) { *dev = nullptr; return PTK_OK; }
    if (is_device_ptr(p)) { *dev = p; return PTK_OK; }
    CK(cudaMemcpyAsync(staging, p, count * sizeof(double), cudaMemcpyHostToDevice, st));
    *dev = staging;
    return PTK_OK;
}

static int copy_out(ptk_ctx* ctx, void* dst, const void* dev_src, size_t bytes, cudaStream_t st) {
    if (!dst || bytes == 0) return PTK_OK;
    CK(cudaMemcpyAsync(dst, dev_src, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return PTK_OK;
}

static int err_to_code(ptk_ctx* ctx, int err) {
    if (err & ERR_POOL) return fail(ctx, PTK_E_CAPACITY, "local map voxel pool exhausted (cfg.map_capacity)");
    if (err & ERR_TABLE) return fail(ctx, PTK_E_CAPACITY, "local map hash table full");
    if (err & ERR_KEYRANGE) return fail(ctx, PTK_E_KEYRANGE, "voxel coordinate outside +-2^20");
    return PTK_OK;
}

// ---- threshold / prediction host math -----------------------------------
static bool has_moved(const ptk_config& c, const LaneHost& LH) {
    if (LH.poses.empty()) return false;
    Rigid d = rigid_mul(rigid_inv(LH.poses.front()), LH.poses.back());
    double motion = sqrt((d.t[0] * d.t[0] + d.t[1] * d.t[1]) + d.t[2] * d.t[2]);
    return motion > 5.0 * c.min_motion_th;
}

static double compute_threshold(const ptk_config& c, Threshold& t) {
    double theta = rot_angle(t.deviation.r);
    double delta_rot = 2.0 * c.max_range * sin(theta / 2.0);
    const double* tt = t.deviation.t;
    double delta_trans = sqrt((tt[0] * tt[0] + tt[1] * tt[1]) + tt[2] * tt[2]);
    double err = delta_trans + delta_rot;
    if (err > c.min_motion_th) { t.sse2 += err * err; t.num += 1; }
    if (t.num < 1) return c.initial_threshold;
    return sqrt(t.sse2 / t.num);
}

static Rigid prediction_model(const LaneHost& LH) {
    size_t n = LH.poses.size();
    if (n < 2) return rigid_identity();
    return rigid_mul(rigid_inv(LH.poses[n - 2]), LH.poses[n - 1]);
}

// ---- kernel launch helpers -----------------------------------------------------------
// `groups_hint`: upper estimate of the 32-point source groups per lane (0 = unknown); blocks beyond
// one per group would only add arrivals to the per-iteration barrier.
static int launch_icp(ptk_ctx* ctx, int l0, int cnt, int groups_hint, cudaStream_t st) {
    {   // iteration 0's searches: every source point of every lane, one thread each (k_icp_search0)
        const int max_groups = (ctx->cfg.max_points + 31) / 32;
        const int groups = groups_hint > 0 ? std::min(groups_hint, max_groups) : max_groups;
        const int wpb = S0_THREADS / 32;
        int gx = std::max(1, std::min((groups + wpb - 1) / wpb, std::max(1, (ctx->num_sms * 16) / cnt)));
        LAUNCH(PS_SEARCH0, st, k_icp_search0<<<dim3(gx, cnt), S0_THREADS, 0, st>>>(ctx->d_lanes + l0, ctx->d_params + l0));
        CK(cudaGetLastError());
    }
    // Wide batches: one THREAD-BLOCK CLUSTER per lane.  The only thing the lane's blocks need from each other
    // is to be running at the same time (their per-iteration barrier spins on a counter), which is exactly
    // what a cluster guarantees - so the launch needs no grid-wide co-residency, the grid may hold more
    // clusters than fit at once, and the lanes that converge early (iterations per scan vary 2-4x between
    // lanes) hand their SMs to queued lanes instead of leaving them idle until the slowest lane ends.
    // (measured: pays off from ~56 lanes on - 64 lanes 19.7k -> 21.1k scans/s; at 48 the cooperative launch is faster)
    if (ctx->icp_cluster > 1 && cnt >= ctx->icp_cluster_min_lanes) {
        int cl = ctx->icp_cluster;
        if (groups_hint > 0) while (cl > 1 && cl / 2 >= groups_hint) cl /= 2;
        LaneDev* dl = ctx->d_lanes + l0;
        const StepParams* dp = ctx->d_params + l0;
        StepOut* dout = ctx->d_outs + l0;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(cl, cnt);
        cfg.blockDim = dim3(ICP_THREADS);
        cfg.dynamicSmemBytes = ICP_SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = cudaSuccess;
        LAUNCH(PS_ICP, st, le = cudaLaunchKernelEx(&cfg, k_icp, dl, dp, dout));
        CK(le);
        return PTK_OK;
    }
    // Few lanes: a cooperative launch with as many blocks per lane as the device holds
    // (all blocks of a cooperative launch must be co-resident: split wide batches)
    int done = 0;
    while (done < cnt) {
        int chunk = std::min(cnt - done, ctx->icp_blocks_total);
        int per = std::max(1, ctx->icp_blocks_total / chunk);
        if (groups_hint > 0) per = std::max(1, std::min(per, groups_hint));
        per = std::min(per, ctx->icp_max_blocks_per_lane);
        if (ctx->xch_local) per = std::min(per, XCH_FLAGS - 2);      // one exchange stamp per block
        LaneDev* dl = ctx->d_lanes + l0 + done;
        StepParams* dp = ctx->d_params + l0 + done;
        StepOut* dout = ctx->d_outs + l0 + done;
        void* args[] = {&dl, &dp, &dout};
        cudaError_t le = cudaSuccess;
        LAUNCH(PS_ICP, st, le = cudaLaunchCooperativeKernel((void*)k_icp, dim3(per, chunk), dim3(ICP_THREADS), args, ICP_SMEM, st));
        CK(le);
        done += chunk;
    }
    return PTK_OK;
}

static int map_update_launch(ptk_ctx* ctx, int l0, int cnt, int nmax, int use_pose, const double* origin_dev,
                             bool do_insert, bool do_prune, cudaStream_t st) {
    int gx = std::max(1, std::min((nmax + 255) / 256, std::max(1, (ctx->num_sms * 4) / cnt)));
    if (do_insert) {
        LAUNCH(PS_MAP_INSERT, st, k_map_insert<<<dim3(gx, cnt), 256, 0, st>>>(ctx->d_lanes + l0, ctx->d_params + l0, ctx->d_outs + l0, use_pose));
        LAUNCH(PS_MAP_COMMIT, st, k_map_commit<<<dim3(gx, cnt), 256, 0, st>>>(ctx->d_lanes + l0, ctx->d_outs + l0, use_pose));
    }
    if (do_prune) {
        int gp = std::max(1, (ctx->num_sms * 4) / cnt);
        LAUNCH(PS_MAP_PRUNE, st, k_map_prune<<<dim3(gp, cnt), 256, 0, st>>>(ctx->d_lanes + l0, ctx->d_outs + l0, origin_dev));
    }
    CK(cudaGetLastError());
    return PTK_OK;
}

static int maybe_rebuild(ptk_ctx* ctx, int l, const StepOut& O, cudaStream_t st) {
    LaneDev& d = ctx->lanes[l].d;
    size_t cap = (size_t)d.m_mask + 1;
    if ((size_t)(O.n_vox + O.n_tomb) * 2 > cap && O.n_tomb > 0) {
        CK(cudaMemsetAsync(d.m_slots, 0xFF, cap * sizeof(MapSlot), st));
        LAUNCH(PS_REBUILD, st, k_map_rebuild<<<dim3(ctx->num_sms, 1), 256, 0, st>>>(ctx->d_lanes + l));
        CK(cudaGetLastError());
    }
    return PTK_OK;
}

// First half of the step for lanes [l0, l0+cnt): parameters, upload, deskew + range filter + the two
// voxel grids (kiss.py:90-105).  Leaves frame_downsample / source on the device.
static int step_prepare(ptk_ctx* ctx, int l0, int cnt, const double* const* xyz, const double* const* ts, const int* n_in,
                        const unsigned int* const* range, const double* guesses, const unsigned char* has_guess,
                        int* nmax_out, cudaStream_t st) {
    const ptk_config& c = ctx->cfg;
    CK(cudaSetDevice(ctx->device));
    int nmax = 0;
    std::vector<int> nv(cnt);
    const int* n = nv.data();
    const int npix = ctx->sen_H * ctx->sen_W;
    if (range && npix <= 0) return fail(ctx, PTK_E_STATE, "ptk_set_sensor has not been called");
    for (int k = 0; k < cnt; ++k) {
        int l = l0 + k;
        LaneHost& LH = ctx->lanes[l];
        nv[k] = range ? npix : n_in[k];
        if (LH.failed) return fail(ctx, PTK_E_STATE, "lane hit a capacity error: its local map is incomplete, ptk_reset it first");
        if (n[k] < 0 || n[k] > c.max_points) return fail(ctx, PTK_E_CAPACITY, "scan larger than cfg.max_points");
        if (!range && n[k] > 0 && !xyz[k]) return fail(ctx, PTK_E_ARG, "xyz is null");
        if (range && !range[k]) return fail(ctx, PTK_E_ARG, "range image is null");
        StepParams& P = ctx->h_params[l];
        memset(&P, 0, sizeof(P));
        int rc = PTK_OK;
        if (range) {
            if (LH.pf_src == (const void*)range[k]) {     // prefetched: the copy overlapped the previous step
                CK(cudaStreamWaitEvent(st, ctx->pf_event, 0));
                P.range = LH.pf_range[LH.pf_buf];
                LH.cur_buf = LH.pf_buf;
                LH.pf_src = nullptr;
            } else if (is_device_ptr(range[k])) { P.range = range[k]; LH.pf_src = nullptr; }
            else {
                LH.pf_src = nullptr;     // any other image invalidates an earlier prefetch (its host buffer may be refilled)
                CK(cudaMemcpyAsync(LH.pf_range[LH.cur_buf], range[k], (size_t)npix * sizeof(u32), cudaMemcpyHostToDevice, st));
                P.range = LH.pf_range[LH.cur_buf];
            }
            P.lut_dir = ctx->d_lut_dir; P.lut_off = ctx->d_lut_off; P.col_ts = ctx->d_col_ts;
            P.col_motion = LH.col_motion; P.W = ctx->sen_W; P.range_unit = ctx->sen_unit;
        } else {
            rc = stage_in(ctx, xyz[k], (size_t)n[k] * 3, LH.in_xyz, &P.xyz, st);
            if (rc) return rc;
        }
        P.n = n[k];
        P.flags = F_RANGE | F_SECOND;
        size_t np = LH.poses.size();
        if (c.deskew && np >= 2) {      // MotionCompensator.deskew_scan: identity with < 2 poses
            if (!range) {
                if (!ts || !ts[k]) return fail(ctx, PTK_E_ARG, "timestamps are null");
                rc = stage_in(ctx, ts[k], (size_t)n[k], LH.in_ts, &P.ts, st);
                if (rc) return rc;
            }
            P.flags |= F_DESKEW;
            Rigid rel = rigid_mul(rigid_inv(LH.poses[np - 2]), LH.poses[np - 1]);
            se3_log(rel, P.delta);
        }
        P.ds1_size = c.voxel_size * 0.5;     // KissICP.voxelize
        P.ds2_size = c.voxel_size * 1.5;
        P.ds1_inv = 1.0 / P.ds1_size; P.ds2_inv = 1.0 / P.ds2_size;
        P.max_range = c.max_range;
        P.min_range = c.min_range;
        double sigma = has_moved(c, LH) ? compute_threshold(c, LH.thr) : c.initial_threshold;   // kiss.py:99
        LH.pending_sigma = sigma;
        if (guesses && (!has_guess || has_guess[k])) {
            rigid_from_mat16(guesses + 16 * (size_t)k, P.guess);
        } else {                                                                                // kiss.py:102-105
            Rigid last = np ? LH.poses.back() : rigid_identity();
            P.guess = rigid_mul(last, prediction_model(LH));
        }
        if (LH.have_last && LH.last_out.n_ds > 0) {
            // the scan tables' front regions: four slots per key the previous scan put there (at least 4096)
            u32 c1 = next_pow2((u32)std::max(4096, 4 * LH.last_out.n_ds)), c2 = next_pow2((u32)std::max(4096, 4 * LH.last_out.n_src));
            P.near1 = c1 <= LH.d.t_mask / 2 ? c1 - 1 : 0;
            P.near2 = c2 <= LH.d.t_mask / 2 ? c2 - 1 : 0;
        }
        P.max_corr = 3 * sigma;                                                                 // kiss.py:112-113
        P.kernel = sigma / 3;
        P.epoch = ++LH.epoch;
        P.xch_epoch = P.epoch;
        nmax = std::max(nmax, n[k]);
    }
    // wide batches: fewer, longer-lived blocks (block dispatch limits the rate); a few lanes: all the blocks we can get
    const int si_tiles = cnt >= 8 ? SI_TILES : 1;
    int g1 = std::max(1, (nmax + 256 * si_tiles - 1) / (256 * si_tiles));
    int gt = std::max(1, ((nmax + TILE - 1) / TILE + CT_TILES - 1) / CT_TILES);     // blocks; each takes CT_TILES tickets
    for (int k = 0; k < cnt; ++k) {
        LaneHost& LH = ctx->lanes[l0 + k];
        StepParams& P = ctx->h_params[l0 + k];
        P.tbase1 = LH.tbase1; P.tbase2 = LH.tbase2; P.release_base = LH.release_base;
        LH.tbase1 += (u32)(gt * CT_TILES); LH.tbase2 += (u32)(gt * CT_TILES); LH.release_base += (u32)c.max_iterations + 1u;
        LH.last_params = P;
    }
    CK(cudaMemcpyAsync(ctx->d_params + l0, ctx->h_params + l0, sizeof(StepParams) * cnt, cudaMemcpyHostToDevice, st));
    LaneDev* dl = ctx->d_lanes + l0;
    StepParams* dp = ctx->d_params + l0;
    if (range) {
        bool any = false;
        for (int k = 0; k < cnt; ++k) any = any || (ctx->h_params[l0 + k].flags & F_DESKEW);
        if (any) LAUNCH(PS_COL_MOTION, st, k_col_motion<<<dim3((ctx->sen_W + 127) / 128, cnt), 128, 0, st>>>(dp));
    }
    if (range) {
        const int ncb = (ctx->sen_W + 255) / 256, nrb = (ctx->sen_H + SI_ROWS - 1) / SI_ROWS;
        LAUNCH(PS_SCAN_INSERT, st, k_scan_insert_range<<<dim3(ncb * nrb, cnt), 256, 0, st>>>(dl, dp, ctx->sen_H, ncb));
    } else {
        LAUNCH(PS_SCAN_INSERT, st, k_scan_insert<<<dim3(g1, cnt), 256, 0, st>>>(dl, dp, si_tiles));
    }
    LAUNCH(PS_COMPACT1, st, k_compact1<<<dim3(gt, cnt), 256, 0, st>>>(dl, dp));
    LAUNCH(PS_COMPACT2, st, k_compact2<<<dim3(gt, cnt), 256, 0, st>>>(dl, dp));
    CK(cudaGetLastError());
    *nmax_out = nmax;
    return PTK_OK;
}

static int step_finish(ptk_ctx* ctx, int l0, int cnt, int nmax, double* out_poses, ptk_stats* stats, cudaStream_t st);
static int issue_prefetch(ptk_ctx* ctx);

// The whole step for lanes [l0, l0+cnt).
// `range` non-null selects the range-image input (one H*W uint32 image per lane; xyz/ts/n unused).
static int run_step(ptk_ctx* ctx, int l0, int cnt, const double* const* xyz, const double* const* ts, const int* n_in,
                    const unsigned int* const* range, const double* guesses, const unsigned char* has_guess,
                    double* out_poses, ptk_stats* stats, cudaStream_t st) {
    int nmax = 0;
    int rc = step_prepare(ctx, l0, cnt, xyz, ts, n_in, range, guesses, has_guess, &nmax, st);
    if (rc) return rc;
    // source size changes slowly from scan to scan: size the ICP grid from the last one
    int groups_hint = 0;
    for (int k = 0; k < cnt; ++k) {
        const LaneHost& LH = ctx->lanes[l0 + k];
        if (!LH.have_last || LH.last_out.n_src <= 0) { groups_hint = 0; break; }
        groups_hint = std::max(groups_hint, (int)(((long long)LH.last_out.n_src * 5 / 4 + 31) / 32) + 1);
    }
    rc = launch_icp(ctx, l0, cnt, groups_hint, st);
    if (rc) return rc;
    return step_finish(ctx, l0, cnt, nmax, out_poses, stats, st);
}

// Second half: local-map update with the new pose (kiss.py:129), counters, and the host bookkeeping of
// kiss.py:116-130 (pose gain metrics, model deviation, pose list).
static int step_finish(ptk_ctx* ctx, int l0, int cnt, int nmax, double* out_poses, ptk_stats* stats, cudaStream_t st) {
    LaneDev* dl = ctx->d_lanes + l0;
    StepOut* dout = ctx->d_outs + l0;
    int rc = map_update_launch(ctx, l0, cnt, nmax, 1, nullptr, true, true, st);
    if (rc) return rc;
    LAUNCH(PS_FINISH, st, k_finish<<<(cnt + 63) / 64, 64, 0, st>>>(dl, dout, cnt));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_outs + l0, dout, sizeof(StepOut) * cnt, cudaMemcpyDeviceToHost, st));
    {   // everything of this step is queued: start moving the next scans while it runs
        int prc = issue_prefetch(ctx);
        if (prc) return prc;
    }
    if (ctx->blocking_sync) {
        // More worker threads than host cores.  PTK_FLEET_WAIT=yield (default): poll the event and give the core away
        // between polls - a thread with launches to make gets it at once, nobody sleeps through its step's end;
        // PTK_FLEET_WAIT=block: sleep on a blocking-sync event (no CPU while waiting, a wake-up latency per step).
        static const bool yield_wait = !(getenv("PTK_FLEET_WAIT") && !strcmp(getenv("PTK_FLEET_WAIT"), "block"));
        if (!ctx->sync_event)
            CK(cudaEventCreateWithFlags(&ctx->sync_event, (yield_wait ? 0 : cudaEventBlockingSync) | cudaEventDisableTiming));
        CK(cudaEventRecord(ctx->sync_event, st));
        if (yield_wait) {
            cudaError_t q;
            while ((q = cudaEventQuery(ctx->sync_event)) == cudaErrorNotReady) sched_yield();
            CK(q);
        } else {
            CK(cudaEventSynchronize(ctx->sync_event));
        }
    } else {
        CK(cudaStreamSynchronize(st));
    }
    prof_collect(ctx);
    int ret = PTK_OK;
    for (int k = 0; k < cnt; ++k) {
        int l = l0 + k;
        LaneHost& LH = ctx->lanes[l];
        const StepOut& O = ctx->h_outs[l];
        const StepParams& P = ctx->h_params[l];
        LH.last_out = O;
        LH.have_last = true;
        LH.last_reg_iters = O.iterations;
        LH.last_reg_nsrc = O.n_src;
        int ec = err_to_code(ctx, O.err);
        if (ec) {
            // points of this scan are missing from the local map (or a key overflowed): the pose is not committed
            // (the Python wrapper raises before it records anything, kiss.py:54-74 lets exceptions propagate) and
            // the lane stays unusable until it is reset
            if (!ret) ret = ec;
            if (ec == PTK_E_CAPACITY) LH.failed = true;
            continue;
        }
        Rigid gain = rigid_mul(rigid_inv(P.guess), O.pose);                   // kiss.py:116
        double dt = sqrt((gain.t[0] * gain.t[0] + gain.t[1] * gain.t[1]) + gain.t[2] * gain.t[2]);
        double om[3], theta;
        so3_log(gain.r, om, theta);
        LH.thr.deviation = gain;                                              // kiss.py:128
        LH.poses.push_back(O.pose);                                           // kiss.py:130
        LH.last_sigma = LH.pending_sigma;
        if (out_poses) rigid_to_mat16(O.pose, out_poses + 16 * (size_t)k);
        if (stats) {
            ptk_stats& S = stats[k];
            memset(&S, 0, sizeof(S));
            S.status = O.status; S.n_in = P.range ? O.n_valid : P.n; S.n_range = O.n_range; S.n_ds = O.n_ds; S.n_src = O.n_src;
            S.n_voxels = O.n_vox; S.iterations = O.iterations; S.n_corr = O.n_corr; S.dx_norm = O.dx_norm;
            S.sigma = LH.pending_sigma; S.err_dt = dt; S.err_drot = fabs(theta); S.map_points = O.map_points;
            S.icp_searches = O.icp_searches;
        }
        if (O.status == 2 && !ret) ret = fail(ctx, PTK_E_NUMERIC, "singular normal equations in ICP");
        int rb = maybe_rebuild(ctx, l, O, st);
        if (rb && !ret) ret = rb;
    }
    return ret;
}

extern "C" int ptk_register_frame(ptk_ctx* ctx, int lane, const double* xyz, const double* timestamps, int n,
                                  const double* initial_guess, double* out_pose, ptk_stats* stats, void* stream) {
    if (!ctx) return PTK_E_ARG;
    if (lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "lane out of range");
    unsigned char hg = initial_guess ? 1 : 0;
    return run_step(ctx, lane, 1, &xyz, &timestamps, &n, nullptr, initial_guess, &hg, out_pose, stats, (cudaStream_t)stream);
}

extern "C" int ptk_register_frame_batch(ptk_ctx* ctx, const double* const* xyz, const double* const* timestamps,
                                        const int* n, const double* guesses, const unsigned char* has_guess,
                                        double* out_poses, ptk_stats* stats, void* stream) {
    if (!ctx || !xyz || !n) return PTK_E_ARG;
    return run_step(ctx, 0, ctx->B, xyz, timestamps, n, nullptr, guesses, has_guess, out_poses, stats, (cudaStream_t)stream);
}

// ---- range-image entry points (kiss.py:54-74 with the projection on the device) ---------
extern "C" int ptk_set_sensor(ptk_ctx* ctx, int H, int W, const double* direction, const double* offset,
                              const double* col_timestamps, double range_unit) {
    if (!ctx || H < 1 || W < 1 || !direction || !(range_unit > 0.0)) return fail(ctx, PTK_E_ARG, "ptk_set_sensor: bad argument");
    if ((long long)H * W > ctx->cfg.max_points) return fail(ctx, PTK_E_CAPACITY, "ptk_set_sensor: H*W exceeds cfg.max_points");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    for (void* p : ctx->sensor_allocs) cudaFree(p);
    ctx->sensor_allocs.clear();
    ctx->d_lut_dir = ctx->d_lut_off = ctx->d_col_ts = nullptr;
    ctx->sen_H = ctx->sen_W = 0;
    size_t np = (size_t)H * W;
    auto up = [&](double** dst, const double* src, size_t count) -> int {
        CK(dalloc(ctx->sensor_allocs, dst, count));
        CK(cudaMemcpy(*dst, src, count * sizeof(double), is_device_ptr(src) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
        return PTK_OK;
    };
    // (H*W,3) tables of the caller -> three planes on the device (coalesced loads in the scan kernels)
    auto up_planes = [&](double** dst, const double* src) -> int {
        std::vector<double> aos(np * 3), soa(np * 3);
        if (is_device_ptr(src)) CK(cudaMemcpy(aos.data(), src, np * 3 * sizeof(double), cudaMemcpyDeviceToHost));
        else memcpy(aos.data(), src, np * 3 * sizeof(double));
        for (size_t i = 0; i < np; ++i)
            for (int c = 0; c < 3; ++c) soa[(size_t)c * np + i] = aos[3 * i + c];
        CK(dalloc(ctx->sensor_allocs, dst, np * 3));
        CK(cudaMemcpy(*dst, soa.data(), np * 3 * sizeof(double), cudaMemcpyHostToDevice));
        return PTK_OK;
    };
    int rc = up_planes(&ctx->d_lut_dir, direction);
    if (rc) return rc;
    if (offset && (rc = up_planes(&ctx->d_lut_off, offset))) return rc;
    std::vector<double> cts;
    if (!col_timestamps) {      // np.linspace(0, 1.0, w, endpoint=False) (kiss.py:34)
        cts.resize(W);
        const double step = 1.0 / (double)W;
        for (int w = 0; w < W; ++w) cts[w] = (double)w * step;
        col_timestamps = cts.data();
    }
    if ((rc = up(&ctx->d_col_ts, col_timestamps, (size_t)W))) return rc;
    for (auto& LH : ctx->lanes) CK(dalloc(ctx->sensor_allocs, &LH.col_motion, (size_t)12 * W));
    ctx->sen_H = H; ctx->sen_W = W; ctx->sen_unit = range_unit;
    return PTK_OK;
}

// Host-to-device copy of the NEXT step's range images on a side stream, overlapped with the current
// step: ptk_prefetch_scan_batch only notes the host pointers; the next ptk_register_scan[_batch] issues
// the copies right after it has launched its own kernels (so they run while the GPU computes) and the
// call after that recognises the pointers and merely waits for the copy event.
static int issue_prefetch(ptk_ctx* ctx) {
    if (ctx->pf_pending.empty()) return PTK_OK;
    const int npix = ctx->sen_H * ctx->sen_W;
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->pf_event, cudaEventDisableTiming));
    }
    for (int l = 0; l < ctx->B && l < (int)ctx->pf_pending.size(); ++l) {
        LaneHost& LH = ctx->lanes[l];
        const unsigned int* src = ctx->pf_pending[l];
        LH.pf_src = nullptr;
        if (!src) continue;
        const int buf = 1 - LH.cur_buf;          // the buffer the step in flight does not read
        CK(cudaMemcpyAsync(LH.pf_range[buf], src, (size_t)npix * sizeof(u32), cudaMemcpyHostToDevice, ctx->copy_stream));
        LH.pf_src = src;
        LH.pf_buf = buf;
    }
    ctx->pf_pending.clear();
    CK(cudaEventRecord(ctx->pf_event, ctx->copy_stream));
    return PTK_OK;
}

extern "C" int ptk_prefetch_scan_batch(ptk_ctx* ctx, const unsigned int* const* range_mm) {
    if (!ctx || !range_mm) return PTK_E_ARG;
    if (ctx->sen_H * ctx->sen_W <= 0) return fail(ctx, PTK_E_STATE, "ptk_set_sensor has not been called");
    ctx->pf_pending.assign(range_mm, range_mm + ctx->B);
    for (auto& p : ctx->pf_pending)
        if (p && is_device_ptr(p)) p = nullptr;
    return PTK_OK;
}

extern "C" int ptk_register_scan(ptk_ctx* ctx, int lane, const unsigned int* range_mm, const double* initial_guess,
                                 double* out_pose, ptk_stats* stats, void* stream) {
    if (!ctx) return PTK_E_ARG;
    if (lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "lane out of range");
    unsigned char hg = initial_guess ? 1 : 0;
    return run_step(ctx, lane, 1, nullptr, nullptr, nullptr, &range_mm, initial_guess, &hg, out_pose, stats, (cudaStream_t)stream);
}

extern "C" int ptk_register_scan_batch(ptk_ctx* ctx, const unsigned int* const* range_mm, const double* guesses,
                                       const unsigned char* has_guess, double* out_poses, ptk_stats* stats, void* stream) {
    if (!ctx || !range_mm) return PTK_E_ARG;
    return run_step(ctx, 0, ctx->B, nullptr, nullptr, nullptr, range_mm, guesses, has_guess, out_poses, stats, (cudaStream_t)stream);
}

// ---- state accessors -----------------------------------------------------------------
extern "C" int ptk_num_poses(const ptk_ctx* ctx, int lane) {
    if (!ctx || lane < 0 || lane >= ctx->B) return PTK_E_ARG;
    return (int)ctx->lanes[lane].poses.size();
}

extern "C" int ptk_get_pose(const ptk_ctx* ctx, int lane, int index, double* out16) {
    if (!ctx || lane < 0 || lane >= ctx->B || !out16) return PTK_E_ARG;
    const auto& p = ctx->lanes[lane].poses;
    int n = (int)p.size();
    if (index < 0) index += n;
    if (index < 0 || index >= n) return PTK_E_ARG;
    rigid_to_mat16(p[index], out16);
    return PTK_OK;
}

extern "C" int ptk_get_prediction_model(const ptk_ctx* ctx, int lane, double* out16) {
    if (!ctx || lane < 0 || lane >= ctx->B || !out16) return PTK_E_ARG;
    rigid_to_mat16(prediction_model(ctx->lanes[lane]), out16);
    return PTK_OK;
}

extern "C" double ptk_last_sigma(const ptk_ctx* ctx, int lane) {
    if (!ctx || lane < 0 || lane >= ctx->B) return 0.0;
    return ctx->lanes[lane].last_sigma;
}

extern "C" int ptk_get_adaptive_threshold(ptk_ctx* ctx, int lane, double* sigma) {
    if (!ctx || lane < 0 || lane >= ctx->B || !sigma) return PTK_E_ARG;
    LaneHost& LH = ctx->lanes[lane];
    *sigma = has_moved(ctx->cfg, LH) ? compute_threshold(ctx->cfg, LH.thr) : ctx->cfg.initial_threshold;
    LH.last_sigma = *sigma;
    return PTK_OK;
}

extern "C" int ptk_update_model_deviation(ptk_ctx* ctx, int lane, const double* T16) {
    if (!ctx || lane < 0 || lane >= ctx->B || !T16) return PTK_E_ARG;
    rigid_from_mat16(T16, ctx->lanes[lane].thr.deviation);
    return PTK_OK;
}

extern "C" int ptk_append_pose(ptk_ctx* ctx, int lane, const double* T16) {
    if (!ctx || lane < 0 || lane >= ctx->B || !T16) return PTK_E_ARG;
    Rigid T;
    rigid_from_mat16(T16, T);
    ctx->lanes[lane].poses.push_back(T);
    return PTK_OK;
}

// ---- stand-alone pieces (scratch lane = index B) -------------------------------------
static int scratch_params(ptk_ctx* ctx, StepParams& P, cudaStream_t st, bool k2, bool k3) {
    int S = ctx->B;
    LaneHost& LH = ctx->lanes[S];
    P.epoch = ++LH.epoch;
    int gt = std::max(1, ((P.n + TILE - 1) / TILE + CT_TILES - 1) / CT_TILES);
    P.tbase1 = LH.tbase1; P.tbase2 = LH.tbase2;
    if (k2) LH.tbase1 += (u32)(gt * CT_TILES);
    if (k3) LH.tbase2 += (u32)(gt * CT_TILES);
    ctx->h_params[S] = P;
    CK(cudaMemcpyAsync(ctx->d_params + S, ctx->h_params + S, sizeof(StepParams), cudaMemcpyHostToDevice, st));
    return PTK_OK;
}

extern "C" int ptk_deskew_scan(ptk_ctx* ctx, const double* xyz, const double* timestamps, int n,
                               const double* start_pose, const double* finish_pose, double* out_xyz, void* stream) {
    if (!ctx || !out_xyz || n < 0 || (n > 0 && (!xyz || !timestamps)) || !start_pose || !finish_pose) return fail(ctx, PTK_E_ARG, "ptk_deskew_scan: bad argument");
    if (n > ctx->cfg.max_points) return fail(ctx, PTK_E_CAPACITY, "scan larger than cfg.max_points");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    LaneHost& LH = ctx->lanes[ctx->B];
    StepParams P;
    memset(&P, 0, sizeof(P));
    P.flags = F_DESKEW;
    Rigid a, b;
    rigid_from_mat16(start_pose, a);
    rigid_from_mat16(finish_pose, b);
    se3_log(rigid_mul(rigid_inv(a), b), P.delta);
    const double *dx, *dt;
    int rc = stage_in(ctx, xyz, (size_t)n * 3, LH.in_xyz, &dx, st);
    if (rc) return rc;
    rc = stage_in(ctx, timestamps, (size_t)n, LH.in_ts, &dt, st);
    if (rc) return rc;
    if (n > 0) {
        double* dout = is_device_ptr(out_xyz) ? out_xyz : ctx->d_tmp;
        LAUNCH(PS_OTHER, st, k_deskew<<<(n + 255) / 256, 256, 0, st>>>(dx, dt, n, P, dout));
        CK(cudaGetLastError());
        if (dout != out_xyz) CK(cudaMemcpyAsync(out_xyz, dout, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    return PTK_OK;
}

// shared by preprocess / voxel_down_sample / get_frame: select on the scratch lane, copy out
static int scratch_select(ptk_ctx* ctx, StepParams P, bool voxel, double* out_xyz, int* out_index, int capacity,
                          int* n_out, cudaStream_t st) {
    int S = ctx->B;
    LaneHost& LH = ctx->lanes[S];
    int rc = scratch_params(ctx, P, st, true, false);
    if (rc) return rc;
    LaneDev* dl = ctx->d_lanes + S;
    StepParams* dp = ctx->d_params + S;
    int g1 = std::max(1, (P.n + 255) / 256), gt = std::max(1, ((P.n + TILE - 1) / TILE + CT_TILES - 1) / CT_TILES);
    if (voxel) LAUNCH(PS_OTHER, st, k_scan_insert<<<dim3(g1, 1), 256, 0, st>>>(dl, dp, 1));
    LAUNCH(PS_OTHER, st, k_compact1<<<dim3(gt, 1), 256, 0, st>>>(dl, dp));
    if (voxel) LAUNCH(PS_OTHER, st, k_clean_tables<<<dim3(ctx->num_sms, 1), 256, 0, st>>>(dl, 1));
    LAUNCH(PS_OTHER, st, k_finish<<<1, 64, 0, st>>>(dl, ctx->d_outs + S, 1));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_outs + S, ctx->d_outs + S, sizeof(StepOut), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const StepOut& O = ctx->h_outs[S];
    int ec = err_to_code(ctx, O.err);
    if (ec) return ec;
    int m = O.n_ds;
    if (n_out) *n_out = m;
    if (capacity >= 0 && m > capacity) return fail(ctx, PTK_E_CAPACITY, "output buffer too small");
    if (m > 0 && out_xyz) {
        double* dout = is_device_ptr(out_xyz) ? out_xyz : ctx->d_tmp;
        LAUNCH(PS_OTHER, st, k_gather_aos<<<(m + 255) / 256, 256, 0, st>>>(LH.d.ds_x, LH.d.ds_y, LH.d.ds_z, m, dout));
        CK(cudaGetLastError());
        if (dout != out_xyz) CK(cudaMemcpyAsync(out_xyz, dout, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (m > 0 && out_index) {
        rc = copy_out(ctx, out_index, LH.d.ds_idx, (size_t)m * sizeof(int), st);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(st));
    return PTK_OK;
}

extern "C" int ptk_preprocess(ptk_ctx* ctx, const double* xyz, int n, double max_range, double min_range,
                              double* out_xyz, int* n_out, void* stream) {
    if (!ctx || n < 0 || (n > 0 && !xyz)) return fail(ctx, PTK_E_ARG, "ptk_preprocess: bad argument");
    if (n > ctx->cfg.max_points) return fail(ctx, PTK_E_CAPACITY, "scan larger than cfg.max_points");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    StepParams P;
    memset(&P, 0, sizeof(P));
    P.n = n;
    P.flags = F_RANGE | F_SELECT_RANGE;
    P.max_range = max_range; P.min_range = min_range;
    int rc = stage_in(ctx, xyz, (size_t)n * 3, ctx->lanes[ctx->B].in_xyz, &P.xyz, st);
    if (rc) return rc;
    return scratch_select(ctx, P, false, out_xyz, nullptr, -1, n_out, st);
}

extern "C" int ptk_voxel_down_sample(ptk_ctx* ctx, const double* xyz, int n, double voxel_size, double* out_xyz,
                                     int* out_index, int* n_out, void* stream) {
    if (!ctx || n < 0 || (n > 0 && !xyz) || !(voxel_size > 0.0)) return fail(ctx, PTK_E_ARG, "ptk_voxel_down_sample: bad argument");
    if (n > ctx->cfg.max_points) return fail(ctx, PTK_E_CAPACITY, "scan larger than cfg.max_points");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    StepParams P;
    memset(&P, 0, sizeof(P));
    P.n = n;
    P.flags = 0;
    P.ds1_size = voxel_size;
    P.ds1_inv = 1.0 / voxel_size;
    int rc = stage_in(ctx, xyz, (size_t)n * 3, ctx->lanes[ctx->B].in_xyz, &P.xyz, st);
    if (rc) return rc;
    return scratch_select(ctx, P, true, out_xyz, out_index, -1, n_out, st);
}

extern "C" int ptk_get_frame(ptk_ctx* ctx, int lane, double* out_xyz, int capacity, int* n_out, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "ptk_get_frame: bad argument");
    LaneHost& LH = ctx->lanes[lane];
    if (!LH.have_last) return fail(ctx, PTK_E_STATE, "ptk_get_frame: no step has run");
    CK(cudaSetDevice(ctx->device));
    StepParams P = LH.last_params;   // the inputs of the last step must still be resident
    P.flags = (P.flags & F_DESKEW) | F_RANGE | F_SELECT_RANGE;
    return scratch_select(ctx, P, false, out_xyz, nullptr, capacity, n_out, (cudaStream_t)stream);
}

extern "C" int ptk_get_points(ptk_ctx* ctx, int lane, int which, double* out_xyz, int* out_index, int capacity,
                              int* n_out, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || which < 0 || which > 1) return fail(ctx, PTK_E_ARG, "ptk_get_points: bad argument");
    LaneHost& LH = ctx->lanes[lane];
    if (!LH.have_last) return fail(ctx, PTK_E_STATE, "ptk_get_points: no step has run");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    int m = which == 0 ? LH.last_out.n_ds : LH.last_out.n_src;
    if (n_out) *n_out = m;
    if (m > capacity) return fail(ctx, PTK_E_CAPACITY, "output buffer too small");
    const LaneDev& d = LH.d;
    if (m > 0 && out_xyz) {
        double* dout = is_device_ptr(out_xyz) ? out_xyz : ctx->d_tmp;
        if (which == 0) LAUNCH(PS_OTHER, st, k_gather_aos<<<(m + 255) / 256, 256, 0, st>>>(d.ds_x, d.ds_y, d.ds_z, m, dout));
        else LAUNCH(PS_OTHER, st, k_gather_aos<<<(m + 255) / 256, 256, 0, st>>>(d.s0_x, d.s0_y, d.s0_z, m, dout));
        CK(cudaGetLastError());
        if (dout != out_xyz) CK(cudaMemcpyAsync(out_xyz, dout, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (m > 0 && out_index) {
        int rc = copy_out(ctx, out_index, which == 0 ? d.ds_idx : d.s_idx, (size_t)m * sizeof(int), st);
        if (rc) return rc;
    }
    CK(cudaStreamSynchronize(st));
    return PTK_OK;
}

extern "C" int ptk_get_trace(ptk_ctx* ctx, int lane, int* out_order, int capacity_iters, int* n_iters, int* n_src,
                             void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "ptk_get_trace: bad argument");
    LaneHost& LH = ctx->lanes[lane];
    if (!LH.have_last) return fail(ctx, PTK_E_STATE, "ptk_get_trace: no registration has run");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    int iters = std::min(LH.last_reg_iters, LH.d.trace_iters);
    int m = LH.last_reg_nsrc;
    if (n_iters) *n_iters = iters;
    if (n_src) *n_src = m;
    if (out_order && m > 0) {
        iters = std::min(iters, capacity_iters);
        for (int it = 0; it < iters; ++it) {
            int rc = copy_out(ctx, out_order + (size_t)it * m, LH.d.trace + (size_t)it * LH.d.cap_points, (size_t)m * sizeof(int), st);
            if (rc) return rc;
        }
    }
    CK(cudaStreamSynchronize(st));
    return PTK_OK;
}

// ---- map taps ------------------------------------------------------------------------
static int lane_counters(ptk_ctx* ctx, int lane, StepOut* O, cudaStream_t st) {
    LAUNCH(PS_OTHER, st, k_finish<<<1, 64, 0, st>>>(ctx->d_lanes + lane, ctx->d_outs + lane, 1));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_outs + lane, ctx->d_outs + lane, sizeof(StepOut), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // k_finish overwrote only the counter fields; keep the pose of the last step
    StepOut& H = ctx->h_outs[lane];
    *O = H;
    return PTK_OK;
}

extern "C" int ptk_map_clear(ptk_ctx* ctx, int lane, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "ptk_map_clear: bad argument");
    CK(cudaSetDevice(ctx->device));
    return lane_reset_device(ctx, lane, (cudaStream_t)stream, false);
}

extern "C" int ptk_map_num_points(ptk_ctx* ctx, int lane, int* n_points, int* n_voxels) {
    if (!ctx || lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "ptk_map_num_points: bad argument");
    CK(cudaSetDevice(ctx->device));
    StepOut O;
    int rc = lane_counters(ctx, lane, &O, 0);
    if (rc) return rc;
    if (n_points) *n_points = O.map_points;
    if (n_voxels) *n_voxels = O.n_vox;
    return PTK_OK;
}

extern "C" int ptk_map_empty(ptk_ctx* ctx, int lane) {
    int nv = 0;
    int rc = ptk_map_num_points(ctx, lane, nullptr, &nv);
    if (rc) return rc;
    return nv == 0 ? 1 : 0;
}

static int map_mutate(ptk_ctx* ctx, int lane, const double* xyz, int n, const double* pose, const double* origin3,
                      bool insert, bool prune, cudaStream_t st) {
    if (!ctx || lane < 0 || lane >= ctx->B || n < 0 || (insert && n > 0 && !xyz)) return fail(ctx, PTK_E_ARG, "map update: bad argument");
    if (n > ctx->cfg.max_points) return fail(ctx, PTK_E_CAPACITY, "cloud larger than cfg.max_points");
    CK(cudaSetDevice(ctx->device));
    LaneHost& LH = ctx->lanes[lane];
    StepOut& H = ctx->h_outs[lane];
    int use_pose = 0;
    if (pose) {
        rigid_from_mat16(pose, H.pose);
        CK(cudaMemcpyAsync(ctx->d_outs + lane, &H, sizeof(StepOut), cudaMemcpyHostToDevice, st));
        use_pose = 1;
    }
    const double* origin_dev = nullptr;
    if (origin3) {
        CK(cudaMemcpyAsync(ctx->d_tmp, origin3, 3 * sizeof(double), cudaMemcpyHostToDevice, st));
        origin_dev = ctx->d_tmp;
    }
    if (insert) {
        const double* dx;
        int rc = stage_in(ctx, xyz, (size_t)n * 3, LH.in_xyz, &dx, st);
        if (rc) return rc;
        LAUNCH(PS_OTHER, st, k_load_ds<<<std::max(1, (n + 255) / 256), 256, 0, st>>>(ctx->d_lanes, lane, dx, n));
        CK(cudaGetLastError());
    }
    int rc = map_update_launch(ctx, lane, 1, std::max(n, 1), use_pose, origin_dev, insert, prune, st);
    if (rc) return rc;
    StepOut O;
    rc = lane_counters(ctx, lane, &O, st);
    if (rc) return rc;
    LH.last_out.n_vox = O.n_vox; LH.last_out.map_points = O.map_points;
    int ec = err_to_code(ctx, O.err);
    if (ec) return ec;
    return maybe_rebuild(ctx, lane, O, st);
}

extern "C" int ptk_map_update(ptk_ctx* ctx, int lane, const double* xyz, int n, const double* pose, void* stream) {
    if (!pose) return fail(ctx, PTK_E_ARG, "ptk_map_update: pose is null");
    return map_mutate(ctx, lane, xyz, n, pose, nullptr, true, true, (cudaStream_t)stream);
}

extern "C" int ptk_map_add_points(ptk_ctx* ctx, int lane, const double* xyz, int n, void* stream) {
    return map_mutate(ctx, lane, xyz, n, nullptr, nullptr, true, false, (cudaStream_t)stream);
}

extern "C" int ptk_map_remove_far(ptk_ctx* ctx, int lane, const double* origin3, void* stream) {
    if (!origin3) return fail(ctx, PTK_E_ARG, "ptk_map_remove_far: origin is null");
    return map_mutate(ctx, lane, nullptr, 0, nullptr, origin3, false, true, (cudaStream_t)stream);
}

static int ensure_big(ptk_ctx* ctx, size_t bytes) {
    if (ctx->big_tmp_bytes >= bytes) return PTK_OK;
    if (ctx->d_big) cudaFree(ctx->d_big);
    ctx->d_big = nullptr; ctx->big_tmp_bytes = 0;
    CK(cudaMalloc(&ctx->d_big, bytes));
    ctx->big_tmp_bytes = bytes;
    return PTK_OK;
}

extern "C" int ptk_map_point_cloud(ptk_ctx* ctx, int lane, double* out_xyz, int capacity, int* n_out, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || capacity < 0) return fail(ctx, PTK_E_ARG, "ptk_map_point_cloud: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_big(ctx, (size_t)std::max(capacity, 1) * 3 * sizeof(double));
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_tmp_i, 0, 2 * sizeof(int), st));
    ctx->launches++;
    k_map_dump<<<ctx->num_sms, 256, 0, st>>>(ctx->d_lanes, lane, nullptr, nullptr, nullptr, (double*)ctx->d_big, capacity,
                                           ctx->d_tmp_i, ctx->d_tmp_i + 1);
    CK(cudaGetLastError());
    int cnt[2];
    CK(cudaMemcpyAsync(cnt, ctx->d_tmp_i, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_out) *n_out = cnt[1];
    if (cnt[1] > capacity) return fail(ctx, PTK_E_CAPACITY, "output buffer too small");
    if (out_xyz && cnt[1] > 0) {
        rc = copy_out(ctx, out_xyz, ctx->d_big, (size_t)cnt[1] * 3 * sizeof(double), st);
        if (rc) return rc;
        CK(cudaStreamSynchronize(st));
    }
    return PTK_OK;
}

extern "C" int ptk_map_dump(ptk_ctx* ctx, int lane, int* keys, int* counts, double* points, int capacity,
                            int* n_voxels, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || capacity < 0) return fail(ctx, PTK_E_ARG, "ptk_map_dump: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    size_t cap = (size_t)std::max(capacity, 1);
    size_t b_keys = cap * 3 * sizeof(int), b_cnt = cap * sizeof(int), b_pts = cap * MAXP * 3 * sizeof(double);
    int rc = ensure_big(ctx, b_pts + b_keys + b_cnt + 64);
    if (rc) return rc;
    double* d_pts = (double*)ctx->d_big;
    int* d_keys = (int*)((char*)ctx->d_big + b_pts);
    int* d_cnt = (int*)((char*)ctx->d_big + b_pts + b_keys);
    CK(cudaMemsetAsync(ctx->d_tmp_i, 0, 2 * sizeof(int), st));
    ctx->launches++;
    k_map_dump<<<ctx->num_sms, 256, 0, st>>>(ctx->d_lanes, lane, d_keys, d_cnt, d_pts, nullptr, capacity, ctx->d_tmp_i,
                                           ctx->d_tmp_i + 1);
    CK(cudaGetLastError());
    int cnt[2];
    CK(cudaMemcpyAsync(cnt, ctx->d_tmp_i, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_voxels) *n_voxels = cnt[0];
    if (cnt[0] > capacity) return fail(ctx, PTK_E_CAPACITY, "output buffer too small");
    size_t v = (size_t)cnt[0];
    if (v > 0) {
        if (keys && (rc = copy_out(ctx, keys, d_keys, v * 3 * sizeof(int), st))) return rc;
        if (counts && (rc = copy_out(ctx, counts, d_cnt, v * sizeof(int), st))) return rc;
        if (points && (rc = copy_out(ctx, points, d_pts, v * MAXP * 3 * sizeof(double), st))) return rc;
        CK(cudaStreamSynchronize(st));
    }
    return PTK_OK;
}

extern "C" int ptk_map_get_correspondences(ptk_ctx* ctx, int lane, const double* xyz, int n, double max_dist,
                                           int* out_order, double* out_target, int* n_corr, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || n < 0 || (n > 0 && !xyz)) return fail(ctx, PTK_E_ARG, "ptk_map_get_correspondences: bad argument");
    if (n > ctx->cfg.max_points) return fail(ctx, PTK_E_CAPACITY, "cloud larger than cfg.max_points");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    const double* dq;
    int rc = stage_in(ctx, xyz, (size_t)n * 3, ctx->lanes[lane].in_xyz, &dq, st);
    if (rc) return rc;
    int* d_ord = ctx->d_tmp_i + 16;
    CK(cudaMemsetAsync(ctx->d_tmp_i, 0, sizeof(int), st));
    if (n > 0) {
        size_t threads = (size_t)n * 32;
        ctx->launches++;
        k_correspondences<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(ctx->d_lanes, lane, dq, n, max_dist, d_ord,
                                                                           ctx->d_tmp, ctx->d_tmp_i);
        CK(cudaGetLastError());
        if (out_order && (rc = copy_out(ctx, out_order, d_ord, (size_t)n * sizeof(int), st))) return rc;
        if (out_target && (rc = copy_out(ctx, out_target, ctx->d_tmp, (size_t)n * 3 * sizeof(double), st))) return rc;
    }
    int nc = 0;
    CK(cudaMemcpyAsync(&nc, ctx->d_tmp_i, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_corr) *n_corr = nc;
    return PTK_OK;
}

extern "C" int ptk_register_point_cloud(ptk_ctx* ctx, int lane, const double* xyz, int n, const double* initial_guess,
                                        double max_correspondance_distance, double kernel, double* out_pose,
                                        ptk_stats* stats, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || n < 0 || (n > 0 && !xyz) || !initial_guess || !out_pose)
        return fail(ctx, PTK_E_ARG, "ptk_register_point_cloud: bad argument");
    if (n > ctx->cfg.max_points) return fail(ctx, PTK_E_CAPACITY, "cloud larger than cfg.max_points");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    LaneHost& LH = ctx->lanes[lane];
    StepParams& P = ctx->h_params[lane];
    memset(&P, 0, sizeof(P));
    rigid_from_mat16(initial_guess, P.guess);
    P.max_corr = max_correspondance_distance;
    P.kernel = kernel;
    P.n = n;
    P.epoch = ++LH.epoch;
    P.tbase1 = LH.tbase1; P.tbase2 = LH.tbase2; P.release_base = LH.release_base;
    LH.release_base += (u32)ctx->cfg.max_iterations + 1u;
    const double* dx;
    int rc = stage_in(ctx, xyz, (size_t)n * 3, LH.in_xyz, &dx, st);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_params + lane, &P, sizeof(StepParams), cudaMemcpyHostToDevice, st));
    LAUNCH(PS_OTHER, st, k_load_src<<<std::max(1, (n + 255) / 256), 256, 0, st>>>(ctx->d_lanes, lane, dx, n, P.guess));
    CK(cudaGetLastError());
    rc = launch_icp(ctx, lane, 1, std::max(1, (n + 31) / 32), st);
    if (rc) return rc;
    StepOut O;
    rc = lane_counters(ctx, lane, &O, st);
    if (rc) return rc;
    LH.last_reg_iters = O.iterations;
    LH.last_reg_nsrc = n;
    LH.have_last = true;
    LH.last_out.n_src = n;
    rigid_to_mat16(O.pose, out_pose);
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->status = O.status; stats->n_in = n; stats->n_src = n; stats->iterations = O.iterations;
        stats->n_corr = O.n_corr; stats->dx_norm = O.dx_norm; stats->n_voxels = O.n_vox; stats->map_points = O.map_points;
        stats->icp_searches = O.icp_searches;
    }
    if (O.status == 2) return fail(ctx, PTK_E_NUMERIC, "singular normal equations in ICP");
    return PTK_OK;
}

// ---- hash-sharded map over several GPUs: one process per GPU drives these between its collectives ----
extern "C" int ptk_shard_config(ptk_ctx* ctx, int rank, int nranks) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || nranks > 32) return fail(ctx, PTK_E_ARG, "ptk_shard_config: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    for (size_t l = 0; l < ctx->lanes.size(); ++l) {
        LaneDev& d = ctx->lanes[l].d;
        d.shard_rank = rank; d.shard_n = nranks;
        int v[2] = {rank, nranks};
        CK(cudaMemcpy((char*)(ctx->d_lanes + l) + offsetof(LaneDev, shard_rank), v, sizeof(v), cudaMemcpyHostToDevice));
    }
    return PTK_OK;
}

// ---- in-kernel exchange through peer memory ------------------------------------------------------------------
// Layout of a rank's buffer: records [nranks][2][5][max_points] f64 | stamps [nranks][XCH_FLAGS] u64 | voxel counts [nranks] i32
static size_t xch_rec_bytes(const ptk_ctx* ctx, int G) { return (size_t)G * 2 * 5 * (size_t)ctx->cfg.max_points * sizeof(double); }
static size_t xch_flag_bytes(int G) { return (size_t)G * XCH_FLAGS * sizeof(unsigned long long); }

extern "C" int ptk_shard_peer_export(ptk_ctx* ctx, unsigned char* handle64, void** local_ptr, unsigned long long* bytes) {
    if (!ctx) return PTK_E_ARG;
    const int G = ctx->lanes[0].d.shard_n;
    if (G < 2 || G > PTK_MAX_PEERS) return fail(ctx, PTK_E_STATE, "ptk_shard_peer_export: call ptk_shard_config with 2..8 ranks first");
    if (ctx->cfg.max_iterations > 1000000) return fail(ctx, PTK_E_ARG, "ptk_shard_peer_export: max_iterations too large");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->xch_local) {
        ctx->xch_bytes = xch_rec_bytes(ctx, G) + xch_flag_bytes(G) + 256;
        CK(cudaMalloc(&ctx->xch_local, ctx->xch_bytes));
        CK(cudaMemset(ctx->xch_local, 0, ctx->xch_bytes));
        CK(cudaDeviceSynchronize());
    }
    if (handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ctx->xch_local));
        memcpy(handle64, &h, 64);
    }
    if (local_ptr) *local_ptr = ctx->xch_local;
    if (bytes) *bytes = ctx->xch_bytes;
    return PTK_OK;
}

extern "C" int ptk_shard_peer_attach(ptk_ctx* ctx, int rank, const unsigned char* handle64, void* direct_ptr) {
    if (!ctx) return PTK_E_ARG;
    const int G = ctx->lanes[0].d.shard_n, me = ctx->lanes[0].d.shard_rank;
    if (!ctx->xch_local) return fail(ctx, PTK_E_STATE, "ptk_shard_peer_attach: ptk_shard_peer_export first");
    if (rank < 0 || rank >= G) return fail(ctx, PTK_E_ARG, "ptk_shard_peer_attach: rank out of range");
    CK(cudaSetDevice(ctx->device));
    if (rank == me) ctx->xch_peer[rank] = ctx->xch_local;
    else if (direct_ptr) ctx->xch_peer[rank] = direct_ptr;            // same process (peer access is the caller's business)
    else if (handle64) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle64, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->xch_peer[rank] = p;
        ctx->xch_ipc[rank] = true;
    } else return fail(ctx, PTK_E_ARG, "ptk_shard_peer_attach: neither handle nor pointer");
    ctx->xch_peer[me] = ctx->xch_local;
    for (int r = 0; r < G; ++r)
        if (!ctx->xch_peer[r]) return PTK_OK;                         // more ranks to come
    // every rank is mapped: hand the pointers to the lanes
    CK(cudaDeviceSynchronize());
    for (size_t l = 0; l < ctx->lanes.size(); ++l) {
        LaneDev& d = ctx->lanes[l].d;
        for (int r = 0; r < G; ++r) {
            char* base = (char*)ctx->xch_peer[r];
            d.xch_rec[r] = (double*)base;
            d.xch_flag[r] = (unsigned long long*)(base + xch_rec_bytes(ctx, G));
            d.xch_nvox[r] = (int*)(base + xch_rec_bytes(ctx, G) + xch_flag_bytes(G));
        }
        CK(cudaMemcpy((char*)(ctx->d_lanes + l) + offsetof(LaneDev, xch_rec), d.xch_rec,
                      offsetof(LaneDev, n_range) - offsetof(LaneDev, xch_rec), cudaMemcpyHostToDevice));
    }
    return PTK_OK;
}

extern "C" int ptk_shard_begin(ptk_ctx* ctx, int lane, const double* xyz, const double* timestamps, int n,
                               const unsigned int* range_mm, const double* initial_guess, int* n_src, int* n_vox_local,
                               void* stream) {
    if (!ctx) return PTK_E_ARG;
    if (lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "lane out of range");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char hg = initial_guess ? 1 : 0;
    LaneHost& LH = ctx->lanes[lane];
    int rc = step_prepare(ctx, lane, 1, &xyz, &timestamps, &n, range_mm ? &range_mm : nullptr, initial_guess, &hg,
                          &LH.pending_nmax, st);
    if (rc) return rc;
    int v[2] = {0, 0};
    CK(cudaMemcpyAsync(&v[0], (char*)(ctx->d_lanes + lane) + offsetof(LaneDev, n_src), sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&v[1], (char*)(ctx->d_lanes + lane) + offsetof(LaneDev, n_vox), sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    LH.shard_nsrc = v[0];
    if (n_src) *n_src = v[0];
    if (n_vox_local) *n_vox_local = v[1];
    return PTK_OK;
}

extern "C" int ptk_shard_search(ptk_ctx* ctx, int lane, int it, double* records, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || !records) return fail(ctx, PTK_E_ARG, "ptk_shard_search: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    int n = ctx->lanes[lane].shard_nsrc;
    if (n > 0) {
        size_t threads = (size_t)n * 32;
        LAUNCH(PS_ICP, st, k_shard_search<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(ctx->d_lanes, lane, ctx->d_params, it, records, n));
        CK(cudaGetLastError());
    }
    return PTK_OK;
}

static void shard_slice(int n, int G, int rank, int* g_lo, int* g_cnt, int* n_roots) {
    int n_groups = (n + 31) / 32;
    int P = 1;
    while (P < n_groups) P <<= 1;
    int sg = std::max(1, P / G);
    *n_roots = P / sg;
    *g_lo = rank * sg;
    *g_cnt = rank < *n_roots ? sg : 0;
}

extern "C" int ptk_shard_system(ptk_ctx* ctx, int lane, const double* gathered, int it, double* partials, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || !gathered || !partials) return fail(ctx, PTK_E_ARG, "ptk_shard_system: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    const LaneHost& LH = ctx->lanes[lane];
    int n = LH.shard_nsrc, G = LH.d.shard_n, rank = LH.d.shard_rank;
    int g_lo, g_cnt, n_roots;
    shard_slice(n, G, rank, &g_lo, &g_cnt, &n_roots);
    CK(cudaMemsetAsync(partials, 0, sizeof(double) * NSUM * 32, st));
    if (g_cnt > 0 && n > 0) {
        LAUNCH(PS_ICP, st, k_shard_system<<<(g_cnt + ICP_WARPS - 1) / ICP_WARPS, ICP_THREADS, 0, st>>>(
                   ctx->d_lanes, lane, ctx->d_params, gathered, G, n, it, g_lo, g_cnt));
        LAUNCH(PS_ICP, st, k_shard_slice_root<<<1, NSUM * 32, 0, st>>>(ctx->d_lanes, lane, n, g_lo, g_cnt, partials, rank));
        CK(cudaGetLastError());
    }
    return PTK_OK;
}

extern "C" int ptk_shard_solve(ptk_ctx* ctx, int lane, const double* partials, int it, int map_empty, int* done, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B || (!map_empty && !partials)) return fail(ctx, PTK_E_ARG, "ptk_shard_solve: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaSetDevice(ctx->device));
    const LaneHost& LH = ctx->lanes[lane];
    int g_lo, g_cnt, n_roots;
    shard_slice(LH.shard_nsrc, LH.d.shard_n, LH.d.shard_rank, &g_lo, &g_cnt, &n_roots);
    LAUNCH(PS_ICP, st, k_shard_solve<<<1, NSUM * 32, 0, st>>>(ctx->d_lanes, lane, ctx->d_params, ctx->d_outs, partials, n_roots, it, map_empty));
    CK(cudaGetLastError());
    int d = 0;
    CK(cudaMemcpyAsync(&d, (char*)(ctx->d_lanes + lane) + offsetof(LaneDev, icp_done), sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (done) *done = d;
    return PTK_OK;
}

extern "C" int ptk_shard_end(ptk_ctx* ctx, int lane, double* out_pose, ptk_stats* stats, void* stream) {
    if (!ctx || lane < 0 || lane >= ctx->B) return fail(ctx, PTK_E_ARG, "ptk_shard_end: bad argument");
    return step_finish(ctx, lane, 1, ctx->lanes[lane].pending_nmax, out_pose, stats, (cudaStream_t)stream);
}

extern "C" int ptk_get_icp_phases(const ptk_ctx* ctx, int lane, long long* cycles6) {
    if (!ctx || lane < 0 || lane >= ctx->B || !cycles6) return PTK_E_ARG;
    const LaneHost& LH = ctx->lanes[lane];
    if (!LH.have_last) return PTK_E_STATE;
    for (int k = 0; k < 6; ++k) cycles6[k] = LH.last_out.icp_cyc[k];
    return PTK_OK;
}

extern "C" int ptk_set_profiling(ptk_ctx* ctx, int on) {
    if (!ctx) return PTK_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    prof_collect(ctx);
    ctx->prof.on = on != 0;
    for (int k = 0; k < PTK_PROF_SLOTS; ++k) { ctx->prof.ms[k] = 0.0; ctx->prof.n[k] = 0; }
    return PTK_OK;
}

extern "C" int ptk_get_profile(ptk_ctx* ctx, double* ms, long long* launches) {
    if (!ctx) return PTK_E_ARG;
    for (int k = 0; k < PTK_PROF_SLOTS; ++k) {
        if (ms) ms[k] = ctx->prof.ms[k];
        if (launches) launches[k] = ctx->prof.n[k];
    }
    return PTK_OK;
}

extern "C" const char* ptk_kernel_name(int slot) { return (slot >= 0 && slot < PS_COUNT) ? kProfNames[slot] : nullptr; }

extern "C" long long ptk_launch_count(const ptk_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" void ptk_control_bytes(int* h2d_per_lane, int* d2h_per_lane) {
    if (h2d_per_lane) *h2d_per_lane = (int)sizeof(StepParams);
    if (d2h_per_lane) *d2h_per_lane = (int)sizeof(StepOut);
}

extern "C" int ptk_host_alloc(void** out, unsigned long long bytes) {
    if (!out) return PTK_E_ARG;
    cudaError_t e = cudaMallocHost(out, (size_t)bytes);
    if (e != cudaSuccess) { cudaGetLastError(); g_create_err = cudaGetErrorString(e); return PTK_E_CUDA; }
    return PTK_OK;
}

extern "C" int ptk_host_free(void* p) {
    if (p) cudaFreeHost(p);
    return PTK_OK;
}

// ---- fleet replay: several contexts, each advanced by its own host thread -----------------------------------------
// The lanes of ONE context move in lock step (every call waits for all of them), so while its ICP kernel - a chain of
// dependent iterations that leaves most of the GPU idle - runs, nothing else of that context can.  A fleet of
// independent sequences does not need the lock step: split it over a few contexts, let every context run through
// its scans on its own thread and stream, and the ICP of one overlaps the streaming kernels of the others.
extern "C" int ptk_set_icp_blocks_per_lane(ptk_ctx* ctx, int blocks) {
    if (!ctx || blocks < 0) return PTK_E_ARG;
    ctx->icp_max_blocks_per_lane = blocks == 0 ? (1 << 20) : blocks;
    return PTK_OK;
}

extern "C" int ptk_fleet_replay(ptk_ctx* const* ctxs, int n_ctx, const unsigned int* const* const* range_mm, int n_scans,
                                double* const* out_poses, ptk_stats* const* stats, void* const* streams) {
    if (!ctxs || n_ctx < 1 || !range_mm || n_scans < 0) return PTK_E_ARG;
    std::vector<int> rcs(n_ctx, PTK_OK);
    // more worker threads than host cores (8 ranks x 8 contexts on a 32-core box, ranks pinned to their share of
    // the cores): a spinning cudaStreamSynchronize per worker would starve the others, so the workers sleep
    const unsigned hw = std::thread::hardware_concurrency();
    cpu_set_t cs;
    int avail = (sched_getaffinity(0, sizeof(cs), &cs) == 0) ? CPU_COUNT(&cs) : (int)hw;
    const bool blocking = n_ctx > 1 && n_ctx + 1 > std::max(1, avail);
    auto work = [&](int g) {
        ptk_ctx* ctx = ctxs[g];
        ctx->blocking_sync = blocking;
        const int B = ctx->B;
        cudaStream_t st = streams ? (cudaStream_t)streams[g] : nullptr;
        for (int s = 0; s < n_scans; ++s) {
            const unsigned int* const* cur = range_mm[g] + (size_t)s * B;
            if (s + 1 < n_scans && !is_device_ptr(cur[0])) {
                int prc = ptk_prefetch_scan_batch(ctx, range_mm[g] + (size_t)(s + 1) * B);
                if (prc) { rcs[g] = prc; return; }
            }
            int rc = ptk_register_scan_batch(ctx, cur, nullptr, nullptr, out_poses ? out_poses[g] + (size_t)s * B * 16 : nullptr,
                                             (stats && stats[g]) ? stats[g] + (size_t)s * B : nullptr, st);
            if (rc) { rcs[g] = rc; ctx->blocking_sync = false; return; }
        }
        ctx->blocking_sync = false;
    };
    if (n_ctx == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < n_ctx; ++g) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < n_ctx; ++g)
        if (rcs[g]) return rcs[g];
    return PTK_OK;
}
