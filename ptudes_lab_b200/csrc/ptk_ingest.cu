// Ingest: raw Ouster UDP packets -> staggered LidarScan field images in HBM (include/ptk.h, "ingest").
//
// Reference path: /root/reference/src/ptudes/data.py:31-77 (OusterLidarData.withScanIdx drives ouster-sdk's
// PacketFormat / ScanBatcher per packet, on the host), fed by pcap.Pcap / OusterRawBagSource
// (/root/reference/src/ptudes/utils.py:171-187, /root/reference/src/ptudes/bag.py:21-97).
// Here: packets of a frame are grouped on the host (ptk_batcher: frame-id logic only, no per-pixel work),
// cross the bus as they came off the wire, and ONE kernel per batch of frames scatters the channel data into
// the (H, W) images: a block per packet, the packet staged in shared memory by one TMA bulk copy, columns
// written to their measurement ids.  Byte work, HBM-bound: reads every packet byte once, writes every pixel once.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ptk.h"

namespace ptk_ingest {

thread_local std::string g_ingest_err;

#define ICK(x)                                                                                   \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            g_ingest_err = std::string(#x) + ": " + cudaGetErrorString(e_);                      \
            return PTK_E_CUDA;                                                                   \
        }                                                                                        \
    } while (0)

// ---- packet layout ---------------------------------------------------------------------------------
struct DevFormat {
    int profile, H, cpp, W;
    int pkt_hdr, col_hdr, ch_size, col_size, pkt_size, ppf;
};

__host__ __device__ inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
__host__ __device__ inline uint32_t rd32(const unsigned char* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__host__ __device__ inline uint64_t rd64(const unsigned char* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

// column header of column c: timestamp, measurement id, status word (bit 0 = valid)
__host__ __device__ inline void col_header(const DevFormat& F, const unsigned char* pkt, int c, uint64_t& ts, int& mid, uint32_t& status) {
    const unsigned char* col = pkt + F.pkt_hdr + (size_t)c * F.col_size;
    ts = rd64(col);
    mid = rd16(col + 8);
    if (F.profile == PTK_PROFILE_LEGACY) status = rd32(col + F.col_size - 4);      // 0xFFFFFFFF valid, 0 not
    else status = rd16(col + 10);
}

// ---- TMA (1-D bulk copy global -> shared, completion on an mbarrier) ---------------------------------
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// A block decodes one packet slot (slot index = frame * packets_per_frame + position): ONE TMA bulk copy brings the
// packet into shared memory, the block scatters it into the images.  With eight resident blocks per SM the loads of
// some overlap the stores of the others: 4.0 TB/s = 62 % of the measured HBM peak on 384 frames per launch.
// The same kernel also runs as a PERSISTENT grid (PTK_DECODE_BLOCKS_PER_SM = n) in which a block walks several slots
// with two buffers, the next packet in flight while the current one is decoded - measured slower (31-44 %): four
// double-buffered blocks per SM keep fewer loads in flight than eight single-buffered ones.
__device__ __forceinline__ void tma_packet(unsigned char* dst, const unsigned char* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(bar)) : "memory");
}

__global__ void __launch_bounds__(256) k_decode_packets(const unsigned char* __restrict__ packets, DevFormat F, ptk_scan_fields out,
                                                        unsigned int* frame_done, int n_slots, int use_tma, int vec_ok) {
    extern __shared__ __align__(16) unsigned char s_buf[];
    __shared__ int s_mid[64];
    __shared__ unsigned long long s_bar[2];
    __shared__ int s_last;
    __shared__ int s_nmiss;
    __shared__ unsigned short s_misscol[256];
    const uint32_t stride = ((uint32_t)F.pkt_size + 15u) & ~15u;
    if (use_tma) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&s_bar[0])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(saddr(&s_bar[1])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if ((int)blockIdx.x < n_slots) tma_packet(s_buf, packets + (size_t)blockIdx.x * F.pkt_size, (uint32_t)F.pkt_size, &s_bar[0]);
        }
        __syncthreads();            // the barriers are initialised before anybody polls them
    }
    int it = 0;
    for (int idx = blockIdx.x; idx < n_slots; idx += gridDim.x, ++it) {
        const int cur = it & 1;
        unsigned char* s_pkt = s_buf + (size_t)cur * stride;
        const int f = idx / F.ppf;
        if (use_tma) {
            // the other buffer was released by the barrier that ended the previous iteration
            const int nxt = idx + gridDim.x;
            if (threadIdx.x == 0 && nxt < n_slots)
                tma_packet(s_buf + (size_t)(cur ^ 1) * stride, packets + (size_t)nxt * F.pkt_size, (uint32_t)F.pkt_size, &s_bar[cur ^ 1]);
            const uint32_t parity = (uint32_t)(it >> 1) & 1u;
            uint32_t ok;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(saddr(&s_bar[cur])), "r"(parity) : "memory");
            } while (!ok);
        } else {
            const unsigned char* g = packets + (size_t)idx * F.pkt_size;
            for (int i = threadIdx.x; i < F.pkt_size; i += blockDim.x) s_pkt[i] = g[i];
            __syncthreads();
        }
        const size_t img = (size_t)f * F.H * F.W, hdr = (size_t)f * F.W;
        if ((int)threadIdx.x < F.cpp) {
            uint64_t ts; int mid; uint32_t st;
            col_header(F, s_pkt, threadIdx.x, ts, mid, st);
            const bool valid = (st & 1u) && mid < F.W;
            s_mid[threadIdx.x] = valid ? mid : -1;
            if (valid) {
                out.timestamp[hdr + mid] = ts;
                out.status[hdr + mid] = st;
                out.measurement_id[hdr + mid] = (unsigned short)mid;
            }
        }
        __syncthreads();
        // channel data -> images.  Fast path (every Ouster format: 16 columns per packet, measurement ids of a packet
        // consecutive and aligned): a thread takes one row of FOUR columns and writes its four pixels with one 16 B
        // (u32 fields) / 8 B (u16 fields) store; four threads cover the 64 B a packet contributes to an image row.
        bool quads = (F.cpp % 4 == 0) && vec_ok;
        if (quads) {
            for (int c = 0; c < F.cpp; c += 4) {
                const int m0 = s_mid[c];
                quads = quads && m0 >= 0 && (m0 & 3) == 0 && s_mid[c + 1] == m0 + 1 && s_mid[c + 2] == m0 + 2 && s_mid[c + 3] == m0 + 3;
            }
        }
        if (quads) {
            const int groups = F.cpp >> 2;
            const int n = F.H * groups;
            uint32_t rmask = 0x0007ffffu, rsh = 0;
            if (F.profile == PTK_PROFILE_LEGACY) rmask = 0x000fffffu;
            else if (F.profile == PTK_PROFILE_RNG15_RFL8_NIR8) { rmask = 0x7fffu; rsh = 3; }
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int g = i % groups, p = i / groups;
                const int c0 = g << 2;
                const unsigned char* ch = s_pkt + F.pkt_hdr + (size_t)c0 * F.col_size + F.col_hdr + (size_t)p * F.ch_size;
                const size_t o = img + (size_t)p * F.W + s_mid[c0];
                uint32_t w0[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w0[j] = *reinterpret_cast<const uint32_t*>(ch + (size_t)j * F.col_size);
                *reinterpret_cast<uint4*>(out.range + o) = make_uint4((w0[0] & rmask) << rsh, (w0[1] & rmask) << rsh, (w0[2] & rmask) << rsh,
                                                                      (w0[3] & rmask) << rsh);
                if (out.reflectivity || out.signal || out.near_ir || out.range2) {
                    unsigned short refl[4], sig[4] = {0, 0, 0, 0}, nir[4];
                    uint32_t r2[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned char* cj = ch + (size_t)j * F.col_size;
                        if (F.profile == PTK_PROFILE_LEGACY) { refl[j] = rd16(cj + 4); sig[j] = rd16(cj + 6); nir[j] = rd16(cj + 8); }
                        else if (F.profile == PTK_PROFILE_RNG19_RFL8_SIG16_NIR16) { refl[j] = cj[4]; sig[j] = rd16(cj + 6); nir[j] = rd16(cj + 8); }
                        else if (F.profile == PTK_PROFILE_RNG15_RFL8_NIR8) { refl[j] = cj[2]; nir[j] = (unsigned short)((unsigned)cj[3] << 4); }
                        else { refl[j] = cj[3]; r2[j] = rd32(cj + 4) & 0x0007ffffu; sig[j] = rd16(cj + 8); nir[j] = rd16(cj + 12); }
                    }
                    auto pack = [](const unsigned short* v) { return make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16)); };
                    if (out.reflectivity) *reinterpret_cast<uint2*>(out.reflectivity + o) = pack(refl);
                    if (out.signal && F.profile != PTK_PROFILE_RNG15_RFL8_NIR8) *reinterpret_cast<uint2*>(out.signal + o) = pack(sig);
                    if (out.near_ir) *reinterpret_cast<uint2*>(out.near_ir + o) = pack(nir);
                    if (out.range2 && F.profile == PTK_PROFILE_RNG19_RFL8_SIG16_NIR16_DUAL) *reinterpret_cast<uint4*>(out.range2 + o) = make_uint4(r2[0], r2[1], r2[2], r2[3]);
                }
            }
        } else {
            // general path: consecutive threads = consecutive columns of one row
            const int n = F.H * F.cpp;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int c = i % F.cpp, p = i / F.cpp;
                const int mid = s_mid[c];
                if (mid < 0) continue;
                const unsigned char* ch = s_pkt + F.pkt_hdr + (size_t)c * F.col_size + F.col_hdr + (size_t)p * F.ch_size;
                const size_t o = img + (size_t)p * F.W + mid;
                const uint32_t w0 = *reinterpret_cast<const uint32_t*>(ch);
                if (F.profile == PTK_PROFILE_LEGACY) {
                    out.range[o] = w0 & 0x000fffffu;
                    if (out.reflectivity) out.reflectivity[o] = rd16(ch + 4);
                    if (out.signal) out.signal[o] = rd16(ch + 6);
                    if (out.near_ir) out.near_ir[o] = rd16(ch + 8);
                } else if (F.profile == PTK_PROFILE_RNG19_RFL8_SIG16_NIR16) {
                    out.range[o] = w0 & 0x0007ffffu;
                    if (out.reflectivity) out.reflectivity[o] = ch[4];
                    if (out.signal) out.signal[o] = rd16(ch + 6);
                    if (out.near_ir) out.near_ir[o] = rd16(ch + 8);
                } else if (F.profile == PTK_PROFILE_RNG15_RFL8_NIR8) {
                    out.range[o] = (w0 & 0x7fffu) << 3;
                    if (out.reflectivity) out.reflectivity[o] = ch[2];
                    if (out.near_ir) out.near_ir[o] = (unsigned short)((unsigned)ch[3] << 4);
                } else {    // RNG19_RFL8_SIG16_NIR16_DUAL
                    out.range[o] = w0 & 0x0007ffffu;
                    if (out.reflectivity) out.reflectivity[o] = ch[3];
                    if (out.range2) out.range2[o] = rd32(ch + 4) & 0x0007ffffu;
                    if (out.signal) out.signal[o] = rd16(ch + 8);
                    if (out.near_ir) out.near_ir[o] = rd16(ch + 12);
                }
            }
        }
        // ScanBatcher's zero fill: whoever finishes the LAST packet slot of a frame clears the columns nobody wrote
        // (status still 0).  The barrier also releases this iteration's buffer and s_mid.
        __syncthreads();
        if (threadIdx.x == 0) {             // release: the barrier ordered the block's stores before this fence
            __threadfence();
            s_last = (atomicAdd(&frame_done[f], 1u) == (unsigned)F.ppf - 1u);
        }
        __syncthreads();
        if (!s_last) continue;              // block-uniform
        __threadfence();
        // (columns checked 256 at a time; the usual frame has none missing and this is one pass of W/256 loads)
        for (int w0 = 0; w0 < F.W; w0 += blockDim.x) {
            if (threadIdx.x == 0) s_nmiss = 0;
            __syncthreads();
            const int w = w0 + threadIdx.x;
            if (w < F.W && !(__ldcg(out.status + hdr + w) & 1u)) {
                s_misscol[atomicAdd(&s_nmiss, 1)] = (unsigned short)w;
                out.timestamp[hdr + w] = 0; out.status[hdr + w] = 0; out.measurement_id[hdr + w] = 0;
            }
            __syncthreads();
            const int nm = s_nmiss;
            for (int i = threadIdx.x; i < nm * F.H; i += blockDim.x) {
                const size_t o = img + (size_t)(i / nm) * F.W + s_misscol[i % nm];
                out.range[o] = 0;
                if (out.range2) out.range2[o] = 0;
                if (out.reflectivity) out.reflectivity[o] = 0;
                if (out.signal) out.signal[o] = 0;
                if (out.near_ir) out.near_ir[o] = 0;
            }
            __syncthreads();
        }
    }
}

DevFormat dev_format(const ptk_packet_format& pf) {
    DevFormat F;
    F.profile = pf.profile; F.H = pf.pixels_per_column; F.cpp = pf.columns_per_packet; F.W = pf.columns_per_frame;
    F.pkt_hdr = pf.packet_header_size; F.col_hdr = pf.col_header_size; F.ch_size = pf.channel_data_size;
    F.col_size = pf.col_size; F.pkt_size = pf.lidar_packet_size; F.ppf = pf.packets_per_frame;
    return F;
}

bool format_ok(const ptk_packet_format* pf) {
    if (!pf) return false;
    ptk_packet_format ref;
    if (ptk_packet_format_init(&ref, pf->profile, pf->pixels_per_column, pf->columns_per_packet, pf->columns_per_frame) != PTK_OK) return false;
    return memcmp(&ref, pf, sizeof(ref)) == 0;
}

// decode `n_frames` frames whose packet slots are in DEVICE memory
int decode_device(const ptk_packet_format& pf, const unsigned char* d_packets, int n_frames, const ptk_scan_fields& out, cudaStream_t st) {
    if (n_frames == 0) return PTK_OK;
    const DevFormat F = dev_format(pf);
    unsigned int* d_done = nullptr;
    ICK(cudaMallocAsync((void**)&d_done, sizeof(unsigned int) * n_frames, st));
    ICK(cudaMemsetAsync(d_done, 0, sizeof(unsigned int) * n_frames, st));
    // a column counts as written when bit 0 of its status is set: start from "nothing written"
    ICK(cudaMemsetAsync(out.status, 0, sizeof(unsigned int) * (size_t)n_frames * F.W, st));
    const int use_tma = (F.pkt_size % 16 == 0) && (((uintptr_t)d_packets) % 16 == 0);
    const size_t stride = ((size_t)F.pkt_size + 15) & ~(size_t)15;
    const int n_slots = n_frames * F.ppf;
    int grid = std::max(1, n_slots), bufs = 1;
    int dev = 0, sms = 0;
    ICK(cudaGetDevice(&dev));
    ICK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (const char* e = getenv("PTK_DECODE_BLOCKS_PER_SM")) {       // tuning: persistent grid of n blocks per SM, two buffers each
        const int v = atoi(e);
        if (v > 0) { grid = std::max(1, std::min(n_slots, sms * v)); bufs = 2; }
    }
    const size_t smem = bufs * stride;
    // (raised once per device and size: the attribute call is not free, and this runs once per decoded frame)
    static std::mutex attr_mu;
    static size_t attr_set[64] = {0};
    if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> lk(attr_mu);
        if (attr_set[dev & 63] < smem) {
            ICK(cudaFuncSetAttribute(k_decode_packets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set[dev & 63] = smem;
        }
    }
    // vector stores need the images 16 B / 8 B aligned and rows a multiple of four columns
    auto al = [](const void* q, uintptr_t a) { return q == nullptr || ((uintptr_t)q % a) == 0; };
    const int vec_ok = (F.W % 4 == 0) && al(out.range, 16) && al(out.range2, 16) && al(out.reflectivity, 8) && al(out.signal, 8) &&
                       al(out.near_ir, 8);
    k_decode_packets<<<grid, 256, smem, st>>>(d_packets, F, out, d_done, n_slots, use_tma, vec_ok);
    ICK(cudaGetLastError());
    ICK(cudaFreeAsync(d_done, st));
    return PTK_OK;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace ptk_ingest
using namespace ptk_ingest;

// ---- C ABI -------------------------------------------------------------------------------------------
extern "C" int ptk_packet_format_init(ptk_packet_format* pf, int profile, int pixels_per_column, int columns_per_packet,
                                      int columns_per_frame) {
    if (!pf || pixels_per_column < 1 || columns_per_packet < 1 || columns_per_packet > 64 || columns_per_frame < columns_per_packet ||
        columns_per_frame % columns_per_packet != 0)
        return PTK_E_ARG;
    memset(pf, 0, sizeof(*pf));
    pf->profile = profile;
    pf->pixels_per_column = pixels_per_column;
    pf->columns_per_packet = columns_per_packet;
    pf->columns_per_frame = columns_per_frame;
    switch (profile) {
        case PTK_PROFILE_LEGACY:
            pf->packet_header_size = 0; pf->col_header_size = 16; pf->channel_data_size = 12; pf->col_footer_size = 4; pf->packet_footer_size = 0;
            break;
        case PTK_PROFILE_RNG19_RFL8_SIG16_NIR16:
            pf->packet_header_size = 32; pf->col_header_size = 12; pf->channel_data_size = 12; pf->col_footer_size = 0; pf->packet_footer_size = 32;
            break;
        case PTK_PROFILE_RNG15_RFL8_NIR8:
            pf->packet_header_size = 32; pf->col_header_size = 12; pf->channel_data_size = 4; pf->col_footer_size = 0; pf->packet_footer_size = 32;
            break;
        case PTK_PROFILE_RNG19_RFL8_SIG16_NIR16_DUAL:
            pf->packet_header_size = 32; pf->col_header_size = 12; pf->channel_data_size = 16; pf->col_footer_size = 0; pf->packet_footer_size = 32;
            break;
        default:
            return PTK_E_ARG;
    }
    pf->col_size = pf->col_header_size + pixels_per_column * pf->channel_data_size + pf->col_footer_size;
    pf->lidar_packet_size = pf->packet_header_size + columns_per_packet * pf->col_size + pf->packet_footer_size;
    pf->packets_per_frame = columns_per_frame / columns_per_packet;
    return PTK_OK;
}

extern "C" int ptk_packet_frame_id(const ptk_packet_format* pf, const unsigned char* packet) {
    if (!pf || !packet) return PTK_E_ARG;
    if (pf->profile == PTK_PROFILE_LEGACY) return rd16(packet + 10);      // first column header
    return rd16(packet + 2);                                              // packet header
}

extern "C" int ptk_decode_packets(const ptk_packet_format* pf, int device, const unsigned char* packets, int n_frames,
                                  const ptk_scan_fields* out, void* stream) {
    if (!format_ok(pf) || !packets || n_frames < 0 || !out || !out->range || !out->timestamp || !out->status || !out->measurement_id)
        return PTK_E_ARG;
    ICK(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    if (is_device_ptr(packets)) return decode_device(*pf, packets, n_frames, *out, st);
    const size_t bytes = (size_t)n_frames * pf->packets_per_frame * pf->lidar_packet_size;
    unsigned char* d = nullptr;
    ICK(cudaMallocAsync((void**)&d, std::max(bytes, (size_t)16), st));
    ICK(cudaMemcpyAsync(d, packets, bytes, cudaMemcpyHostToDevice, st));
    int rc = decode_device(*pf, d, n_frames, *out, st);
    cudaFreeAsync(d, st);
    if (rc == PTK_OK) ICK(cudaStreamSynchronize(st));     // the host buffer is the caller's again
    return rc;
}

// ---- ScanBatcher over whole frames --------------------------------------------------------------------
struct ptk_batcher {
    ptk_packet_format pf;
    int device = -1;
    int frames = 0;
    size_t frame_bytes = 0;
    unsigned char* h_ring = nullptr;       // [frames][ppf][packet]: pinned when a device is attached
    unsigned char* d_ring = nullptr;       // the same slots on the device
    std::vector<cudaEvent_t> copied;       // per ring slot: its H2D copy has completed
    std::vector<char> copy_pending;
    struct Frame { int ring; int frame_id; int n_packets; std::vector<char> present; };
    std::deque<Frame> closed;
    bool open = false;
    Frame cur;
    int next_ring = 0;
    long long dropped = 0;
};

static int batcher_close(ptk_batcher* b) {
    // lost packets read as zero bytes: no column of theirs has its valid bit set
    for (int s = 0; s < b->pf.packets_per_frame; ++s)
        if (!b->cur.present[s])
            memset(b->h_ring + (size_t)b->cur.ring * b->frame_bytes + (size_t)s * b->pf.lidar_packet_size, 0, b->pf.lidar_packet_size);
    b->closed.push_back(b->cur);
    b->open = false;
    return PTK_OK;
}

static int batcher_open(ptk_batcher* b, int frame_id) {
    if ((int)b->closed.size() >= b->frames - 1) { g_ingest_err = "ptk_batcher: every frame slot holds a closed frame; decode or pop first"; return PTK_E_STATE; }
    const int r = b->next_ring;
    b->next_ring = (b->next_ring + 1) % b->frames;
    if (b->device >= 0 && b->copy_pending[r]) {       // the slot's previous frame may still be on its way to the device
        ICK(cudaEventSynchronize(b->copied[r]));
        b->copy_pending[r] = 0;
    }
    b->cur.ring = r; b->cur.frame_id = frame_id; b->cur.n_packets = 0;
    b->cur.present.assign(b->pf.packets_per_frame, 0);
    b->open = true;
    return PTK_OK;
}

extern "C" int ptk_batcher_create(ptk_batcher** out, int device, const ptk_packet_format* pf, int frames) {
    if (!out || !format_ok(pf) || frames < 2) return PTK_E_ARG;
    ptk_batcher* b = new ptk_batcher();
    b->pf = *pf; b->device = device; b->frames = frames;
    b->frame_bytes = (size_t)pf->packets_per_frame * pf->lidar_packet_size;
    auto fail = [&](int rc) { ptk_batcher_destroy(b); return rc; };
    if (device >= 0) {
        if (cudaSetDevice(device) != cudaSuccess) { g_ingest_err = "cudaSetDevice failed"; cudaGetLastError(); return fail(PTK_E_CUDA); }
        if (cudaMallocHost((void**)&b->h_ring, b->frame_bytes * frames) != cudaSuccess ||
            cudaMalloc((void**)&b->d_ring, b->frame_bytes * frames) != cudaSuccess) {
            g_ingest_err = "ptk_batcher: allocation failed"; cudaGetLastError(); return fail(PTK_E_CUDA);
        }
        b->copied.resize(frames); b->copy_pending.assign(frames, 0);
        for (int i = 0; i < frames; ++i)
            if (cudaEventCreateWithFlags(&b->copied[i], cudaEventDisableTiming) != cudaSuccess) { b->copied.resize(i); return fail(PTK_E_CUDA); }
    } else {
        b->h_ring = (unsigned char*)malloc(b->frame_bytes * frames);
        if (!b->h_ring) return fail(PTK_E_CAPACITY);
    }
    *out = b;
    return PTK_OK;
}

extern "C" int ptk_batcher_destroy(ptk_batcher* b) {
    if (!b) return PTK_OK;
    if (b->device >= 0) {
        cudaSetDevice(b->device);
        for (cudaEvent_t e : b->copied) cudaEventDestroy(e);
        if (b->h_ring) cudaFreeHost(b->h_ring);
        if (b->d_ring) cudaFree(b->d_ring);
    } else {
        free(b->h_ring);
    }
    delete b;
    return PTK_OK;
}

extern "C" int ptk_batcher_push(ptk_batcher* b, const unsigned char* packet, int* frames_ready) {
    if (!b || !packet) return PTK_E_ARG;
    const int fid = ptk_packet_frame_id(&b->pf, packet);
    if (b->open && b->cur.frame_id != fid) {
        if (b->cur.frame_id == ((fid + 1) & 0xffff)) {       // a straggler of the previous frame: ScanBatcher drops it
            ++b->dropped;
            if (frames_ready) *frames_ready = (int)b->closed.size();
            return PTK_OK;
        }
        int rc = batcher_close(b);
        if (rc) return rc;
    }
    if (!b->open) {
        int rc = batcher_open(b, fid);
        if (rc) return rc;
    }
    // the slot of a packet follows from the measurement id of any VALID column (sensors send whole, aligned column
    // groups; a column without the valid bit may carry a blank header); a packet without a valid column has nothing
    // to contribute
    DevFormat F = dev_format(b->pf);
    int mid = -1;
    for (int c = 0; c < F.cpp; ++c) {
        uint64_t ts; int m; uint32_t st;
        col_header(F, packet, c, ts, m, st);
        if ((st & 1u) && m >= c && (m - c) % F.cpp == 0 && (m - c) / F.cpp < F.ppf) { mid = m - c; break; }
    }
    if (mid < 0) { ++b->dropped; if (frames_ready) *frames_ready = (int)b->closed.size(); return PTK_OK; }
    const int slot = mid / F.cpp;
    memcpy(b->h_ring + (size_t)b->cur.ring * b->frame_bytes + (size_t)slot * F.pkt_size, packet, F.pkt_size);
    if (!b->cur.present[slot]) { b->cur.present[slot] = 1; ++b->cur.n_packets; }
    if (frames_ready) *frames_ready = (int)b->closed.size();
    return PTK_OK;
}

extern "C" int ptk_batcher_flush(ptk_batcher* b, int* frames_ready) {
    if (!b) return PTK_E_ARG;
    if (b->open) {
        int rc = batcher_close(b);
        if (rc) return rc;
    }
    if (frames_ready) *frames_ready = (int)b->closed.size();
    return PTK_OK;
}

extern "C" int ptk_batcher_peek(ptk_batcher* b, const unsigned char** packets, int* frame_id, int* n_packets) {
    if (!b) return PTK_E_ARG;
    if (b->closed.empty()) { g_ingest_err = "ptk_batcher: no closed frame"; return PTK_E_STATE; }
    const ptk_batcher::Frame& f = b->closed.front();
    if (packets) *packets = b->h_ring + (size_t)f.ring * b->frame_bytes;
    if (frame_id) *frame_id = f.frame_id;
    if (n_packets) *n_packets = f.n_packets;
    return PTK_OK;
}

extern "C" int ptk_batcher_pop(ptk_batcher* b) {
    if (!b) return PTK_E_ARG;
    if (b->closed.empty()) { g_ingest_err = "ptk_batcher: no closed frame"; return PTK_E_STATE; }
    b->closed.pop_front();
    return PTK_OK;
}

extern "C" int ptk_batcher_decode(ptk_batcher* b, const ptk_scan_fields* out, int* frame_id, int* n_packets, void* stream) {
    if (!b || !out || !out->range || !out->timestamp || !out->status || !out->measurement_id) return PTK_E_ARG;
    if (b->device < 0) { g_ingest_err = "ptk_batcher: created without a device"; return PTK_E_STATE; }
    if (b->closed.empty()) { g_ingest_err = "ptk_batcher: no closed frame"; return PTK_E_STATE; }
    ICK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const ptk_batcher::Frame f = b->closed.front();
    unsigned char* d = b->d_ring + (size_t)f.ring * b->frame_bytes;
    ICK(cudaMemcpyAsync(d, b->h_ring + (size_t)f.ring * b->frame_bytes, b->frame_bytes, cudaMemcpyHostToDevice, st));
    ICK(cudaEventRecord(b->copied[f.ring], st));
    b->copy_pending[f.ring] = 1;
    int rc = decode_device(b->pf, d, 1, *out, st);
    if (rc) return rc;
    if (frame_id) *frame_id = f.frame_id;
    if (n_packets) *n_packets = f.n_packets;
    b->closed.pop_front();
    return PTK_OK;
}

// ---- pcap reader (classic libpcap file format) ----------------------------------------------------------
struct ptk_pcap {
    FILE* f = nullptr;
    bool swap = false, nanos = false;
    uint32_t linktype = 1;
    struct Key { uint32_t src, dst; uint16_t id; bool operator<(const Key& o) const { return src != o.src ? src < o.src : (dst != o.dst ? dst < o.dst : id < o.id); } };
    struct Frag { std::vector<unsigned char> data; std::vector<std::pair<int, int>> have; int total = -1; long long seq = 0; };
    std::map<Key, Frag> frags;
    long long seq = 0;
    std::vector<unsigned char> rec;
    // pcapng: one entry per interface description block of the current section
    bool ng = false;
    struct Iface { uint32_t linktype; double tick; };
    std::vector<Iface> ifaces;
};

static uint32_t sw32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

extern "C" int ptk_pcap_open(ptk_pcap** out, const char* path) {
    if (!out || !path) return PTK_E_ARG;
    FILE* f = fopen(path, "rb");
    if (!f) { g_ingest_err = std::string("cannot open ") + path; return PTK_E_ARG; }
    unsigned char h[24];
    if (fread(h, 1, 24, f) != 24) { fclose(f); g_ingest_err = "pcap: short file"; return PTK_E_ARG; }
    uint32_t magic;
    memcpy(&magic, h, 4);
    ptk_pcap* p = new ptk_pcap();
    if (magic == 0xa1b2c3d4u) { }
    else if (magic == 0xd4c3b2a1u) p->swap = true;
    else if (magic == 0xa1b23c4du) p->nanos = true;
    else if (magic == 0x4d3cb2a1u) { p->swap = true; p->nanos = true; }
    else if (magic == 0x0a0d0d0au) { p->ng = true; fseek(f, 0, SEEK_SET); }      // pcapng: blocks, parsed as they come
    else { fclose(f); delete p; g_ingest_err = "pcap: neither a pcap nor a pcapng file"; return PTK_E_ARG; }
    if (!p->ng) {
        uint32_t lt;
        memcpy(&lt, h + 20, 4);
        p->linktype = p->swap ? sw32(lt) : lt;
    }
    p->f = f;
    *out = p;
    return PTK_OK;
}

extern "C" int ptk_pcap_close(ptk_pcap* p) {
    if (!p) return PTK_OK;
    if (p->f) fclose(p->f);
    delete p;
    return PTK_OK;
}

extern "C" int ptk_pcap_next(ptk_pcap* p, unsigned char* buf, int cap, int* len, int* dst_port, double* ts) {
    if (!p || !buf || !len) return PTK_E_ARG;
    while (true) {
        uint32_t incl = 0, linktype = p->linktype;
        double t = 0.0;
        if (!p->ng) {
            unsigned char rh[16];
            if (fread(rh, 1, 16, p->f) != 16) return 0;
            uint32_t sec, frac;
            memcpy(&sec, rh, 4); memcpy(&frac, rh + 4, 4); memcpy(&incl, rh + 8, 4);
            if (p->swap) { sec = sw32(sec); frac = sw32(frac); incl = sw32(incl); }
            if (incl > (1u << 26)) { g_ingest_err = "pcap: corrupt record"; return PTK_E_ARG; }
            p->rec.resize(incl);
            if (incl && fread(p->rec.data(), 1, incl, p->f) != incl) return 0;
            t = (double)sec + (double)frac * (p->nanos ? 1e-9 : 1e-6);
        } else {
            // pcapng block: type, total length, body, total length again
            unsigned char bh[8];
            if (fread(bh, 1, 8, p->f) != 8) return 0;
            uint32_t type, total;
            memcpy(&type, bh, 4); memcpy(&total, bh + 4, 4);
            if (type == 0x0a0d0d0au) {                  // section header: its byte-order magic decides the endianness
                unsigned char bo[4];
                if (fread(bo, 1, 4, p->f) != 4) return 0;
                uint32_t m;
                memcpy(&m, bo, 4);
                p->swap = (m == 0x4d3c2b1au);
                if (p->swap) total = sw32(total);
                if (total < 16 || total > (1u << 26)) { g_ingest_err = "pcapng: corrupt section header"; return PTK_E_ARG; }
                fseek(p->f, (long)total - 12, SEEK_CUR);
                p->ifaces.clear();
                continue;
            }
            if (p->swap) { type = sw32(type); total = sw32(total); }
            if (total < 12 || total > (1u << 26)) { g_ingest_err = "pcapng: corrupt block"; return PTK_E_ARG; }
            std::vector<unsigned char>& body = p->rec;
            body.resize(total - 8);
            if (fread(body.data(), 1, total - 8, p->f) != total - 8) return 0;
            auto u16 = [&](size_t o) { uint16_t v; memcpy(&v, body.data() + o, 2); return p->swap ? (uint16_t)((v >> 8) | (v << 8)) : v; };
            auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, body.data() + o, 4); return p->swap ? sw32(v) : v; };
            if (type == 1) {                            // interface description: link type + timestamp resolution option
                ptk_pcap::Iface I{u16(0), 1e-6};
                size_t o = 8;
                while (o + 4 <= body.size() - 4) {
                    const uint16_t code = u16(o), olen = u16(o + 2);
                    if (code == 0) break;
                    if (code == 9 && olen >= 1) {       // if_tsresol
                        const unsigned char r = body[o + 4];
                        I.tick = (r & 0x80) ? 1.0 / (double)(1ull << (r & 0x7f)) : 1.0;
                        if (!(r & 0x80)) for (int k = 0; k < r; ++k) I.tick /= 10.0;
                    }
                    o += 4 + ((olen + 3u) & ~3u);
                }
                p->ifaces.push_back(I);
                continue;
            }
            if (type != 6) continue;                    // only enhanced packet blocks carry what we want
            if (body.size() < 24) continue;
            const uint32_t ifid = u32(0);
            if (ifid >= p->ifaces.size()) continue;
            const uint64_t ticks = ((uint64_t)u32(4) << 32) | u32(8);
            incl = u32(12);
            if ((size_t)incl + 20 > body.size()) continue;
            t = (double)ticks * p->ifaces[ifid].tick;
            linktype = p->ifaces[ifid].linktype;
            memmove(body.data(), body.data() + 20, incl);
            body.resize(incl);
        }
        const unsigned char* d = p->rec.data();
        int n = (int)incl, off = 0;
        // link layer -> start of the IP header
        if (linktype == 1) {                       // Ethernet (+ VLAN tags)
            if (n < 14) continue;
            int et = (d[12] << 8) | d[13];
            off = 14;
            while ((et == 0x8100 || et == 0x88a8) && n >= off + 4) { et = (d[off + 2] << 8) | d[off + 3]; off += 4; }
            if (et != 0x0800) continue;
        } else if (linktype == 113) {              // Linux cooked capture
            if (n < 16 || ((d[14] << 8) | d[15]) != 0x0800) continue;
            off = 16;
        } else if (linktype == 0) {                // BSD loopback: 4-byte family
            off = 4;
        } else if (linktype == 101 || linktype == 228 || linktype == 12) {   // raw IP
            off = 0;
        } else { g_ingest_err = "pcap: unsupported link type"; return PTK_E_ARG; }
        if (n < off + 20 || (d[off] >> 4) != 4) continue;
        const int ihl = (d[off] & 15) * 4;
        const int tot = std::min((d[off + 2] << 8) | d[off + 3], n - off);
        if (d[off + 9] != 17 || tot < ihl) continue;
        const int fl = (d[off + 6] << 8) | d[off + 7];
        const bool mf = fl & 0x2000;
        const int foff = (fl & 0x1fff) * 8;
        const unsigned char* pay = d + off + ihl;
        int plen = tot - ihl;
        std::vector<unsigned char>* dgram = nullptr;
        std::vector<unsigned char> whole;
        ptk_pcap::Key key{0, 0, 0};
        if (mf || foff) {                            // a fragment: reassemble
            memcpy(&key.src, d + off + 12, 4); memcpy(&key.dst, d + off + 16, 4);
            key.id = (uint16_t)((d[off + 4] << 8) | d[off + 5]);
            ptk_pcap::Frag& F = p->frags[key];
            if (F.have.empty()) F.seq = p->seq++;
            bool dup = false;
            for (auto& h : F.have) if (h.first == foff) dup = true;
            if (!dup) {
                if ((int)F.data.size() < foff + plen) F.data.resize(foff + plen);
                memcpy(F.data.data() + foff, pay, plen);
                F.have.push_back({foff, plen});
            }
            if (!mf) F.total = foff + plen;
            int got = 0;
            for (auto& h : F.have) got += h.second;
            if (F.total >= 0 && got >= F.total) {
                whole.swap(F.data);
                whole.resize(F.total);
                p->frags.erase(key);
                dgram = &whole;
            } else {
                if (p->frags.size() > 64) {          // forget the oldest incomplete datagram
                    auto old = p->frags.begin();
                    for (auto it = p->frags.begin(); it != p->frags.end(); ++it) if (it->second.seq < old->second.seq) old = it;
                    p->frags.erase(old);
                }
                continue;
            }
            pay = dgram->data(); plen = (int)dgram->size();
        }
        if (plen < 8) continue;
        const int dport = (pay[2] << 8) | pay[3];
        const int ulen = std::min((pay[4] << 8) | pay[5], plen);
        const int body = ulen - 8;
        if (body < 0) continue;
        *len = body;
        if (dst_port) *dst_port = dport;
        if (ts) *ts = t;
        memcpy(buf, pay + 8, std::min(body, cap));
        return 1;
    }
}

// ---- LZ4 frame format (lz4 Frame Format Description 1.6; block format: token / literals / offset / match) -------
extern "C" int ptk_lz4_frame_decompress(const unsigned char* src, unsigned long long n, unsigned char* dst, unsigned long long cap,
                                        unsigned long long* out_len) {
    if (!src || !dst || !out_len) return PTK_E_ARG;
    auto bad = [](const char* m) { g_ingest_err = std::string("lz4: ") + m; return PTK_E_ARG; };
    size_t ip = 0, op = 0;
    if (n < 7 || rd32(src) != 0x184D2204u) return bad("not an LZ4 frame");
    const unsigned flg = src[4];
    if ((flg >> 6) != 1) return bad("unsupported frame version");
    const bool block_checksum = flg & 0x10, content_size = flg & 0x08, content_checksum = flg & 0x04, dict_id = flg & 0x01;
    ip = 6 + (content_size ? 8 : 0) + (dict_id ? 4 : 0) + 1;          // magic, FLG, BD, [size], [dict], header checksum
    if (ip > n) return bad("truncated header");
    while (true) {
        if (ip + 4 > n) return bad("truncated frame");
        const uint32_t bs = rd32(src + ip);
        ip += 4;
        if (bs == 0) break;                                            // end mark
        const size_t len = bs & 0x7fffffffu;
        if (ip + len > n) return bad("truncated block");
        if (bs & 0x80000000u) {                                        // stored block
            if (op + len > cap) { g_ingest_err = "lz4: output buffer too small"; return PTK_E_CAPACITY; }
            memcpy(dst + op, src + ip, len);
            op += len;
        } else {
            size_t b = ip;
            const size_t bend = ip + len;
            while (b < bend) {
                const unsigned token = src[b++];
                size_t lit = token >> 4;
                if (lit == 15) { unsigned c; do { if (b >= bend) return bad("bad literal length"); c = src[b++]; lit += c; } while (c == 255); }
                if (b + lit > bend) return bad("literals run past the block");
                if (op + lit > cap) { g_ingest_err = "lz4: output buffer too small"; return PTK_E_CAPACITY; }
                memcpy(dst + op, src + b, lit);
                op += lit; b += lit;
                if (b >= bend) break;                                  // the last sequence has no match
                if (b + 2 > bend) return bad("truncated offset");
                const size_t offset = src[b] | (src[b + 1] << 8);
                b += 2;
                size_t ml = (token & 15) + 4;
                if ((token & 15) == 15) { unsigned c; do { if (b >= bend) return bad("bad match length"); c = src[b++]; ml += c; } while (c == 255); }
                if (offset == 0 || offset > op) return bad("match offset out of range");
                if (op + ml > cap) { g_ingest_err = "lz4: output buffer too small"; return PTK_E_CAPACITY; }
                for (size_t k = 0; k < ml; ++k) dst[op + k] = dst[op + k - offset];     // overlapping copies are the point
                op += ml;
            }
        }
        ip += len + (block_checksum ? 4 : 0);
    }
    (void)content_checksum;
    *out_len = op;
    return PTK_OK;
}

extern "C" const char* ptk_ingest_last_error(void) { return g_ingest_err.c_str(); }
