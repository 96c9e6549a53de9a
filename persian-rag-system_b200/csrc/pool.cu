// pool.cu -- encoder output epilogue: attention-masked mean pooling (+ L2 normalisation) fused
// in one kernel.
//
// Replaces the tail of SentenceTransformer.encode at src/retrieval.py:98 and
// src/create_embeddings.py:97-101 (sentence-transformers Pooling(mean) then Normalize):
//   out[b,:] = sum_t hidden[b,t,:] * mask[b,t] / max(sum_t mask[b,t], 1e-9)
//   if normalize: out[b,:] /= max(||out[b,:]||_2, 1e-12)
// Work split: a thread-block CLUSTER per sequence, the CTAs of a cluster take consecutive TOKEN
// slices (whole rows: a slice is one contiguous piece of HBM), so B = 256 sequences of 512 tokens
// become 2 048 short CTAs that balance over the 148 SMs.  A CTA first puts its slice of the mask
// into shared memory (one load latency instead of one per row batch), then every thread streams a
// 16-byte column of R rows at a time, 8 rows in flight, skipping masked rows.  The per-CTA partial
// sums meet ACROSS the cluster through distributed shared memory: CTA j adds up columns
// [j*H/S, (j+1)*H/S) of all S partials in rank order, the squared norm takes one more exchange
// (no workspace, no second kernel, no atomics, deterministic).  [T, H] is streamed from HBM exactly
// once (bytes = B*T*H*sizeof(h) + B*T*8 + B*H*4), fp32 accumulate.  The result stays on the device
// so it can be handed straight to prs_index_search_device.
// Shapes the cluster path does not cover (H not a multiple of the lane width, more than 256
// 16-byte lanes per row, slices above 8 192 tokens) use the one-CTA-per-sequence kernel.
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"
#include "host_common.h"

namespace cg = cooperative_groups;

namespace prs {

constexpr int POOL_THREADS = 256;
constexpr int POOL_MAXV = 8;          // H <= POOL_THREADS * POOL_MAXV

template <typename T> __device__ __forceinline__ float pool_ld(const T* p);
template <> __device__ __forceinline__ float pool_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float pool_ld<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float pool_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_norm_kernel(const T* __restrict__ hidden, const long long* __restrict__ mask,
                                                                 int T_len, int H, int normalize, float* __restrict__ out) {
    __shared__ float s_red[POOL_THREADS / 32];
    __shared__ float s_tot;
    const int b = blockIdx.x, tid = threadIdx.x;
    const T* hb = hidden + (size_t)b * T_len * H;
    const long long* mb = mask + (size_t)b * T_len;
    float acc[POOL_MAXV];
#pragma unroll
    for (int i = 0; i < POOL_MAXV; ++i) acc[i] = 0.f;
    float cnt = 0.f;
    for (int t = 0; t < T_len; ++t) {
        const float m = (float)mb[t];            // sentence-transformers multiplies by the float mask
        cnt += m;
        if (m != 0.f) {
            const T* row = hb + (size_t)t * H;
#pragma unroll
            for (int i = 0; i < POOL_MAXV; ++i) {
                const int h = tid + i * POOL_THREADS;
                if (h < H) acc[i] = fmaf(pool_ld<T>(row + h), m, acc[i]);
            }
        }
    }
    const float denom = fmaxf(cnt, 1e-9f);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < POOL_MAXV; ++i) {
        acc[i] = acc[i] / denom;
        const int h = tid + i * POOL_THREADS;
        if (h < H) sq = fmaf(acc[i], acc[i], sq);
    }
    if (normalize) {
#pragma unroll
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = sq;
        __syncthreads();
        if (tid == 0) {
            float tot = 0.f;
            for (int w = 0; w < POOL_THREADS / 32; ++w) tot += s_red[w];
            s_tot = fmaxf(sqrtf(tot), 1e-12f);
        }
        __syncthreads();
        const float nrm = s_tot;
#pragma unroll
        for (int i = 0; i < POOL_MAXV; ++i) acc[i] = acc[i] / nrm;
    }
#pragma unroll
    for (int i = 0; i < POOL_MAXV; ++i) {
        const int h = tid + i * POOL_THREADS;
        if (h < H) out[(size_t)b * H + h] = acc[i];
    }
}

// ---- cluster kernel: grid (S, B), cluster (S, 1, 1) ----
// every lane moves 16 bytes per row: 4 fp32 columns or 8 16-bit columns
template <typename T> struct PoolVec;
template <> struct PoolVec<float> {
    static constexpr int CPL = 4;
    __device__ static __forceinline__ void ld(const float* p, float (&f)[4]) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
};
template <> struct PoolVec<__half> {
    static constexpr int CPL = 8;
    __device__ static __forceinline__ void ld(const __half* p, float (&f)[8]) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
};
template <> struct PoolVec<__nv_bfloat16> {
    static constexpr int CPL = 8;
    __device__ static __forceinline__ void ld(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    }
};

constexpr int POOL_NST = 3;             // stages of the bulk-copy ring (default; at most POOL_NST_MAX): 36 KB in flight per CTA, 4 CTAs per SM
constexpr int POOL_NST_MAX = 8;
#ifndef POOL_PER
#define POOL_PER 4                      // rows per thread and stage: a stage holds R * POOL_PER rows (<= 16 KB)
#endif

__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

// convert + accumulate one 16-byte lane
template <typename T> __device__ __forceinline__ void pool_fma(const uint4& v, float m, float (&acc)[PoolVec<T>::CPL]);
template <> __device__ __forceinline__ void pool_fma<float>(const uint4& v, float m, float (&acc)[4]) {
    acc[0] = fmaf(__uint_as_float(v.x), m, acc[0]); acc[1] = fmaf(__uint_as_float(v.y), m, acc[1]);
    acc[2] = fmaf(__uint_as_float(v.z), m, acc[2]); acc[3] = fmaf(__uint_as_float(v.w), m, acc[3]);
}
template <> __device__ __forceinline__ void pool_fma<__half>(const uint4& v, float m, float (&acc)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); acc[2 * i] = fmaf(t.x, m, acc[2 * i]); acc[2 * i + 1] = fmaf(t.y, m, acc[2 * i + 1]); }
}
template <> __device__ __forceinline__ void pool_fma<__nv_bfloat16>(const uint4& v, float m, float (&acc)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[2 * i] = fmaf(__uint_as_float(w[i] << 16), m, acc[2 * i]); acc[2 * i + 1] = fmaf(__uint_as_float(w[i] & 0xFFFF0000u), m, acc[2 * i + 1]); }
}

template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_norm_cluster_kernel(const T* __restrict__ hidden, const long long* __restrict__ mask,
                                                                         int T_len, int H, int TS, int R, int NST, int normalize, float* __restrict__ out) {
    constexpr int CPL = PoolVec<T>::CPL;         // columns per lane
    extern __shared__ __align__(128) unsigned char pool_smem[];
    __shared__ uint64_t s_full[POOL_NST_MAX];
    __shared__ float s_red[POOL_THREADS / 32];
    __shared__ int s_hi[POOL_THREADS / 32];
    __shared__ float s_cnt;                      // this CTA's mask total   (read by the cluster)
    __shared__ float s_sq;                       // this CTA's share of the squared norm (read by the cluster)
    __shared__ int s_last;                       // last unmasked row of the slice
    cg::cluster_group cluster = cg::this_cluster();
    const int S = (int)cluster.num_blocks(), j = (int)cluster.block_rank(), b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
    const int VPR = H / CPL;                     // 16-byte lanes per row
    const uint32_t row_bytes = (uint32_t)VPR * 16u;
    const int CH = R * POOL_PER;                 // rows per stage
    const uint32_t stage_bytes = (uint32_t)CH * row_bytes;
    unsigned char* s_stage = pool_smem;                                               // [NST][CH rows]
    float* s_part = reinterpret_cast<float*>(pool_smem + (size_t)NST * stage_bytes);   // [H]    this CTA's partial sums (read by the cluster)
    float* s_acc = s_part + H;                   // [R][H]   per row-group sums
    float* s_m = s_acc + (size_t)R * H;          // [TS]     mask slice as float
    const int t_begin = min(T_len, j * TS), nT = min(T_len, t_begin + TS) - t_begin;
    const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(hidden) + ((size_t)b * T_len + t_begin) * row_bytes;
    // The ring is filled BEFORE the mask is known (the first NST pieces of the slice): the data stream starts together
    // with the mask load instead of one DRAM latency after it.  A slice that turns out to be padding wastes those copies.
    const int spec = min(NST, (nT + CH - 1) / CH);
    if (tid == 0) {
        for (int q = 0; q < NST; ++q) mbar_init(&s_full[q], 1);
        mbar_fence_init();
        for (int q = 0; q < spec; ++q) {
            const uint32_t bytes = (uint32_t)min(CH, nT - q * CH) * row_bytes;
            mbar_arrive_expect_tx(&s_full[q], bytes);
            bulk_g2s(s_stage + (size_t)q * stage_bytes, gsrc + (size_t)(q * CH) * row_bytes, bytes, &s_full[q]);
        }
    }
    // mask slice -> shared memory; its total and the last unmasked row
    const long long* mb = mask + (size_t)b * T_len + t_begin;
    float c_part = 0.f;
    int hi = -1;
    for (int i = tid; i < nT; i += nthr) {
        const float m = (float)__ldg(mb + i);    // sentence-transformers multiplies by the float mask
        s_m[i] = m; c_part += m;
        if (m != 0.f) hi = i;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        c_part += __shfl_xor_sync(0xffffffffu, c_part, o);
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { s_red[warp] = c_part; s_hi[warp] = hi; }
    __syncthreads();
    if (tid == 0) {
        float n = 0.f;
        for (int w = 0; w < (nthr >> 5); ++w) { n += s_red[w]; hi = max(hi, s_hi[w]); }
        s_cnt = n; s_last = hi;
    }
    __syncthreads();
    // rows [0, nrows) of the slice are copied: up to the last unmasked row, at least what is already under way
    // (masked rows in between are multiplied by their 0, like the reference does)
    const int first = 0, nrows = max(s_last + 1, min(nT, spec * CH));
    const int nchunks = (nrows + CH - 1) / CH;
    const int r = tid / VPR, c = tid - r * VPR;
    const bool active = r < R;
    float acc[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
    // thread-constant shared-space addresses: this thread's 16-byte lane of its first row in stage 0, its first mask value
    const uint32_t lane_off = smem_u32(s_stage) + (uint32_t)r * row_bytes + (uint32_t)c * 16u;
    const uint32_t step_bytes = (uint32_t)R * row_bytes;
    uint32_t m_addr = smem_u32(s_m) + (uint32_t)(first + r) * 4u;
    int st = 0;
    uint32_t ph = 0, st_off = 0;
    for (int q = 0; q < nchunks; ++q) {
        mbar_wait(&s_full[st], ph);
        const int count = nrows - q * CH;        // rows of this piece (>= CH except for the last one)
        if (active) {
            float m[POOL_PER];
            uint4 v[POOL_PER];
            if (count >= CH) {
                // every masked row between the first and the last unmasked one is multiplied by its 0 like the reference does
#pragma unroll
                for (int p = 0; p < POOL_PER; ++p) {
                    m[p] = lds_f32(m_addr + (uint32_t)(p * R) * 4u);
                    v[p] = lds_v4(lane_off + st_off + (uint32_t)p * step_bytes);
                }
            } else {
#pragma unroll
                for (int p = 0; p < POOL_PER; ++p) {
                    const bool in = r + p * R < count;          // the rest of the stage holds stale rows
                    m[p] = in ? lds_f32(m_addr + (uint32_t)(p * R) * 4u) : 0.f;
                    v[p] = in ? lds_v4(lane_off + st_off + (uint32_t)p * step_bytes) : make_uint4(0u, 0u, 0u, 0u);
                }
            }
#pragma unroll
            for (int p = 0; p < POOL_PER; ++p) pool_fma<T>(v[p], m[p], acc);
        }
        m_addr += (uint32_t)CH * 4u;
        __syncthreads();                         // everybody is done with the stage: refill it
        if (tid == 0 && q + NST < nchunks) {
            const int qn = q + NST;
            const uint32_t bytes = (uint32_t)min(CH, nrows - qn * CH) * row_bytes;
            mbar_arrive_expect_tx(&s_full[st], bytes);
            bulk_g2s(s_stage + st_off, gsrc + (size_t)(first + qn * CH) * row_bytes, bytes, &s_full[st]);
        }
        st_off += stage_bytes;
        if (++st == NST) { st = 0; st_off = 0; ph ^= 1u; }
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) s_acc[(size_t)r * H + c * CPL + k] = acc[k];
    }
    __syncthreads();
    for (int h = tid; h < H; h += nthr) { float t = 0.f; for (int g = 0; g < R; ++g) t += s_acc[(size_t)g * H + h]; s_part[h] = t; }
    cluster.sync();
    // CTA j finishes columns [h_lo, h_hi): partials added in rank order
    const int HS = (H + S - 1) / S, h_lo = min(H, j * HS), h_hi = min(H, h_lo + HS);
    float n = 0.f;
    for (int g = 0; g < S; ++g) n += *cluster.map_shared_rank(&s_cnt, g);
    const float denom = fmaxf(n, 1e-9f);
    float* s_mean = s_acc;                       // s_acc is free again; HS <= H
    float sq = 0.f;
    for (int h = h_lo + tid; h < h_hi; h += nthr) {
        float t = 0.f;
        for (int g = 0; g < S; ++g) t += cluster.map_shared_rank(s_part, g)[h];
        const float mean = t / denom;
        s_mean[h - h_lo] = mean;
        sq = fmaf(mean, mean, sq);
    }
    float nrm = 1.f;
    if (normalize) {
#pragma unroll
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) s_red[warp] = sq;         // s_red is free again (two barriers since its last reader)
        __syncthreads();
        if (tid == 0) { float t = 0.f; for (int w = 0; w < (nthr >> 5); ++w) t += s_red[w]; s_sq = t; }
        cluster.sync();
        float total = 0.f;
        for (int g = 0; g < S; ++g) total += *cluster.map_shared_rank(&s_sq, g);
        nrm = fmaxf(sqrtf(total), 1e-12f);
    }
    for (int h = h_lo + tid; h < h_hi; h += nthr) out[(size_t)b * H + h] = s_mean[h - h_lo] / nrm;      // written by this very thread above
    cluster.sync();                              // nobody leaves while a peer still reads its shared memory
}

}  // namespace prs

using namespace prs;

// Split of one sequence over a cluster of S token slices: about three CTAs per SM, all resident at once (a CTA streams
// at ring / latency, so long sequences in a ragged batch finish sooner when they are split), but no more: every CTA pays
// the same start-up and exchange, and clusters of 8 fit only 48 at a time.  Slices of at least 16 tokens.
static int pool_cluster_size(int B, int T_len) {
    int S = 1;
    while (S < 8 && (long long)B * S < 148 * 3 && T_len / (2 * S) >= 16) S *= 2;
#ifdef PRS_EXPERIMENTS
    if (const char* e = getenv("PRS_POOL_S")) { const int v = atoi(e); if (v >= 1 && v <= 8) S = v; }
#endif
    return S;
}

template <typename T>
static cudaError_t launch_pool_cluster(const void* hidden, const int64_t* mask, int B, int T_len, int H, int normalize, float* out, int device, cudaStream_t st) {
    const int VPR = H / PoolVec<T>::CPL;                      // <= POOL_THREADS (checked by the caller)
    const int R = POOL_THREADS / VPR;                         // rows a CTA moves at a time
    const int threads = (R * VPR + 31) & ~31;
    const int S = pool_cluster_size(B, T_len);
    const int TS = (T_len + S - 1) / S;
    int NST = POOL_NST;
#ifdef PRS_EXPERIMENTS
    if (const char* e = getenv("PRS_POOL_NST")) { const int v = atoi(e); if (v >= 1 && v <= POOL_NST_MAX) NST = v; }
#endif
    const size_t smem = (size_t)NST * POOL_PER * R * VPR * 16 + ((size_t)H * (1 + R) + (size_t)(TS > 0 ? TS : 1)) * sizeof(float);   // <= 64 + 8 + 8 + 32 KB
    static std::atomic<bool> attr_done[64];
    if (device >= 0 && device < 64 && !attr_done[device].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(pool_norm_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
        if (e != cudaSuccess) return e;
        attr_done[device].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)S, (unsigned)B, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, pool_norm_cluster_kernel<T>, (const T*)hidden, (const long long*)mask, T_len, H, TS, R, NST, normalize, out);
}

extern "C" int prs_pool_norm(const void* hidden, int dtype, const int64_t* mask, int B, int T, int H, int normalize, float* out,
                             int device, void* stream) {
    if (B < 0 || T < 0 || H < 1) { set_error("pool_norm: bad shape"); return PRS_EINVAL; }
    if (H > POOL_THREADS * POOL_MAXV) { set_error("pool_norm: H=%d > %d not supported", H, POOL_THREADS * POOL_MAXV); return PRS_EUNSUP; }
    if (B == 0) return 0;
    if (!hidden || !mask || !out) { set_error("pool_norm: null pointer"); return PRS_EINVAL; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("pool_norm: no CUDA device %d", device); return PRS_ECUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    if (B > 65535) { set_error("pool_norm: B=%d > 65535 sequences per call", B); return PRS_EINVAL; }
    const int cpl = dtype == PRS_F32 ? 4 : 8;      // 16 bytes per lane and row
    const int S = pool_cluster_size(B, T);
    if (H % cpl == 0 && H / cpl <= POOL_THREADS && (T + S - 1) / S <= 8192 && ((uintptr_t)hidden & 15) == 0 && dtype >= PRS_F32 && dtype <= PRS_BF16) {
        cudaError_t e = dtype == PRS_F32 ? launch_pool_cluster<float>(hidden, mask, B, T, H, normalize, out, device, st)
                        : dtype == PRS_F16 ? launch_pool_cluster<__half>(hidden, mask, B, T, H, normalize, out, device, st)
                                           : launch_pool_cluster<__nv_bfloat16>(hidden, mask, B, T, H, normalize, out, device, st);
        if (e != cudaSuccess) { set_error("pool_norm: cluster launch failed: %s", cudaGetErrorString(e)); return PRS_ECUDA; }
        PRS_LAUNCH_CHECK();
        return 0;
    }
    switch (dtype) {
        case PRS_F32: pool_norm_kernel<float><<<B, POOL_THREADS, 0, st>>>((const float*)hidden, (const long long*)mask, T, H, normalize, out); break;
        case PRS_F16: pool_norm_kernel<__half><<<B, POOL_THREADS, 0, st>>>((const __half*)hidden, (const long long*)mask, T, H, normalize, out); break;
        case PRS_BF16: pool_norm_kernel<__nv_bfloat16><<<B, POOL_THREADS, 0, st>>>((const __nv_bfloat16*)hidden, (const long long*)mask, T, H, normalize, out); break;
        default: set_error("pool_norm: unsupported dtype %d", dtype); return PRS_EINVAL;
    }
    PRS_LAUNCH_CHECK();
    return 0;
}
