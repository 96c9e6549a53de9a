// pool.cu -- encoder output epilogue: attention-masked mean pooling (+ L2 normalisation) fused
// in one kernel.
//
// Replaces the tail of SentenceTransformer.encode at src/retrieval.py:98 and
// src/create_embeddings.py:97-101 (sentence-transformers Pooling(mean) then Normalize):
//   out[b,:] = sum_t hidden[b,t,:] * mask[b,t] / max(sum_t mask[b,t], 1e-9)
//   if normalize: out[b,:] /= max(||out[b,:]||_2, 1e-12)
// Work split: a thread-block CLUSTER per sequence, one CTA per chunk of H (128 fp32 or 256 16-bit
// columns: H = 768 fp16 -> 3 CTAs, fp32 -> 6), so B = 32 sequences already fill the machine.  Inside
// a CTA each of the 8 warps takes every 8th token and a lane moves 16 bytes per row (8 rows in
// flight), partial sums meet in shared memory; the squared norm is reduced ACROSS the cluster's
// CTAs through distributed shared memory (no workspace, no second kernel, no atomics).  [T, H] is
// streamed from HBM exactly once (bytes = B*T*H*sizeof(h) + B*T*8 + B*H*4), fp32 accumulate.  The
// result stays on the device so it can be handed straight to prs_index_search_device.
// Shapes the cluster path does not cover (H not a multiple of the lane width, or more than 8 chunks)
// use the one-CTA-per-sequence kernel.
#include <cooperative_groups.h>

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

// ---- cluster kernel: grid (nchunks, B), cluster (nchunks, 1, 1), 256 threads ----
// every lane moves 16 bytes per row: 4 fp32 columns or 8 16-bit columns; a CTA owns 32 lanes' worth
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

template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_norm_cluster_kernel(const T* __restrict__ hidden, const long long* __restrict__ mask,
                                                                         int T_len, int H, int normalize, float* __restrict__ out) {
    constexpr int POOL_CPL = PoolVec<T>::CPL;            // columns per lane
    constexpr int POOL_CHUNK = 32 * POOL_CPL;            // columns per CTA (128 fp32 / 256 16-bit)
    __shared__ float s_acc[POOL_THREADS / 32][POOL_CHUNK];
    __shared__ float s_cnt[POOL_THREADS / 32];
    __shared__ float s_sq;                       // this CTA's share of the squared norm (read by the cluster)
    cg::cluster_group cluster = cg::this_cluster();
    const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col = chunk * POOL_CHUNK + lane * POOL_CPL;
    const bool live = col < H;                   // H % CPL == 0: a lane is entirely inside or outside
    const T* hb = hidden + (size_t)b * T_len * H + col;
    const long long* mb = mask + (size_t)b * T_len;
    float acc[POOL_CPL];
#pragma unroll
    for (int c = 0; c < POOL_CPL; ++c) acc[c] = 0.f;
    float cnt = 0.f;
    constexpr int NW = POOL_THREADS / 32, UNR = 8;     // rows in flight per lane (16 was measured slower for 16-bit: registers)
    for (int t0 = warp; t0 < T_len; t0 += NW * UNR) {
        float m[UNR];
        float x[UNR][POOL_CPL];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int t = t0 + u * NW;
            m[u] = t < T_len ? (float)__ldg(mb + t) : 0.f;          // sentence-transformers multiplies by the float mask
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int t = t0 + u * NW;
            if (live && t < T_len && m[u] != 0.f) PoolVec<T>::ld(hb + (size_t)t * H, x[u]);
            else {
#pragma unroll
                for (int c = 0; c < POOL_CPL; ++c) x[u][c] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            cnt += m[u];
#pragma unroll
            for (int c = 0; c < POOL_CPL; ++c) acc[c] = fmaf(x[u][c], m[u], acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < POOL_CPL; ++c) s_acc[warp][lane * POOL_CPL + c] = acc[c];
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    float mean = 0.f, sq = 0.f;
    const int mycol = chunk * POOL_CHUNK + tid;
    if (tid < POOL_CHUNK) {
        float tot = 0.f, n = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { tot += s_acc[w][tid]; n += s_cnt[w]; }
        mean = tot / fmaxf(n, 1e-9f);
        if (mycol < H) sq = mean * mean;
    }
    if (normalize) {
        // CTA share of ||mean||^2, then the cluster total through distributed shared memory
#pragma unroll
        for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        __syncthreads();
        if (lane == 0) s_cnt[warp] = sq;          // s_cnt is free again
        __syncthreads();
        if (tid == 0) s_sq = ((s_cnt[0] + s_cnt[1]) + (s_cnt[2] + s_cnt[3])) + ((s_cnt[4] + s_cnt[5]) + (s_cnt[6] + s_cnt[7]));
        cluster.sync();
        float total = 0.f;
        for (unsigned r = 0; r < cluster.num_blocks(); ++r) total += *cluster.map_shared_rank(&s_sq, r);
        cluster.sync();                           // nobody leaves while a peer still reads its s_sq
        mean = mean / fmaxf(sqrtf(total), 1e-12f);
    }
    if (tid < POOL_CHUNK && mycol < H) out[(size_t)b * H + mycol] = mean;
}

}  // namespace prs

using namespace prs;

template <typename T>
static cudaError_t launch_pool_cluster(const void* hidden, const int64_t* mask, int B, int T_len, int H, int normalize, float* out, cudaStream_t st) {
    constexpr int POOL_CHUNK = 32 * PoolVec<T>::CPL;
    const int nchunks = (H + POOL_CHUNK - 1) / POOL_CHUNK;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nchunks, (unsigned)B, 1);
    cfg.blockDim = dim3(POOL_THREADS, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nchunks; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, pool_norm_cluster_kernel<T>, (const T*)hidden, (const long long*)mask, T_len, H, normalize, out);
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
    if (H % cpl == 0 && H <= 8 * 32 * cpl && dtype >= PRS_F32 && dtype <= PRS_BF16) {
        cudaError_t e = dtype == PRS_F32 ? launch_pool_cluster<float>(hidden, mask, B, T, H, normalize, out, st)
                        : dtype == PRS_F16 ? launch_pool_cluster<__half>(hidden, mask, B, T, H, normalize, out, st)
                                           : launch_pool_cluster<__nv_bfloat16>(hidden, mask, B, T, H, normalize, out, st);
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
