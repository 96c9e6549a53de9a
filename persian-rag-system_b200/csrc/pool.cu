// pool.cu -- encoder output epilogue: attention-masked mean pooling (+ L2 normalisation) fused
// in one kernel.
//
// Replaces the tail of SentenceTransformer.encode at src/retrieval.py:98 and
// src/create_embeddings.py:97-101 (sentence-transformers Pooling(mean) then Normalize):
//   out[b,:] = sum_t hidden[b,t,:] * mask[b,t] / max(sum_t mask[b,t], 1e-9)
//   if normalize: out[b,:] /= max(||out[b,:]||_2, 1e-12)
// Work split: a thread-block CLUSTER per sequence, one CTA per 128-column chunk of H (H = 768 -> 6
// CTAs, 384 -> 3), so B = 32 sequences already fill the machine.  Inside a CTA each of the 8 warps
// takes every 8th token and a lane owns 4 consecutive columns (8- or 16-byte loads, 8 rows in
// flight), partial sums meet in shared memory; the squared norm is reduced ACROSS the cluster's
// CTAs through distributed shared memory (no workspace, no second kernel, no atomics).  [T, H] is
// streamed from HBM exactly once (bytes = B*T*H*sizeof(h) + B*T*8 + B*H*4), fp32 accumulate.  The
// result stays on the device so it can be handed straight to prs_index_search_device.
// Shapes the cluster path does not cover (H % 4 != 0 or H > 1024) use the one-CTA-per-sequence kernel.
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
constexpr int POOL_CHUNK = 128;       // columns per CTA
constexpr int POOL_CPL = 4;           // columns per lane
template <typename T> struct PoolVec;
template <> struct PoolVec<float> {
    __device__ static __forceinline__ void ld(const float* p, float (&f)[4]) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
};
template <> struct PoolVec<__half> {
    __device__ static __forceinline__ void ld(const __half* p, float (&f)[4]) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    }
};
template <> struct PoolVec<__nv_bfloat16> {
    __device__ static __forceinline__ void ld(const __nv_bfloat16* p, float (&f)[4]) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
        f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xFFFF0000u);
        f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xFFFF0000u);
    }
};

template <typename T>
__global__ void __launch_bounds__(POOL_THREADS) pool_norm_cluster_kernel(const T* __restrict__ hidden, const long long* __restrict__ mask,
                                                                         int T_len, int H, int normalize, float* __restrict__ out) {
    __shared__ float s_acc[POOL_THREADS / 32][POOL_CHUNK];
    __shared__ float s_cnt[POOL_THREADS / 32];
    __shared__ float s_sq;                       // this CTA's share of the squared norm (read by the cluster)
    cg::cluster_group cluster = cg::this_cluster();
    const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int col = chunk * POOL_CHUNK + lane * POOL_CPL;
    const bool live = col < H;                   // H % 4 == 0: a lane is entirely inside or outside
    const T* hb = hidden + (size_t)b * T_len * H + col;
    const long long* mb = mask + (size_t)b * T_len;
    float acc[POOL_CPL] = {0.f, 0.f, 0.f, 0.f};
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
            else { x[u][0] = x[u][1] = x[u][2] = x[u][3] = 0.f; }
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
    if (H % POOL_CPL == 0 && H <= 8 * POOL_CHUNK && dtype >= PRS_F32 && dtype <= PRS_BF16) {
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
