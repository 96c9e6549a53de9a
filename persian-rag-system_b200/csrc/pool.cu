// pool.cu -- encoder output epilogue: attention-masked mean pooling (+ L2 normalisation) fused
// in one kernel.
//
// Replaces the tail of SentenceTransformer.encode at src/retrieval.py:98 and
// src/create_embeddings.py:97-101 (sentence-transformers Pooling(mean) then Normalize):
//   out[b,:] = sum_t hidden[b,t,:] * mask[b,t] / max(sum_t mask[b,t], 1e-9)
//   if normalize: out[b,:] /= max(||out[b,:]||_2, 1e-12)
// One CTA per sequence; threads stride the hidden dimension (coalesced), the token loop streams
// [T, H] once from HBM (bytes = B*T*H*sizeof(h) + B*T*8 + B*H*4), fp32 accumulate.  The result
// stays on the device so it can be handed straight to prs_index_search_device.
#include "common.cuh"
#include "host_common.h"

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

}  // namespace prs

using namespace prs;

extern "C" int prs_pool_norm(const void* hidden, int dtype, const int64_t* mask, int B, int T, int H, int normalize, float* out,
                             int device, void* stream) {
    if (B < 0 || T < 0 || H < 1) { set_error("pool_norm: bad shape"); return PRS_EINVAL; }
    if (H > POOL_THREADS * POOL_MAXV) { set_error("pool_norm: H=%d > %d not supported", H, POOL_THREADS * POOL_MAXV); return PRS_EUNSUP; }
    if (B == 0) return 0;
    if (!hidden || !mask || !out) { set_error("pool_norm: null pointer"); return PRS_EINVAL; }
    int arch = prs_device_arch(device);
    if (arch < 0) return arch;
    DeviceGuard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case PRS_F32: pool_norm_kernel<float><<<B, POOL_THREADS, 0, st>>>((const float*)hidden, (const long long*)mask, T, H, normalize, out); break;
        case PRS_F16: pool_norm_kernel<__half><<<B, POOL_THREADS, 0, st>>>((const __half*)hidden, (const long long*)mask, T, H, normalize, out); break;
        case PRS_BF16: pool_norm_kernel<__nv_bfloat16><<<B, POOL_THREADS, 0, st>>>((const __nv_bfloat16*)hidden, (const long long*)mask, T, H, normalize, out); break;
        default: set_error("pool_norm: unsupported dtype %d", dtype); return PRS_EINVAL;
    }
    PRS_LAUNCH_CHECK();
    return 0;
}
