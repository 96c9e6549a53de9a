// hybrid.cu -- batched hybrid fusion (dense + BM25) on the device.
//
// Replaces the fusion loop of RetrievalSystem.retrieve_hybrid, src/retrieval.py:174-220:
//   dense_results = retrieve_dense(query, 2k)   -> score 1/(1+d) per hit (:108), ids outside [0, len(chunks)) dropped (:106)
//   bm25_results  = retrieve_bm25(query, 2k)
//   each list divided by its own maximum (0 when the maximum is not positive), weighted 0.6 / 0.4,
//   summed per chunk id in a dict (dense hits inserted first, then BM25-only hits), list sorted by
//   score descending with Python's STABLE sort, cut to k.
// Here both top-2k lists are already on the device (flat scan / sparse scoring kernels), one CTA fuses
// one query, and only the final [nq, k] lists ever leave the GPU.  Arithmetic is float64 (what the
// reference computes under its pinned numpy 1.24: np.float32 distance + Python int -> float64).
#include "common.cuh"
#include "host_common.h"

namespace prs {

constexpr int HY_THREADS = 128;

__device__ __forceinline__ double block_max(double v, double* red, int tid) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double m = red[0];
#pragma unroll
    for (int w = 1; w < HY_THREADS / 32; ++w) m = fmax(m, red[w]);
    __syncthreads();
    return m;
}

__global__ void __launch_bounds__(HY_THREADS) hybrid_fuse_kernel(
    const float* __restrict__ Dd, const long long* __restrict__ Id, int kd, const double* __restrict__ Ss,
    const long long* __restrict__ Is, int ks, long long n_chunks, double wd, double ws, int top_k,
    double* __restrict__ S_out, long long* __restrict__ I_out) {
    extern __shared__ __align__(16) unsigned char hsm[];
    const int n = kd + ks;
    double* sc = reinterpret_cast<double*>(hsm);                 // [n] fused score of slot (dense slots first)
    long long* id = reinterpret_cast<long long*>(sc + n);        // [n] chunk row, -1 = slot not in the union
    double* red = reinterpret_cast<double*>(id + n);             // [HY_THREADS / 32]
    int* s_cnt = reinterpret_cast<int*>(red + HY_THREADS / 32);
    const long long q = blockIdx.x;
    const int tid = threadIdx.x;
    const double NEG = -1.7976931348623157e308;

    // dense hits: similarity 1/(1+d); invalid rows dropped before the maximum, like the reference
    double mx = NEG;
    for (int j = tid; j < kd; j += HY_THREADS) {
        const long long r = Id[q * kd + j];
        const bool ok = r >= 0 && r < n_chunks;
        const double s = ok ? 1.0 / (1.0 + (double)Dd[q * kd + j]) : 0.0;
        id[j] = ok ? r : -1;
        sc[j] = s;
        if (ok) mx = fmax(mx, s);
    }
    const double maxd = block_max(mx, red, tid);
    mx = NEG;
    for (int j = tid; j < ks; j += HY_THREADS) {
        const long long r = Is[q * ks + j];
        const bool ok = r >= 0 && r < n_chunks;
        const double s = ok ? Ss[q * ks + j] : 0.0;
        id[kd + j] = ok ? r : -1;
        sc[kd + j] = s;
        if (ok) mx = fmax(mx, s);
    }
    const double maxs = block_max(mx, red, tid);
    if (tid == 0) *s_cnt = 0;
    // normalise and weight (score / max if max > 0 else 0)
    for (int j = tid; j < n; j += HY_THREADS) {
        const bool dense = j < kd;
        const double m = dense ? maxd : maxs;
        sc[j] = (m > 0.0 ? sc[j] / m : 0.0) * (dense ? wd : ws);
    }
    __syncthreads();
    // union keyed by chunk row: a BM25 hit that is also a dense hit adds to the dense slot
    for (int j = tid; j < ks; j += HY_THREADS) {
        const long long r = id[kd + j];
        if (r < 0) continue;
        for (int i = 0; i < kd; ++i) {
            if (id[i] == r) { sc[i] = sc[i] + sc[kd + j]; id[kd + j] = -1; break; }     // one BM25 hit per row: no race
        }
    }
    __syncthreads();
    // stable descending order = rank by (score desc, insertion position asc)
    for (int j = tid; j < n; j += HY_THREADS) {
        if (id[j] < 0) continue;
        const double s = sc[j];
        int rank = 0;
        for (int i = 0; i < n; ++i) rank += (id[i] >= 0) && (sc[i] > s || (sc[i] == s && i < j));
        atomicAdd(s_cnt, 1);
        if (rank < top_k) { S_out[q * top_k + rank] = s; I_out[q * top_k + rank] = id[j]; }
    }
    __syncthreads();
    for (int j = *s_cnt + tid; j < top_k; j += HY_THREADS) { S_out[q * top_k + j] = 0.0; I_out[q * top_k + j] = -1; }
}

}  // namespace prs

using namespace prs;

extern "C" int prs_hybrid_fuse_device(const float* D_dense, const int64_t* I_dense, int kd, const double* S_sparse,
                                      const int64_t* I_sparse, int ks, int64_t nq, int64_t n_chunks, double dense_weight,
                                      double sparse_weight, int top_k, double* S_out, int64_t* I_out, int device, void* stream) {
    if (kd < 0 || ks < 0 || kd > PRS_MAX_K * 2 || ks > PRS_MAX_K * 2 || top_k < 1 || nq < 0) { set_error("hybrid_fuse: bad arguments"); return PRS_EINVAL; }
    if (nq == 0) return 0;
    if ((kd > 0 && (!D_dense || !I_dense)) || (ks > 0 && (!S_sparse || !I_sparse)) || !S_out || !I_out) { set_error("hybrid_fuse: null pointer"); return PRS_EINVAL; }
    DeviceGuard g(device);
    const size_t smem = (size_t)(kd + ks) * 16 + (HY_THREADS / 32) * 8 + 16;
    PRS_CUDA(cudaFuncSetAttribute(hybrid_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hybrid_fuse_kernel<<<(unsigned)nq, HY_THREADS, smem, (cudaStream_t)stream>>>(D_dense, (const long long*)I_dense, kd, S_sparse,
                                                                                (const long long*)I_sparse, ks, (long long)n_chunks,
                                                                                dense_weight, sparse_weight, top_k, S_out, (long long*)I_out);
    PRS_LAUNCH_CHECK();
    return 0;
}
