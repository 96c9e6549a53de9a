// topk_merge.cuh -- final merge of the per-CTA candidate lists (one CTA per query).
#pragma once
#include "common.cuh"

namespace prs {

// ---------------------------------------------------------------------------------------------
// Final merge over CTAs ("parts"): one CTA per query.
// out_mode 0: D = score (IP)   1: D = -score (direct L2)   2: D = max(0, qnorm - score) (expanded L2)
// ---------------------------------------------------------------------------------------------
constexpr int MERGE_THREADS = 256;
constexpr int MERGE_ONESHOT = 4096;     // inputs up to this many keys are sorted in shared memory in one shot

// One warp sorts buf[0, cnt) (cnt <= 128, unordered, in shared memory) descending in registers and writes back
// max(cnt rounded up, k) slots (missing ones as 0 = empty).  Called by the 32 lanes of one warp; the caller
// synchronises the CTA before and after.
template <int NPL>
__device__ __forceinline__ void warp_sort_buf_npl(u64* buf, int cnt, int k, int lane) {
    u64 v[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) { const int e = lane * NPL + i; v[i] = e < cnt ? buf[e] : 0ull; }
    __syncwarp();
    warp_sort_desc<NPL>(v, lane);
#pragma unroll
    for (int i = 0; i < NPL; ++i) buf[lane * NPL + i] = v[i];
    for (int e = 32 * NPL + lane; e < k; e += 32) buf[e] = 0ull;
}
__device__ __forceinline__ void warp_sort_buf_desc(u64* buf, int cnt, int k, int lane) {
    if (cnt <= 32) warp_sort_buf_npl<1>(buf, cnt, k, lane);
    else if (cnt <= 64) warp_sort_buf_npl<2>(buf, cnt, k, lane);
    else warp_sort_buf_npl<4>(buf, cnt, k, lane);
}

// CTA-level top-k over `parts` lists of L keys each (list p = keys [p*L, (p+1)*L), sorted
// descending, 0 = empty).  buf: sortn keys of shared memory (power of two, >= k + MERGE_THREADS);
// heads: NT keys of shared memory; s_n: two ints.
// Inputs that fit (parts*k <= sortn) are done in one memory round trip:
//   1. every thread fetches its keys (independent loads) and the list heads go to `heads`;
//   2. bound = k-th largest head (rank counting, no sort): k distinct candidates are >= bound, so
//      nothing below it can be in the result;
//   3. the few survivors (>= bound) are compacted into buf and sorted (typically 16-64 keys, i.e.
//      10-21 bitonic stages instead of 66 for the whole input).
// Measured on B200 (148 lists x 10): 28 us for the full in-smem sort -> see profiles/.
// Larger inputs stream through block_topk_stream.  On return buf[0..k) holds the result
// (descending, 0 = empty); returns the number of valid entries.
// NT = threads of the calling CTA (all of them call): 256 in the merge kernels, 192 in the fused tail of the scan kernel.
// STREAM = false drops the streaming fall-back (callers that guarantee parts * L <= min(sortn, MERGE_ONESHOT): the fused
// tail of the scan kernel, whose code size must stay small).
// ONESHOT bounds parts * L of such a caller (fewer unrolled fetches).
template <int NT, bool STREAM = true, int ONESHOT = MERGE_ONESHOT, class Fetch>
__device__ __forceinline__ int block_topk_lists(Fetch fetch, int parts, int L, int k, u64* buf, int sortn, u64* heads, int* s_n, int tid) {
    const long long total = (long long)parts * L;
    if (STREAM) { if (total > sortn || total > ONESHOT) return block_topk_stream(fetch, total, k, buf, sortn, s_n, tid, NT, 1); }
    constexpr int NREG = (ONESHOT + NT - 1) / NT;              // 16 in the merge kernels
    u64 v[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        const int idx = i * NT + tid;
        v[i] = idx < (int)total ? fetch((long long)idx) : 0ull;
    }
    const int nheads = parts < NT ? parts : NT;
    const u64 myhead = tid < nheads ? fetch((long long)tid * L) : 0ull;
    heads[tid] = myhead;
    if (tid == 0) { s_n[0] = 0; s_n[1] = 0; }
    __syncthreads();
    if (myhead) {
        int rank = 0;
        for (int u = 0; u < nheads; ++u) rank += heads[u] > myhead;
        if (rank == k - 1) { s_n[1] = 1; buf[sortn - 1] = myhead; }   // publish the bound (read back before buf is refilled)
    }
    __syncthreads();
    const u64 bound = s_n[1] ? buf[sortn - 1] : 1ull;          // fewer than k heads: keep every non-empty key
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NREG; ++i) {
        if (v[i] >= bound) buf[atomicAdd(&s_n[0], 1)] = v[i];   // count <= total <= sortn
    }
    __syncthreads();
    const int cnt = s_n[0];
    if (cnt <= 128) {
        // the usual case (k .. 3k survivors): ONE warp sorts them in registers -- no block-wide barrier per bitonic stage
        if (tid < 32) warp_sort_buf_desc(buf, cnt, k, tid);
        named_bar_sync(1, NT);
        return cnt < k ? cnt : k;
    }
    int n2 = 2;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + tid; i < n2; i += NT) buf[i] = 0ull;
    if (n2 < k) for (int i = n2 + tid; i < k; i += NT) buf[i] = 0ull;
    named_bar_sync(1, NT);
    block_sort_desc(buf, n2, tid, NT, 1);
    return cnt < k ? cnt : k;
}

// ---------------------------------------------------------------------------------------------
// 16-bit storage, squared L2: exact values for the selected rows.
// The tcgen05 scan ranks rows by the EXPANDED form 2 q.x - ||x||^2 accumulated in fp32; its
// cancellation error (~1e-7 * ||x||^2) is far below the distance of ordinary neighbours but not of
// near-duplicate ones (the reference's fine-tuned indices: ||x||^2 ~ 25, nearest distances down to
// 5e-5, SURVEY.md findings 2/7).  The merge therefore recomputes the squared distance of the k
// selected rows in the DIRECT form sum (q~ - x)^2 -- what faiss does for nq < 20 -- from the stored
// rows and the query rounded to the storage type, and orders the result by (distance, id).
// ---------------------------------------------------------------------------------------------
struct Rerank {
    const unsigned char* x;   // corpus, T64 layout (nullptr = no re-rank)
    const uint16_t* qlow;     // [nq padded][pitch] 16-bit queries in TMEM-slot order (prep kernel), or nullptr:
    const void* q;            // ... then the ORIGINAL queries [nq, d] of type qdtype, rounded to the storage type on the fly
    int qdtype, d;
    int pitch, is_bf16;
};

// slot row of query q in qlow (inverse of the lane spreading in the scan kernel's epilogue)
__device__ __forceinline__ long long qlow_row(int q) {
    const int qi = q & 127;
    return (long long)(q & ~127) + (long long)((qi & 3) * 32 + (qi >> 2));
}

__device__ __forceinline__ void cvt8(const uint4& v, int is_bf16, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (is_bf16) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
        else { const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
}
// one query element as the scan kernel sees it: rounded to the storage type, back in fp32
__device__ __forceinline__ float round_to_storage(float v, int is_bf16) {
    return is_bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : __half2float(__float2half_rn(v));
}
__device__ __forceinline__ float load_query_elem(const void* q, int qdtype, size_t i) {
    if (qdtype == PRS_F32) return reinterpret_cast<const float*>(q)[i];
    if (qdtype == PRS_F16) return __half2float(reinterpret_cast<const __half*>(q)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(q)[i]);
}

// dd[j] = sum (q~ - x_row(j))^2 for the n keys in buf (one warp per key, 16-byte pieces per lane)
template <int NT>
__device__ __forceinline__ void direct_l2_of_keys(const u64* buf, int n, const Rerank& rr, int q, float* dd, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const uint16_t* qrow = rr.qlow ? rr.qlow + (size_t)qlow_row(q) * rr.pitch : nullptr;
    for (int j = warp; j < n; j += NT / 32) {
        const long long row = (long long)key_id<PRS_TIE_LOW_ID>(buf[j]);
        float acc = 0.f;
        for (int ch = lane; ch < (rr.pitch >> 3); ch += 32) {
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(rr.x + t64_offset(row, ch, rr.pitch)));
            float xf[8], qf[8];
            cvt8(xv, rr.is_bf16, xf);
            if (qrow) {
                const uint4 qv = __ldg(reinterpret_cast<const uint4*>(qrow + ch * 8));
                cvt8(qv, rr.is_bf16, qf);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = ch * 8 + e;
                    qf[e] = c < rr.d ? round_to_storage(load_query_elem(rr.q, rr.qdtype, (size_t)q * rr.d + c), rr.is_bf16) : 0.f;
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { const float t = qf[e] - xf[e]; acc = fmaf(t, t, acc); }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) dd[j] = (acc == acc) ? acc : INFINITY;          // NaN rows sort last, ordered by id
    }
}

// shared memory one merge needs (buf | heads | s_n | dd | ids)
__host__ __device__ inline size_t merge_smem_bytes(int sortn, int nt, int k) { return (size_t)sortn * 8 + (size_t)nt * 8 + 16 + (size_t)(k + 2) * 12; }

// Merge of ONE query (all NT threads of the CTA call it): per-part candidate lists -> D / I rows of query q.
template <int NT, bool STREAM = true, int ONESHOT = MERGE_ONESHOT>
__device__ __forceinline__ void merge_query(const u64* __restrict__ cand, int parts, int nq, int k, int sortn, int out_mode,
                                            const float* __restrict__ qnorm, long long id_offset, const Rerank& rr,
                                            float* __restrict__ D, long long* __restrict__ I, int q, int tid, unsigned char* msm) {
    u64* buf = reinterpret_cast<u64*>(msm);
    u64* heads = buf + sortn;
    int* s_n = reinterpret_cast<int*>(heads + NT);
    float* dd = reinterpret_cast<float*>(s_n + 4);               // [k] direct-form distances (re-rank only)
    // every (part, query) list has all k slots written, empty ones as key 0 (scan kernels' contract)
    auto fetch = [&](long long i) -> u64 {
        const int part = (int)((unsigned)i / (unsigned)k), j = (int)i - part * k;
        return __ldcg(cand + ((size_t)part * nq + q) * k + j);
    };
    const int n = block_topk_lists<NT, STREAM, ONESHOT>(fetch, parts, k, k, buf, sortn, heads, s_n, tid);
    if (out_mode == 2 && rr.x) {
        __syncthreads();
        direct_l2_of_keys<NT>(buf, n, rr, q, dd, tid);
        __syncthreads();
        // order by (distance asc, id asc): rank counting (n <= 1024; ids are unique)
        for (int j = tid; j < k; j += NT) {
            if (j < n) {
                const float dj = dd[j];
                const uint32_t idj = key_id<PRS_TIE_LOW_ID>(buf[j]);
                int rank = 0;
                for (int i = 0; i < n; ++i) {
                    const float di = dd[i];
                    rank += (di < dj) || (di == dj && key_id<PRS_TIE_LOW_ID>(buf[i]) < idj);
                }
                D[(size_t)q * k + rank] = dj;
                I[(size_t)q * k + rank] = (long long)idj + id_offset;
            } else {
                D[(size_t)q * k + j] = 3.402823466e+38f;
                I[(size_t)q * k + j] = -1;
            }
        }
        return;
    }
    for (int j = tid; j < k; j += NT) {
        float dv;
        long long iv;
        if (j < n) {
            const u64 key = buf[j];
            const float s = key_score(key);
            dv = out_mode == 0 ? s : (out_mode == 1 ? -s : fmaxf(0.f, qnorm[q] - s));
            iv = (long long)key_id<PRS_TIE_LOW_ID>(key) + id_offset;
        } else {
            dv = out_mode == 0 ? -3.402823466e+38f : 3.402823466e+38f;
            iv = -1;
        }
        D[(size_t)q * k + j] = dv;
        I[(size_t)q * k + j] = iv;
    }
}

__global__ void __launch_bounds__(MERGE_THREADS) merge_cand_kernel(
    const u64* __restrict__ cand, int parts, int nq, int k, int sortn,
    int out_mode, const float* __restrict__ qnorm, long long id_offset, const Rerank rr, float* __restrict__ D,
    long long* __restrict__ I, uint32_t* __restrict__ zero, long long zero_words) {
    extern __shared__ __align__(16) unsigned char msm[];
    // the scan kernel's bootstrap words, cleared for the next search when no preparation kernel will do it
    for (long long i = (long long)blockIdx.x * MERGE_THREADS + threadIdx.x; i < zero_words; i += (long long)gridDim.x * MERGE_THREADS) zero[i] = 0u;
    merge_query<MERGE_THREADS>(cand, parts, nq, k, sortn, out_mode, qnorm, id_offset, rr, D, I, (int)blockIdx.x, (int)threadIdx.x, msm);
}

// ---------------------------------------------------------------------------------------------
// Wide-k path (16 < k <= 1024 on the tcgen05 scan).
// tau: the sampling pass left the best score of every sampled tile per query (gmax[q][M]); the
// k-th largest of them belongs to k distinct rows, so it is a lower bound of the query's true
// k-th best -- the admission threshold of the collecting pass.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MERGE_THREADS) tau_from_maxima_kernel(const float* __restrict__ gmax, long long M, int k, int sortn,
                                                                        float* __restrict__ tau) {
    extern __shared__ __align__(16) unsigned char msm[];
    u64* buf = reinterpret_cast<u64*>(msm);
    int* s_n = reinterpret_cast<int*>(buf + sortn);
    const int q = blockIdx.x, tid = threadIdx.x;
    const float* row = gmax + (size_t)q * M;
    auto fetch = [&](long long i) -> u64 { return ((u64)f2ord(__ldcg(row + i)) << 32) | (u64)(uint32_t)i; };
    const int n = block_topk_stream(fetch, M, k, buf, sortn, s_n, tid, MERGE_THREADS, 1);
    if (tid == 0) tau[q] = (n >= k) ? key_score(buf[k - 1]) : -INFINITY;
}

// select: the collecting pass left, per (part, query), an unsorted slice of keys >= tau[q]; one CTA
// per query streams all slices through the block top-k and writes ONE sorted list of k keys
// (0 = empty), which the ordinary merge / merge+exchange kernel then finishes as a single part.
__global__ void __launch_bounds__(MERGE_THREADS) select_collected_kernel(const u64* __restrict__ coll, const int* __restrict__ coll_cnt,
                                                                         int parts, int nq, int cap, int k, int sortn,
                                                                         u64* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char msm[];
    u64* buf = reinterpret_cast<u64*>(msm);
    int* offs = reinterpret_cast<int*>(buf + sortn);             // [parts + 1]
    int* s_n = offs + parts + 1;
    const int q = blockIdx.x, tid = threadIdx.x;
    for (int p = tid; p < parts; p += MERGE_THREADS) offs[p + 1] = __ldcg(coll_cnt + (size_t)p * nq + q);
    if (tid == 0) offs[0] = 0;
    __syncthreads();
    if (tid == 0) for (int p = 0; p < parts; ++p) offs[p + 1] += offs[p];
    __syncthreads();
    const long long total = offs[parts];
    auto fetch = [&](long long i) -> u64 {
        int lo = 0, hi = parts;                                  // last part with offs[part] <= i
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= (int)i) lo = mid; else hi = mid; }
        return __ldcg(coll + ((size_t)lo * nq + q) * (size_t)cap + ((int)i - offs[lo]));
    };
    const int n = block_topk_stream(fetch, total, k, buf, sortn, s_n, tid, MERGE_THREADS, 1);
    for (int j = tid; j < k; j += MERGE_THREADS) out[(size_t)q * k + j] = (j < n) ? buf[j] : 0ull;
}

}  // namespace prs
