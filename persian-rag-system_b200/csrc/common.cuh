// common.cuh -- shared device/host helpers for libprs (sm_100a only).
//
// Top-k bookkeeping everywhere in this library works on 64-bit KEYS so that selection is a
// plain integer max and ties are resolved deterministically:
//     key = ord(score) << 32 | tie(id)         larger key == better candidate
// ord() maps a float to a uint32 that sorts like the float; `score` is always "larger is
// better" (inner product, or minus the squared L2 distance); tie(id) = ~id when the lower id
// must win (dense path, faiss order) or id when the higher id must win (sparse path, the
// order of np.argsort(scores, kind="stable")[::-1], SURVEY.md finding 6).  key 0 = empty slot.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/prs.h"

typedef unsigned long long u64;

#define PRS_TIE_LOW_ID  0
#define PRS_TIE_HIGH_ID 1

namespace prs {

__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t b = o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
// NaN never wins (faiss: comparisons with NaN are false, so NaN never enters the heap)
__device__ __forceinline__ float sanitize(float s) { return (s == s) ? s : -INFINITY; }

template <int TIE>
__device__ __forceinline__ u64 make_key(float score, uint32_t id) {
    return ((u64)f2ord(score) << 32) | (u64)(TIE == PRS_TIE_LOW_ID ? ~id : id);
}
__device__ __forceinline__ u64 make_key_rt(float score, uint32_t id, int tie_high) {
    return ((u64)f2ord(score) << 32) | (u64)(tie_high ? id : ~id);
}
__device__ __forceinline__ float key_score(u64 k) { return ord2f((uint32_t)(k >> 32)); }
template <int TIE>
__device__ __forceinline__ uint32_t key_id(u64 k) { uint32_t lo = (uint32_t)k; return TIE == PRS_TIE_LOW_ID ? ~lo : lo; }
__device__ __forceinline__ uint32_t key_id_rt(u64 k, int tie_high) { uint32_t lo = (uint32_t)k; return tie_high ? lo : ~lo; }

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, m);
    hi = __shfl_xor_sync(0xffffffffu, hi, m);
    return ((u64)hi << 32) | lo;
}
__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return ((u64)hi << 32) | lo;
}

// ---------------------------------------------------------------------------------------------
// Warp-level bitonic sort, DESCENDING, of 32*NPL keys held NPL per lane.
// Element e lives in lane e / NPL, register e % NPL, so the small strides stay inside a lane.
// ---------------------------------------------------------------------------------------------
template <int NPL>
__device__ __forceinline__ void warp_sort_desc(u64 (&v)[NPL], int lane) {
    constexpr int N = 32 * NPL;
#pragma unroll
    for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride < NPL) {
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    if ((i & stride) == 0) {
                        const int e = lane * NPL + i;
                        const bool desc = ((e & size) == 0);         // block direction
                        u64 a = v[i], b = v[i + stride];
                        const bool sw = desc ? (a < b) : (a > b);
                        v[i] = sw ? b : a;
                        v[i + stride] = sw ? a : b;
                    }
                }
            } else {
                const int lm = stride / NPL;
#pragma unroll
                for (int i = 0; i < NPL; ++i) {
                    const int e = lane * NPL + i;
                    const bool desc = ((e & size) == 0);
                    const bool lower = ((lane & lm) == 0);           // I hold the lower-index element
                    u64 mine = v[i];
                    u64 other = shfl_xor_u64(mine, lm);
                    // lower index keeps the larger key in a descending block
                    const bool keep_max = (lower == desc);
                    v[i] = keep_max ? (mine > other ? mine : other) : (mine < other ? mine : other);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Block-level bitonic sort, DESCENDING, of n (power of two) keys in shared memory.
// Must be called by all `nthreads` threads; uses bar.sync on barrier `bar_id`.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int bar_id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void block_sort_desc(u64* s, int n, int tid, int nthreads, int bar_id) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int p = tid; p < (n >> 1); p += nthreads) {
                const int i = ((p & ~(stride - 1)) << 1) | (p & (stride - 1));
                const int j = i + stride;
                const bool desc = ((i & size) == 0);
                u64 a = s[i], b = s[j];
                const bool sw = desc ? (a < b) : (a > b);
                if (sw) { s[i] = b; s[j] = a; }
            }
            named_bar_sync(bar_id, nthreads);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming block-level top-k: feeds `total` keys (fetch(i) -> key, 0 = skip) through a
// shared-memory buffer of `sortn` keys (power of two, sortn >= k + nthreads), keeping the k
// largest.  On return buf[0..min(count,k)) holds them sorted descending; returns how many.
// All `nthreads` threads must call it.  s_cnt: one shared int.
// ---------------------------------------------------------------------------------------------
template <class Fetch>
__device__ __forceinline__ int block_topk_stream(Fetch fetch, long long total, int k, u64* buf, int sortn,
                                                 int* s_cnt, int tid, int nthreads, int bar_id) {
    if (tid == 0) *s_cnt = 0;
    named_bar_sync(bar_id, nthreads);
    u64 thr = 0;
    for (long long base = 0; base < total; base += nthreads) {
        const long long i = base + tid;
        u64 key = (i < total) ? fetch(i) : 0ull;
        if (key > thr) {
            const int pos = atomicAdd(s_cnt, 1);
            buf[pos] = key;                       // pos < sortn: count <= sortn - nthreads before the round
        }
        named_bar_sync(bar_id, nthreads);
        const int cnt = *s_cnt;
        if (cnt > sortn - nthreads) {
            for (int p = cnt + tid; p < sortn; p += nthreads) buf[p] = 0ull;
            named_bar_sync(bar_id, nthreads);
            block_sort_desc(buf, sortn, tid, nthreads, bar_id);
            const int keep = cnt < k ? cnt : k;
            thr = (keep == k) ? buf[k - 1] : 0ull;
            if (tid == 0) *s_cnt = keep;
            named_bar_sync(bar_id, nthreads);
        }
    }
    int cnt = *s_cnt;
    named_bar_sync(bar_id, nthreads);
    // final sort over the smallest power of two covering cnt
    int n2 = 2;
    while (n2 < cnt) n2 <<= 1;
    for (int p = cnt + tid; p < n2; p += nthreads) buf[p] = 0ull;
    named_bar_sync(bar_id, nthreads);
    block_sort_desc(buf, n2, tid, nthreads, bar_id);
    return cnt < k ? cnt : k;
}

// ---------------------------------------------------------------------------------------------
// HBM layout of 16-bit corpora ("T64"): blocks of 64 rows, each block contiguous (64 * pitch * 2
// bytes).  Inside a block the data is k-block-major -- [pitch/64 k-blocks][64 rows][128 bytes] --
// and the eight 16-byte chunks of every 128-byte row piece are stored pre-swizzled
// (chunk ^ (row & 7)), i.e. exactly the image a 128-byte-swizzle TMA tensor load would leave in
// shared memory.  One plain bulk copy of a contiguous k-block range is then directly consumable
// by tcgen05.mma (K-major, SWIZZLE_128B descriptor), every byte of a DRAM page is used by the
// same copy, and no tensor map is needed.
// ---------------------------------------------------------------------------------------------
constexpr int BLK_ROWS = 64;
constexpr int KBLOCK_BYTES = BLK_ROWS * 128;       // one k-block (64 elements) of one row block
// byte offset of the 16-byte chunk `chunk` (8 elements) of row `row`; pitch in elements (multiple of 64)
__host__ __device__ __forceinline__ size_t t64_offset(long long row, int chunk, int pitch) {
    const int r = (int)(row & (BLK_ROWS - 1));
    return (size_t)(row >> 6) * ((size_t)BLK_ROWS * pitch * 2) + (size_t)(chunk >> 3) * KBLOCK_BYTES + (size_t)r * 128 +
           (size_t)((((chunk & 7) ^ (r & 7))) << 4);
}

__host__ __device__ inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA) PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace prs
