// flat_umma.cuh -- tcgen05 flat scan with a fused per-query top-k (fp16 / bf16 storage).
//
// Replaces faiss IndexFlat::search (reference call site src/retrieval.py:102) for query batches
// that make the scan a real GEMM while it is still HBM-bound (8 <= nq; 128 queries per pass).
//
//   D[128 queries x TILE_N rows] (+)= A[128 x K] * B[TILE_N x K]^T      (tcgen05.mma, kind::f16)
//
//   A  = the query block, converted to the storage type once, RESIDENT IN TENSOR MEMORY for the
//        whole kernel (tcgen05.st, one TMEM lane per query).  Shared memory therefore holds
//        nothing but the streaming ring, and the MMA never re-reads queries from smem or L2.
//   B  = corpus rows, K-major.  The corpus sits in HBM in the T64 layout (common.cuh): 64-row
//        blocks, k-block-major, pre-swizzled, so each pipeline stage is ONE contiguous bulk copy
//        (cp.async.bulk, SASS UBLKCP) of `kbs` k-blocks (8 KB each) through an mbarrier ring.
//   D  = fp32 accumulators in TMEM, double buffered (2 x TILE_N columns).
//   epilogue (4 warps, one thread per query): tcgen05.ld its lane's TILE_N scores, apply the L2
//        bias (2 q.x - ||x||^2), compare with the thread's k-th best, and on the rare hit insert
//        into a thread-private sorted list held in 16 registers.  The score matrix never exists in
//        HBM; each CTA emits one sorted top-k per query (cand[cta][query][k]).
//
// TMEM budget (512 columns): A uses pitch/2 columns (two 16-bit values per column), D uses
// 2*TILE_N = 128 (TILE_N = 64 = one row block), so pitch <= 768.
//
// Warp roles (192 threads): warp 0 = TMA producer + TMEM allocator, warp 1 = MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#pragma once
#include <algorithm>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "host_common.h"
#include "topk_merge.cuh"
#include "xchg.cuh"

// Timing experiments (profiles/r1_umma_variant_experiments.log) change what the kernel computes, so they
// exist only in builds made with -DPRS_EXPERIMENTS; the shipped library has no such switches.
#ifdef PRS_EXPERIMENTS
#define UMMA_DBG(p) ((p).dbg)
#else
#define UMMA_DBG(p) 0
#endif

namespace prs {

constexpr int UMMA_MAX_K = 16;         // thread-private sorted lists are 16 registers per thread
constexpr int UMMA_THREADS = 192;
constexpr int UMMA_M = 128;            // queries per pass
constexpr int UMMA_MAX_STAGES = 24;
constexpr int UMMA_TMEM_COLS = 512;
constexpr int UMMA_FUSED_ONESHOT = 160 * UMMA_MAX_K;   // candidate keys one merge of the one-launch search can meet (<= 160 CTAs x k <= 16)

struct UmmaParams {
    const unsigned char* x; // corpus, T64 layout
    const uint16_t* qlow;   // [128, pitch] 16-bit queries of this pass, slot ordered (zero padded)
    const float* xnorm;     // [n_rows] squared norms of the stored rows
    long long n_rows;
    int pitch, nq, k, l2, stages, is_bf16, kbs;   // kbs: k-blocks per pipeline stage
    int nbuf;               // accumulator buffers in TMEM: 2 when pitch <= 512, else 1
    int reboot;             // mode 0: republish the CTA's running best and refresh the cross-CTA bound at tiles 1, 3, 15, 63, ...
    int dbg;                // experiments only (PRS_UMMA_DEBUG): 1 no bootstrap+no inserts, 2 epilogue releases without reading, 16 no corpus copies
    int nq_total, q0;
    u64* cand;              // [grid][nq_total][k]
    int* cand_cnt;          // [grid][nq_total]
    int tile_step;          // 1 = every tile; S > 1 = every S-th tile (threshold sampling pass of the wide-k path)
    int mode;               // 0: per-part sorted top-k lists (k <= 16), pruned by the cross-CTA bootstrap bound;
                            // 1: collect every score >= tau[q];  2: best score of every visited tile -> gmax (threshold sampling)
    float* gmax;            // mode 2: [nq_total][gmax_stride] best score per (query, sampled tile)
    long long gmax_stride;
    const float* tau;       // mode 1: [nq_total] admission threshold per query (a lower bound of its k-th best)
    u64* coll;              // mode 1: [parts][nq_total][coll_cap] unsorted keys
    int* coll_cnt;          // mode 1: [parts][nq_total]
    int coll_cap;           // mode 1: entries per (part, query) slice, >= 2k (a full slice is compacted to its k best)
    uint32_t* boot;         // per query block: [parts][128] ord(best score of the first tile) + 1 counter; zeroed per search
    long long boot_stride;  // words between the bootstrap arrays of consecutive query blocks
    // ---- one-launch search (nq <= 128, k <= 16): the query preparation and the merge live in this kernel ----
    const void* q;          // fuse_prep: ORIGINAL queries [nq, d] (qdtype), converted by the epilogue threads themselves
    int qdtype, d, fuse_prep, fuse_merge;
    float* qnorm;           // fuse_prep: [nq] ||q~||^2 out (written by part 0)
    unsigned int* gbar;     // fuse_merge: [0] arrivals, [1] generation of the grid barrier (self-resetting)
    int sortn, out_mode, largest;
    long long id_offset;
    float* D;
    long long* I;
    Rerank rr;
    int* status;            // host-mapped report word (grid-barrier timeout), may be null
    int use_xchg;           // row-sharded search: the tail PUSHES the local lists to the peers (xchg_pull_kernel finishes)
    uint32_t xgen;
    XchgView xv;
};

// ---------------- PTX wrappers (tcgen05 / TMA) ----------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// same with a compile-time accumulate flag and without the optional disable-output-lane operand (fewest instructions per MMA)
template <int ACC>
__device__ __forceinline__ void umma_ts_f16_c(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC) : "memory");
}
// one thread of a converged warp (elect.sync): what single-thread issue loops (TMA, tcgen05.mma) should branch on
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// ---- thread-block cluster helpers (large batches: one corpus stage feeds CL query blocks) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy global -> the same shared-memory offset of every CTA in `mask`; each destination's
// mbarrier (same offset) receives the complete_tx for these bytes
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// MMA completion arrives on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// ---- CTA pair (cta_group::2): the two CTAs of a cluster execute ONE MMA of M = 256 -- each holds its own 128 query rows
// (A) and accumulator rows (D) in its tensor memory and HALF of the corpus tile (B) in its shared memory; the even CTA issues.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int ACC>
__device__ __forceinline__ void umma_ts_f16_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "n"(ACC) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {       // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// three-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start[0,14) LBO[16,30) SBO[32,46) version[46,48)=1 layout[61,64)=2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// k-th largest of n distinct 64-bit keys (n >= k), by bisection on the key bits.  One thread, rare path
// (a collection slice of the wide-k scan filled up); kept out of line so the hot loop stays small.
__device__ __noinline__ u64 slice_kth_largest(const u64* s, int n, int k) {
    u64 prefix = 0ull;
    for (int b = 63; b >= 0; --b) {
        const u64 c = prefix | (1ull << b);
        int cnt = 0;
        for (int i = 0; i < n; ++i) cnt += s[i] >= c;
        if (cnt >= k) prefix = c;
    }
    return prefix;
}

// One-launch search, prologue: an epilogue thread converts ITS query row (fp32 / fp16 / bf16, d a multiple of 8) to
// the storage type on the way into its TMEM lane and returns ||q~||^2 of the rounded row.  Out of line on purpose, and
// with scalar arguments: taking the address of the kernel's parameter block (or indexing its arrays dynamically) makes
// the compiler copy it to local memory and drop the uniform-datapath branches of the hot loops (measured: +15 % scan time).
// NORM = false (inner product: nobody reads ||q~||^2) drops the convert-back + FMA pair per element -- 3 of the 5
// instructions per element pair of a loop that runs in a warp with a scheduler to itself.
template <bool NORM>
__device__ __forceinline__ float stage_query_row_t(const void* q, int qdtype, int d, int is_bf16, uint32_t a_addr, int ncol, bool qvalid, size_t qoff) {
    float qn = 0.f;
    auto pack2 = [&](float a, float b) -> uint32_t {
        if (is_bf16) {
            const __nv_bfloat16 x0 = __float2bfloat16_rn(a), x1 = __float2bfloat16_rn(b);
            if (NORM) {
                const float f0 = __bfloat162float(x0), f1 = __bfloat162float(x1);
                qn = fmaf(f0, f0, qn); qn = fmaf(f1, f1, qn);
            }
            return (uint32_t)__bfloat16_as_ushort(x0) | ((uint32_t)__bfloat16_as_ushort(x1) << 16);
        }
        const __half x0 = __float2half_rn(a), x1 = __float2half_rn(b);
        if (NORM) {
            const float f0 = __half2float(x0), f1 = __half2float(x1);
            qn = fmaf(f0, f0, qn); qn = fmaf(f1, f1, qn);
        }
        return (uint32_t)__half_as_ushort(x0) | ((uint32_t)__half_as_ushort(x1) << 16);
    };
    const int nch = d >> 3;                           // 16-byte chunks (8 elements) that hold data
    if (qdtype == PRS_F32) {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(q) + qoff);
        for (int c0 = 0; c0 < ncol; c0 += 64) {         // 64 columns = 128 elements = 32 float4 loads in flight
            float4 t[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int ch = (c0 >> 2) + (i >> 1);
                t[i] = (qvalid && ch < nch) ? __ldg(src + (size_t)ch * 2 + (i & 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (c0 + g * 32 >= ncol) break;
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[2 * i] = pack2(t[g * 16 + i].x, t[g * 16 + i].y);
                    v[2 * i + 1] = pack2(t[g * 16 + i].z, t[g * 16 + i].w);
                }
                tmem_st32(a_addr + (uint32_t)(c0 + g * 32), v);
            }
        }
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(q) + qoff);
        const int src_bf16 = qdtype == PRS_BF16;
        for (int c0 = 0; c0 < ncol; c0 += 96) {         // 24 loads of 16 bytes in flight
            uint4 t[24];
#pragma unroll
            for (int i = 0; i < 24; ++i) {
                const int ch = (c0 >> 2) + i;
                t[i] = (qvalid && ch < nch && c0 + 4 * i < ncol) ? __ldg(src + ch) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                if (c0 + g * 32 >= ncol) break;
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float f[8];
                    cvt8(t[g * 8 + i], src_bf16, f);
                    v[4 * i] = pack2(f[0], f[1]); v[4 * i + 1] = pack2(f[2], f[3]);
                    v[4 * i + 2] = pack2(f[4], f[5]); v[4 * i + 3] = pack2(f[6], f[7]);
                }
                tmem_st32(a_addr + (uint32_t)(c0 + g * 32), v);
            }
        }
    }
    return qn;
}
__device__ __noinline__ float stage_query_row(const void* q, int qdtype, int d, int is_bf16, uint32_t a_addr, int ncol, bool qvalid, size_t qoff,
                                              int need_norm) {
    return need_norm ? stage_query_row_t<true>(q, qdtype, d, is_bf16, a_addr, ncol, qvalid, qoff)
                     : stage_query_row_t<false>(q, qdtype, d, is_bf16, a_addr, ncol, qvalid, qoff);
}

// The two merges of the one-launch search, out of line with scalar arguments (see stage_query_row): inlined, they
// triple the size of the scan kernel and ptxas then stops using uniform-datapath branches in its hot loops.
__device__ __noinline__ void fused_merge_plain(const u64* cand, int parts, int nq, int k, int sortn, int out_mode, const float* qnorm,
                                               long long id_offset, Rerank rr, float* D, long long* I, int q, int tid, unsigned char* smem) {
    merge_query<UMMA_THREADS, false, UMMA_FUSED_ONESHOT>(cand, parts, nq, k, sortn, out_mode, qnorm, id_offset, rr, D, I, q, tid, smem);
}
// row-sharded variant of the tail: local merge + push to the peers' exchange buffers.  Arguments come through shared
// memory (the peers' pointers are indexed by rank at run time; see stage_query_row for why not from the parameter block).
struct FusedPushArgs {
    XchgView xv;
    Rerank rr;
    const u64* cand;
    const float* qnorm;
    long long id_offset;
    uint32_t gen;
    int parts, nq, k, sortn, out_mode;
};
__device__ __noinline__ void fused_push(const FusedPushArgs* a, int q, int tid, unsigned char* smem) {
    xchg_push_query<UMMA_THREADS, false, UMMA_FUSED_ONESHOT>(a->cand, a->parts, a->nq, a->k, a->sortn, a->out_mode, a->qnorm, a->id_offset, a->rr,
                                                             a->xv, a->gen, q, tid, smem);
}

// One-launch search, tail: grid barrier, then the CTAs merge the queries among themselves.  Row-sharded searches only
// PUSH here (local merge + stores into the peers' buffers); waiting for the peers and the G-way merge are a second, tiny
// kernel: with the whole exchange reachable from this kernel ptxas stops emitting uniform-datapath branches in the hot
// loops (the scan got 15 % slower), and pushing early lets the NVLink stores overlap the peers' tails.
__device__ __forceinline__ void fused_tail(const UmmaParams& p, int part, int nparts, int tid, unsigned char* base, int* s_flag) {
    // ---------------- one-launch search: grid barrier, then the CTAs merge the queries among themselves ----------------
    // The launch is cooperative (all CTAs co-resident).  Every CTA's lists are written (the __syncthreads above);
    // thread 0 publishes them with a fence + arrival and waits for the generation to move.  The barrier resets
    // itself (the last arriver zeroes the count), so nothing has to be prepared per search.  The wait is bounded.
    if (tid == 0) {
        const unsigned g0 = ld_relaxed_gpu(p.gbar + 1);
        __threadfence();
        const unsigned arrived = atomicAdd(p.gbar, 1u);
        int ok = 1;
        if (arrived == gridDim.x - 1) {
            p.gbar[0] = 0u;
            __threadfence();
            atomicAdd(p.gbar + 1, 1u);
        } else {
            unsigned long long t0 = 0;
            int spins = 0;
            while (ld_relaxed_gpu(p.gbar + 1) == g0) {
                __nanosleep(20);
                if ((++spins & 4095) == 0) {
                    const unsigned long long now = global_timer_ns();
                    if (!t0) t0 = now;
                    else if (now - t0 > 4000000000ull) { ok = 0; break; }
                }
            }
        }
        __threadfence();
        *s_flag = ok;
    }
    __syncthreads();
    const bool ok = *s_flag != 0;
    // the bootstrap words this CTA used are cleared for the next search on this workspace
    for (int i = tid; i < UMMA_M; i += UMMA_THREADS) p.boot[(size_t)part * UMMA_M + i] = 0u;
    if (part == 0 && tid == 0) p.boot[(size_t)nparts * UMMA_M] = 0u;
    FusedPushArgs* xa = reinterpret_cast<FusedPushArgs*>(base + 96 * 1024);
    if (p.use_xchg) {
        if (tid == 0) {
#pragma unroll
            for (int r = 0; r < XCHG_MAX_RANKS; ++r) {
                xa->xv.vals[r] = p.xv.vals[r]; xa->xv.vals2[r] = p.xv.vals2[r]; xa->xv.ids[r] = p.xv.ids[r]; xa->xv.flags[r] = p.xv.flags[r];
            }
            xa->xv.cap = p.xv.cap; xa->xv.nq_cap = p.xv.nq_cap; xa->xv.G = p.xv.G; xa->xv.rank = p.xv.rank;
            xa->rr = p.rr; xa->cand = p.cand; xa->qnorm = p.qnorm; xa->id_offset = p.id_offset; xa->gen = p.xgen;
            xa->parts = nparts; xa->nq = p.nq_total; xa->k = p.k; xa->sortn = p.sortn; xa->out_mode = p.out_mode;
        }
        __syncthreads();
    }
    for (int q = part; q < p.nq; q += nparts) {
        if (p.use_xchg) {
            // (a barrier time-out leaves this rank's flags unset: the peers' pull kernels then time out and report)
            if (ok) fused_push(xa, q, tid, base);
            __syncthreads();
            continue;
        }
        if (!ok) {
            for (int j = tid; j < p.k; j += UMMA_THREADS) {
                p.D[(size_t)q * p.k + j] = p.largest ? -3.402823466e+38f : 3.402823466e+38f;
                p.I[(size_t)q * p.k + j] = -1;
            }
            if (tid == 0 && p.status) { *(volatile int*)p.status = 2; __threadfence_system(); }
            continue;
        }
        fused_merge_plain(p.cand, nparts, p.nq_total, p.k, p.sortn, p.out_mode, p.qnorm, p.id_offset, p.rr, p.D, p.I, q, tid, base);
        __syncthreads();
    }
}

// CL = thread-block cluster size.  CL == 1: one CTA per SM streams its own tiles (bandwidth-bound
// batches, nq <= 128).  CL = 2 / 4 (nq > 128): the CTAs of a cluster hold DIFFERENT 128-query blocks
// in tensor memory and share every corpus stage -- the stage's bulk copies are dealt round-robin to
// the CTAs and each one is multicast into all CL shared memories (cp.async.bulk ...
// .multicast::cluster); a stage is recycled when the MMAs of all CL CTAs have retired
// (tcgen05.commit ... .multicast::cluster onto every CTA's `empty` barrier).  One pass over HBM then
// serves CL*128 queries.
//
// NB = T64 row blocks per MMA tile (tile = 64*NB corpus rows).  Measured on B200: a tcgen05.mma
// costs the issuing thread ~54 cycles whatever its width, so with N = 64 (32 cycles of tensor work)
// a tile costs ~1.7k + 54 * 4 * (pitch/64) cycles -- above the HBM time of a 64-row tile once
// pitch <= 512.  NB = 2 (N = 128) halves the MMA count per row and is used whenever the query block
// leaves room in tensor memory for two 128-column accumulator buffers (pitch <= 512: d = 384 went
// from 74 % to 86 % of the HBM roofline); pitch 768 keeps NB = 1 with double buffering (a single
// 128-column buffer serialises MMA and epilogue: 93.7 % vs 95.3 %).
template <int CL, int NB, int PAIR = 0>
__global__ void __launch_bounds__(UMMA_THREADS, 1) flat_scan_umma_kernel(const UmmaParams p) {
    static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
    constexpr int TILE_N = NB * BLK_ROWS;
    // one k-block of a tile in THIS CTA's shared memory: NB adjacent 8 KB block pieces; a pair CTA holds HALF of the tile's
    // rows -- one block of the two (NB = 2) or rows [32 * rank, 32 * rank + 32) of the block (NB = 1: 4 KB per k-block)
    constexpr int UMMA_KB_STAGE_BYTES = PAIR ? NB * KBLOCK_BYTES / 2 : NB * KBLOCK_BYTES;
    // which query block of the pass this CTA serves = its rank in the (CL, 1, 1) cluster.  Taken from blockIdx (clusters are
    // consecutive blocks) rather than %cluster_ctarank: the compiler then knows it is uniform, which the pair's epilogue needs
    // to keep its uniform-datapath branches (see the codegen note in DESIGN.md)
    const int crank = CL > 1 ? (int)(blockIdx.x % CL) : 0;
    const int part = (int)blockIdx.x / CL;                            // cluster index = candidate-list slot
    const int nparts = (int)gridDim.x / CL;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
    extern __shared__ unsigned char umma_smem_raw[];
    const uint32_t STAGE_BYTES = (uint32_t)p.kbs * UMMA_KB_STAGE_BYTES;
    const int nbuf = p.nbuf;                                           // accumulator buffers (1 or 2)
    const uint32_t D_OFF = UMMA_TMEM_COLS - (uint32_t)nbuf * TILE_N;   // accumulator columns at the top

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // 1024-byte aligned carve-up (SWIZZLE_128B atoms are 1024 bytes)
    const uint32_t raw = smem_u32(umma_smem_raw);
    unsigned char* base = umma_smem_raw + (((raw + 1023u) & ~1023u) - raw);
    unsigned char* ring = base;
    float* snorm = reinterpret_cast<float*>(ring + (size_t)p.stages * STAGE_BYTES);      // [4][TILE_N]
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(snorm) + 4 * TILE_N * 4);
    uint64_t* empty = full + UMMA_MAX_STAGES;
    uint64_t* tmem_full = empty + UMMA_MAX_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* peer_ready = tmem_empty + 2;                             // pair: the odd CTA's queries are in ITS tensor memory
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(peer_ready + 1);
    int* s_flag = reinterpret_cast<int*>(tmem_ptr + 1);

    const int kblocks = p.pitch >> 6;
    const int kstages = kblocks / p.kbs;           // pipeline stages per tile
    const long long n_blocks = (p.n_rows + BLK_ROWS - 1) / BLK_ROWS;
    const long long n_tiles_all = (n_blocks + NB - 1) / NB;
    const long long n_tiles = (n_tiles_all + p.tile_step - 1) / p.tile_step;   // tiles this launch visits: t * tile_step

    if (tid == 0) {
        // pair: the leader's `full` also takes the peer's "my half landed", its `tmem_empty` the peer's four epilogue warps;
        // one multicast commit frees a stage in both CTAs
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], (PAIR && crank == 0) ? 2 : 1); mbar_init(&empty[s], PAIR ? 1 : CL); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], PAIR ? 8 : 4); }
        mbar_init(peer_ready, 1);
        mbar_fence_init();
    }
    if (warp == 0) { if (PAIR) tmem_alloc_pair(tmem_ptr, UMMA_TMEM_COLS); else tmem_alloc(tmem_ptr, UMMA_TMEM_COLS); }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // every CTA's barriers exist before a peer multicasts onto them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // The producer starts streaming at once; the epilogue warps stage the queries into TMEM
    // meanwhile and release the MMA warp through named barrier 2 (128 + 32 threads).
    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            const size_t blk_bytes = (size_t)BLK_ROWS * p.pitch * 2;
            for (long long ti = part; ti < n_tiles; ti += nparts) {
                const long long t = ti * p.tile_step;
                const int nblk = (int)((n_blocks - NB * t < NB) ? (n_blocks - NB * t) : NB);   // the last tile may be short
                const unsigned char* src = p.x + (size_t)(NB * t) * blk_bytes;
                if (PAIR && NB == 2) {
                    // this CTA's half of the tile: row block NB*t + crank (nothing if the last tile has one block only)
                    const bool have = crank < nblk;
                    for (int ks = 0; ks < kstages; ++ks) {
                        mbar_wait(&empty[s], ph ^ 1u);
                        mbar_arrive_expect_tx(&full[s], have ? (uint32_t)p.kbs * KBLOCK_BYTES : 0u);
                        if (have) {
                            const unsigned char* g = src + (size_t)crank * blk_bytes + (size_t)(ks * p.kbs) * KBLOCK_BYTES;
                            // the kbs k-blocks of one row block are contiguous in HBM and in the stage: ONE copy
                            bulk_g2s(ring + (size_t)s * STAGE_BYTES, g, (uint32_t)p.kbs * KBLOCK_BYTES, &full[s]);
                        }
                        if (++s == p.stages) { s = 0; ph ^= 1u; }
                    }
                    continue;
                }
                if (PAIR && NB == 1) {
                    // 64-row tiles: rows [32 * crank, +32) of the block -- the first / second 4 KB of every 8 KB k-block piece
                    for (int ks = 0; ks < kstages; ++ks) {
                        mbar_wait(&empty[s], ph ^ 1u);
                        mbar_arrive_expect_tx(&full[s], (uint32_t)p.kbs * (KBLOCK_BYTES / 2));
                        for (int kbi = 0; kbi < p.kbs; ++kbi)
                            bulk_g2s(ring + (size_t)s * STAGE_BYTES + (size_t)kbi * (KBLOCK_BYTES / 2),
                                     src + (size_t)(ks * p.kbs + kbi) * KBLOCK_BYTES + (size_t)crank * (KBLOCK_BYTES / 2), KBLOCK_BYTES / 2, &full[s]);
                        if (++s == p.stages) { s = 0; ph ^= 1u; }
                    }
                    continue;
                }
                for (int ks = 0; ks < kstages; ++ks) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (UMMA_DBG(p) & 16) {                     // experiment: no copies at all (what do MMA + epilogue cost alone?)
                        mbar_arrive_expect_tx(&full[s], 0u);
                        if (++s == p.stages) { s = 0; ph ^= 1u; }
                        continue;
                    }
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(p.kbs * nblk) * KBLOCK_BYTES);
                    unsigned char* dst = ring + (size_t)s * STAGE_BYTES;
                    int c = 0;
                    for (int kbi = 0; kbi < p.kbs; ++kbi) {
                        for (int bl = 0; bl < nblk; ++bl, ++c) {
                            // k-block (ks*kbs + kbi) of row block NB*t + bl: 8 KB contiguous in HBM (T64 layout)
                            const unsigned char* g = src + (size_t)bl * blk_bytes + (size_t)(ks * p.kbs + kbi) * KBLOCK_BYTES;
                            unsigned char* d = dst + (size_t)kbi * UMMA_KB_STAGE_BYTES + (size_t)bl * KBLOCK_BYTES;
                            if (CL == 1) bulk_g2s(d, g, KBLOCK_BYTES, &full[s]);
                            else if ((c % CL) == crank) bulk_g2s_multicast(d, g, KBLOCK_BYTES, &full[s], CMASK);
                        }
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (one thread) ----------------
        // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1,
        // a/b_format [7,10)/[10,13) (0 F16, 1 BF16), K-major A and B, N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t fmt = p.is_bf16 ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TILE_N >> 3) << 17) |
                               ((uint32_t)((PAIR ? 2 * UMMA_M : UMMA_M) >> 4) << 24);
        int s = 0, it = 0;
        uint32_t ph = 0;
        named_bar_sync(2, 160);                               // queries are in TMEM
        tc_fence_after();
        if (PAIR && crank != 0) {
            // The odd CTA of a pair issues no MMA.  Its idle issuer thread tells the leader that this CTA's queries are
            // staged, then forwards "my half of the stage landed" to the leader's `full` barrier, stage after stage.
            if (elect_one()) {
                mbar_arrive_remote(peer_ready, 0u);
                for (long long ti = part; ti < n_tiles; ti += nparts) {
                    for (int ks = 0; ks < kstages; ++ks) {
                        mbar_wait(&full[s], ph);
                        mbar_arrive_remote(&full[s], 0u);
                        if (++s == p.stages) { s = 0; ph ^= 1u; }
                    }
                }
            }
        } else
        // ONE elected thread runs the whole issue loop (elect.sync: the compiler then emits straight-line uniform-datapath
        // code; with `if (lane == 0)` every tcgen05.mma was wrapped in an ELECT / BRA.U.ANY retry loop).  Per MMA the
        // loop is two adds and the instruction itself: the shared-memory descriptor of a stage is built once and only its
        // 14-bit address field is stepped (32 bytes per k16 slice, one 8 KB block piece per k-block), the accumulate
        // flag is a compile-time predicate except for the first slice of a tile.  Measured before: ~12 instructions and
        // ~100 cycles per MMA in this warp, i.e. the ISSUE loop -- not the tensor pipe (36 % busy) -- bound large batches.
        if (elect_one()) {
            constexpr uint64_t DESC_HI = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            if (PAIR) { mbar_wait(peer_ready, 0u); tc_fence_after(); }
#ifdef PRS_EXPERIMENTS
            long long w_acc = 0, w_full = 0, t_begin = clock64();      // PRS_UMMA_DEBUG & 32: where does the issue thread wait?
#endif
            for (long long ti = part; ti < n_tiles; ti += nparts, ++it) {
                const int b = nbuf == 2 ? (it & 1) : 0;
                const uint32_t aph = (uint32_t)(nbuf == 2 ? (it >> 1) : it) & 1u;
#ifdef PRS_EXPERIMENTS
                const long long c0 = clock64();
#endif
                mbar_wait(&tmem_empty[b], aph ^ 1u);
#ifdef PRS_EXPERIMENTS
                w_acc += clock64() - c0;
#endif
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_OFF + (uint32_t)(b * TILE_N);
                uint32_t a_tmem = tmem_base;
                for (int ks = 0; ks < kstages; ++ks) {
#ifdef PRS_EXPERIMENTS
                    const long long c1 = clock64();
#endif
                    mbar_wait(&full[s], ph);
#ifdef PRS_EXPERIMENTS
                    w_full += clock64() - c1;
#endif
                    tc_fence_after();
                    uint64_t bdesc = DESC_HI | (uint64_t)((smem_u32(ring + (size_t)s * STAGE_BYTES) >> 4) & 0x3FFFu);
                    for (int kbi = 0; kbi < p.kbs; ++kbi) {
                        if (PAIR) {
                            if (ks == 0 && kbi == 0) umma_ts_f16_pair<0>(d_tmem, a_tmem, bdesc, idesc);
                            else umma_ts_f16_pair<1>(d_tmem, a_tmem, bdesc, idesc);
#pragma unroll
                            for (int k4 = 1; k4 < 4; ++k4) umma_ts_f16_pair<1>(d_tmem, a_tmem + (uint32_t)(k4 * 8), bdesc + (uint64_t)(k4 * 2), idesc);
                        } else {
                            if (ks == 0 && kbi == 0) umma_ts_f16_c<0>(d_tmem, a_tmem, bdesc, idesc);     // first slice of the tile overwrites
                            else umma_ts_f16_c<1>(d_tmem, a_tmem, bdesc, idesc);
#pragma unroll
                            for (int k4 = 1; k4 < 4; ++k4) umma_ts_f16_c<1>(d_tmem, a_tmem + (uint32_t)(k4 * 8), bdesc + (uint64_t)(k4 * 2), idesc);
                        }
                        a_tmem += 32u;
                        bdesc += (uint64_t)(UMMA_KB_STAGE_BYTES >> 4);
                    }
                    if (PAIR) umma_commit_pair(&empty[s]);     // frees the stage in both CTAs of the pair
                    else if (CL == 1) umma_commit(&empty[s]);  // frees the smem stage when these MMAs retire
                    else umma_commit_multicast(&empty[s], CMASK);   // ... in every CTA of the cluster
                    if (ks == kstages - 1) { if (PAIR) umma_commit_pair(&tmem_full[b]); else umma_commit(&tmem_full[b]); }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
#ifdef PRS_EXPERIMENTS
            if ((UMMA_DBG(p) & 32) && blockIdx.x == 0)
                printf("issue thread of CTA 0: %d tiles, total %lld cycles, waiting for the accumulator %lld, for stages %lld\n", it, clock64() - t_begin, w_acc, w_full);
#endif
        }
    } else {
        // ---------------- epilogue: one thread per query ----------------
        // Queries are spread over the four lane quarters (query i -> TMEM lane (i&3)*32 + (i>>2)) so
        // that a partly filled pass loads the four epilogue warps evenly.
        const int qd = warp & 3;
        const int m = qd * 32 + lane;
        const int qi = (lane << 2) | qd;
        const int my_nq = p.nq - crank * UMMA_M;                          // valid queries of this CTA's block
        const bool qvalid = qi < my_nq;
        float* wnorm = snorm + (warp - 2) * TILE_N;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16) + D_OFF;
        {
            // stage this thread's query row into TMEM: lane m, columns [0, pitch/2).  The MMA warp waits
            // for this, so the row is fetched with 24 independent 16-byte loads in flight per round
            // (three TMEM stores per round) instead of 8; empty slots store zeros without reading.
            const uint32_t a_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
            const int ncol = p.pitch >> 1;                     // 32-bit TMEM columns of the A operand
            if (!p.fuse_prep) {
                const uint4* qrow = reinterpret_cast<const uint4*>(p.qlow + ((size_t)crank * UMMA_M + m) * p.pitch);
                auto stage = [&](auto NG, int c0) {                 // NG x 32 columns starting at c0
                    constexpr int G_ = decltype(NG)::value;
                    uint32_t v[G_][32];
#pragma unroll
                    for (int g = 0; g < G_; ++g) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            uint4 t4 = make_uint4(0u, 0u, 0u, 0u);
                            if (qvalid) t4 = __ldg(&qrow[((c0 + g * 32) >> 2) + i]);
                            v[g][4 * i] = t4.x; v[g][4 * i + 1] = t4.y; v[g][4 * i + 2] = t4.z; v[g][4 * i + 3] = t4.w;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < G_; ++g) tmem_st32(a_addr + (uint32_t)(c0 + g * 32), v[g]);
                };
                int c0 = 0;
                for (; c0 + 96 <= ncol; c0 += 96) stage(std::integral_constant<int, 3>{}, c0);
                if (ncol - c0 == 64) stage(std::integral_constant<int, 2>{}, c0);
                else if (ncol - c0 == 32) stage(std::integral_constant<int, 1>{}, c0);
            } else {
                // one-launch search: this thread converts ITS query row on the way into tensor memory (out of line:
                // the conversion code is large and must not sit between the hot loops of the three warp roles)
                const size_t qg = (size_t)p.q0 + (size_t)crank * UMMA_M + qi;
                const float qn = stage_query_row(p.q, p.qdtype, p.d, p.is_bf16, a_addr, ncol, qvalid, qg * (size_t)p.d, p.l2);
                if (part == 0 && qvalid && p.qnorm) p.qnorm[qg] = qn;
            }
            tmem_st_wait();
            tc_fence_before();
            named_bar_sync(2, 160);
        }
        const int G = nparts;
        uint32_t* const boot = p.boot + (size_t)crank * p.boot_stride;    // one bootstrap array per query block
        float thr = qvalid ? -INFINITY : INFINITY;   // admission threshold = max(bootstrap bound, this CTA's k-th best); empty slots admit nothing
        bool boot_done = false;
        // mode 1 (wide k): fixed threshold tau[q]; every score that reaches it goes, unsorted, into this
        // thread's private slice of the collection buffer (no atomics: one thread owns (part, query))
        const bool collect = p.mode == 1;
        const size_t qglobal = (size_t)p.q0 + (size_t)crank * UMMA_M + qi;
        u64* const cslice = collect ? p.coll + ((size_t)part * p.nq_total + qglobal) * (size_t)p.coll_cap : nullptr;
        int ccount = 0;
        if (collect) { boot_done = true; thr = qvalid ? __ldg(p.tau + qglobal) : INFINITY; }
        if (p.mode == 2) boot_done = true;
        // thread-private top-k of this CTA for this query: 16 sorted keys in registers (key 0 = empty)
        // The list is RIGHT-ALIGNED: slots [0, 16 - k) hold an all-ones sentinel no key can displace, the k live slots
        // follow, so the k-th best is always top[15] -- a fixed register.  (With the list left-aligned the k-dependent
        // reads made the compiler keep all of top[] in local memory: 16 loads + 16 stores + a chain of 16 predicated
        // loads per insertion, ~1 500 cycles in a warp that has its scheduler to itself.)
        u64 top[UMMA_MAX_K];
#pragma unroll
        for (int j = 0; j < UMMA_MAX_K; ++j) top[j] = (j < UMMA_MAX_K - p.k) ? ~0ull : 0ull;
        uint32_t best_ord = 0u;                      // ord() of this thread's best score so far (0 = none yet)

        // Bootstrap bound.  Every CTA publishes, per query, the best score of its FIRST tile; groups
        // of 4 CTAs are folded to their maximum and the k-th largest group maximum is taken.  Those
        // are scores of k distinct rows, so the value is a lower bound of the global k-th best:
        // pruning with it is exact, and it replaces the ~k*ln(rows per CTA / k) warm-up insertions
        // each CTA would otherwise need to find the threshold on its own.
        auto refresh_boot = [&]() {
            const int published = (int)__ldcg(reinterpret_cast<const unsigned int*>(boot + (size_t)G * UMMA_M));
            // all loads are independent and issued back to back (an L2 round trip under a saturated
            // HBM pipe is ~1 us: 148 dependent ones would cost more than the whole bootstrap saves)
            constexpr int NGRP = 40, FOLD = 4;              // covers up to 160 CTAs
            uint32_t gm[NGRP];
#pragma unroll
            for (int g = 0; g < NGRP; ++g) {
                uint32_t a[FOLD];
#pragma unroll
                for (int f = 0; f < FOLD; ++f) {
                    const int c = g * FOLD + f;
                    a[f] = (c < G) ? __ldcg(boot + (size_t)c * UMMA_M + m) : 0u;
                }
                gm[g] = max(max(a[0], a[1]), max(a[2], a[3]));
            }
            uint32_t best[UMMA_MAX_K];                      // right-aligned like top[]: the k-th largest ends up in best[15]
#pragma unroll
            for (int j = 0; j < UMMA_MAX_K; ++j) best[j] = (j < UMMA_MAX_K - p.k) ? 0xFFFFFFFFu : 0u;
#pragma unroll
            for (int g = 0; g < NGRP; ++g) {
                if (g * FOLD >= G) break;                   // uniform: clusters of 2 publish 74 values, not 148
                uint32_t v = gm[g];
#pragma unroll
                for (int j = 0; j < UMMA_MAX_K; ++j) {
                    const uint32_t hi = max(best[j], v);
                    v = min(best[j], v);
                    best[j] = hi;
                }
            }
            const uint32_t kth = best[UMMA_MAX_K - 1];
            if (kth) thr = fmaxf(thr, ord2f(kth));
            boot_done = published >= 4 * G;
        };

        // one 64-column half of the tile: TMEM -> registers
        auto load_half = [&](uint32_t (&v)[BLK_ROWS], int b, int h) {
            tmem_ld32(lane_addr + (uint32_t)(b * TILE_N + h * BLK_ROWS), *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tmem_ld32(lane_addr + (uint32_t)(b * TILE_N + h * BLK_ROWS + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            tmem_ld_wait();
            if (p.l2) {
#pragma unroll
                for (int j = 0; j < BLK_ROWS; ++j) v[j] = __float_as_uint(fmaf(2.f, __uint_as_float(v[j]), -wnorm[h * BLK_ROWS + j]));
            }
        };
        // fast path: one compare per score into a bit mask (no branches, small code)
        auto survivors = [&](const uint32_t (&v)[BLK_ROWS], int nv, uint32_t (&hm)[2]) {
            // Once the thresholds are tight almost no tile holds an admissible score: take the maximum of the 64
            // scores first (3-input max: 32 instructions) and build the bit masks (128+ instructions) only when it
            // reaches the threshold.  The four epilogue warps have a scheduler each and no other warp to hide their
            // latencies behind, so their instruction count per tile is what bounds tensor-bound batches.
            {
                float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
                for (int j = 0; j < BLK_ROWS; j += 4) {
                    m0 = fmax3(m0, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                    m1 = fmax3(m1, __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
                // (invalid columns of a short tile and NaNs are sorted out by the masks; a NaN maximum compares false,
                // exactly like the per-score test)
                if (!(fmaxf(m0, m1) >= thr) && nv >= BLK_ROWS) { hm[0] = hm[1] = 0u; return; }
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                hm[hh] = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) hm[hh] |= (__uint_as_float(v[hh * 32 + j]) >= thr ? 1u : 0u) << j;
                const int left = nv - hh * 32;
                if (left < 32) hm[hh] &= (left <= 0) ? 0u : ((1u << left) - 1u);
                if (!qvalid) hm[hh] = 0u;
            }
        };
        // rare path: a score reached the threshold.  One copy of the code for all columns, written for instruction-level
        // parallelism: the epilogue warp is alone on its scheduler, so a dependent chain costs its full latency per link.
        auto insert_hits = [&](const uint32_t (&v)[BLK_ROWS], const uint32_t (&hm)[2], long long rbase) {
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                uint32_t mask = hh ? hm[1] : hm[0];
#pragma unroll 1
                while (mask) {
                    const int j = __ffs(mask) - 1 + hh * 32;
                    mask &= mask - 1;
                    // v[j]: a six-level select tree (63 selects, depth 6) instead of a 64-long dependent chain
                    uint32_t s5[32], s4[16], s3[8], s2[4];
#pragma unroll
                    for (int i = 0; i < 32; ++i) s5[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
                    for (int i = 0; i < 16; ++i) s4[i] = (j & 2) ? s5[2 * i + 1] : s5[2 * i];
#pragma unroll
                    for (int i = 0; i < 8; ++i) s3[i] = (j & 4) ? s4[2 * i + 1] : s4[2 * i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) s2[i] = (j & 8) ? s3[2 * i + 1] : s3[2 * i];
                    const uint32_t s1a = (j & 16) ? s2[1] : s2[0], s1b = (j & 16) ? s2[3] : s2[2];
                    const uint32_t bits = (j & 32) ? s1b : s1a;
                    const float sc = __uint_as_float(bits);
                    if (collect) {
                        if (ccount == p.coll_cap) {
                            // slice full (thousands of rows tie with tau): keep exactly its k best and raise
                            // the admission threshold.  This thread visits rows in ascending id order, so a
                            // later row that only TIES with the k-th kept score can never displace it
                            // (lower id wins): from here on a score must be strictly better.
                            const u64 kth = slice_kth_largest(cslice, p.coll_cap, p.k);
                            int w = 0;
                            for (int i = 0; i < p.coll_cap; ++i) { const u64 kk = cslice[i]; if (kk >= kth) cslice[w++] = kk; }
                            ccount = w;
                            thr = fmaxf(thr, nextafterf(key_score(kth), INFINITY));
                        }
                        if (sc >= thr) cslice[ccount++] = make_key<PRS_TIE_LOW_ID>(sc, (uint32_t)(rbase + j));
                    } else if (sc >= thr) {
                        // sorted insertion (descending), every slot on its own: slots above the key keep their value, the
                        // first slot below it takes the key, the rest take their upper neighbour (the last one falls off)
                        const u64 key = make_key<PRS_TIE_LOW_ID>(sc, (uint32_t)(rbase + j));
                        bool below[UMMA_MAX_K];
#pragma unroll
                        for (int i = 0; i < UMMA_MAX_K; ++i) below[i] = key > top[i];
                        u64 nt[UMMA_MAX_K];
#pragma unroll
                        for (int i = 0; i < UMMA_MAX_K; ++i) {
                            const u64 shifted = (i > 0 && below[i - 1]) ? top[i > 0 ? i - 1 : 0] : key;
                            nt[i] = below[i] ? shifted : top[i];
                        }
#pragma unroll
                        for (int i = 0; i < UMMA_MAX_K; ++i) top[i] = nt[i];
                        best_ord = max(best_ord, (uint32_t)(key >> 32));
                        if (top[UMMA_MAX_K - 1]) thr = fmaxf(thr, key_score(top[UMMA_MAX_K - 1]));
                    }
                }
            }
        };

        int it = 0;
        for (long long ti = part; ti < n_tiles; ti += nparts, ++it) {
            const long long t = ti * p.tile_step;
            const int b = nbuf == 2 ? (it & 1) : 0;
            const uint32_t aph = (uint32_t)(nbuf == 2 ? (it >> 1) : it) & 1u;
            const long long row0 = t * TILE_N;
            const int nvalid = (int)((p.n_rows - row0 < TILE_N) ? (p.n_rows - row0) : TILE_N);
            if (p.l2) {
                __syncwarp();
#pragma unroll
                for (int i = 0; i < TILE_N / 32; ++i) {
                    const int c = lane + 32 * i;
                    wnorm[c] = (c < nvalid) ? p.xnorm[row0 + c] : 0.f;
                }
                __syncwarp();
            }
            mbar_wait(&tmem_full[b], aph);
            tc_fence_after();
            uint32_t v[BLK_ROWS];
            uint32_t hm[2];
            if (UMMA_DBG(p) & 2) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(&tmem_empty[b], 0u); else mbar_arrive(&tmem_empty[b]); }
                continue;
            }
            if (p.mode == 2) {
                // threshold sampling: only the best score of this tile per query (branch free, no lists)
                float mx = -INFINITY;
#pragma unroll 1
                for (int h = 0; h < NB; ++h) {
                    load_half(v, b, h);
                    if (h == NB - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(&tmem_empty[b], 0u); else mbar_arrive(&tmem_empty[b]); }
                    }
                    const int nv = nvalid - h * BLK_ROWS;
#pragma unroll
                    for (int j = 0; j < BLK_ROWS; ++j) mx = (j < nv) ? fmaxf(mx, __uint_as_float(v[j])) : mx;
                }
                if (qvalid) p.gmax[qglobal * p.gmax_stride + ti] = sanitize(mx);
                continue;
            }
            if (UMMA_DBG(p) & 1) { thr = INFINITY; boot_done = true; }
            else if (it == 0 && p.mode == 0) {
                // bootstrap pass over the first tile (both halves): best score per query, then the bound
                float mx = -INFINITY;
#pragma unroll 1
                for (int h = 0; h < NB; ++h) {
                    load_half(v, b, h);
                    const int nv = nvalid - h * BLK_ROWS;
#pragma unroll
                    for (int j = 0; j < BLK_ROWS; ++j) mx = (j < nv) ? fmaxf(mx, __uint_as_float(v[j])) : mx;
                }
                boot[(size_t)part * UMMA_M + m] = f2ord(mx);
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(reinterpret_cast<unsigned int*>(boot + (size_t)G * UMMA_M), 1u);
                // bounded wait for the other CTAs (they start together and do the same work); a late
                // CTA only weakens the bound, and the refresh is repeated while it is incomplete
                for (int spin = 0; spin < 24; ++spin) {
                    if ((int)__ldcg(reinterpret_cast<const unsigned int*>(boot + (size_t)G * UMMA_M)) >= 4 * G) break;
                    __nanosleep(64);
                }
                __threadfence();
                refresh_boot();
                if (UMMA_DBG(p) & 8) thr = INFINITY;
            } else if (p.reboot && p.mode == 0 && (it & (it + 1)) == 0 && (it == 1 || (__ffs(it + 1) & 1))) {
                // it = 1, 3, 15, 63, 255, ...: publish this CTA's best score SO FAR and take the bound again.  The k-th
                // largest of the CTAs' running maxima tightens like 1 / (rows seen): between two refreshes a warp then
                // meets only a handful of admissible scores, instead of ~k ln(rows / k) per thread when every CTA has
                // to find its threshold alone (large batches were bound by exactly those insertions: 35 % of the
                // epilogue's instructions and warp divergence on 4 of 5 tiles).
                if (best_ord) boot[(size_t)part * UMMA_M + m] = best_ord;
                refresh_boot();
            } else if (!boot_done && (it & (it + 1)) == 0) {
                refresh_boot();          // first-tile maxima still arriving: it = 1, 3, 7, 15, ...
            }
            // 64 columns at a time: registers, mask, the rare insertions.  After the LAST load the
            // accumulator buffer goes straight back to the MMA warp.
#pragma unroll 1
            for (int h = 0; h < NB; ++h) {
                load_half(v, b, h);
                if (h == NB - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR && crank != 0) mbar_arrive_remote(&tmem_empty[b], 0u); else mbar_arrive(&tmem_empty[b]); }
                }
                survivors(v, nvalid - h * BLK_ROWS, hm);
                if (hm[0] | hm[1]) insert_hits(v, hm, row0 + h * BLK_ROWS);
            }
        }
        if (qvalid && collect) p.coll_cnt[(size_t)part * p.nq_total + qglobal] = ccount;
        if (qvalid && p.mode == 0) {
            const size_t o = (size_t)part * p.nq_total + p.q0 + crank * UMMA_M + qi;
            int n = 0;
#pragma unroll
            for (int j = 0; j < UMMA_MAX_K; ++j) {
                const int jj = j - (UMMA_MAX_K - p.k);                                 // the live slots are the last k
                if (jj >= 0) { p.cand[o * p.k + jj] = top[j]; n += top[j] != 0ull; }    // all k slots, 0 = empty (sorted: empties last)
            }
            p.cand_cnt[o] = n;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // no CTA leaves while a peer's commits can still arrive on its barriers
    tc_fence_after();
    if (warp == 0) { if (PAIR) tmem_dealloc_pair(tmem_base, UMMA_TMEM_COLS); else tmem_dealloc(tmem_base, UMMA_TMEM_COLS); }

    if (CL == 1 && p.fuse_merge) fused_tail(p, part, nparts, tid, base, s_flag);
}

// One launch per search: queries (fp32 / fp16 / bf16) -> [128-padded, pitch] 16-bit rows in TMEM-slot
// order, zero padded; ||q~||^2 of the ROUNDED query (what the expanded L2 form needs); and the
// bootstrap array of the scan kernel zeroed.  One warp per slot row.
template <typename TQ>
__global__ void __launch_bounds__(128) prep_queries_kernel(const TQ* __restrict__ q, long long nq, int d, int pitch, long long nq_pad,
                                                           int is_bf16, uint16_t* __restrict__ out, float* __restrict__ qnorm,
                                                           uint32_t* __restrict__ boot, long long boot_words) {
    // one CTA per slot row (pass * 128 + TMEM lane); <= 6 independent elements per thread, so the
    // kernel is one memory round trip long
    __shared__ float s_part[4];
    const int tid = threadIdx.x;
    for (long long i = (long long)blockIdx.x * 128 + tid; i < boot_words; i += (long long)gridDim.x * 128) boot[i] = 0u;
    const long long r = blockIdx.x;
    if (r >= nq_pad) return;
    const int mm = (int)(r & 127);
    const long long qsrc = (r & ~127ll) + (((mm & 31) << 2) | (mm >> 5));    // query held by that lane
    constexpr int PER = 6;                                                   // pitch <= 768
    float v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = tid + i * 128;
        v[i] = (qsrc < nq && c < d) ? (float)q[qsrc * d + c] : 0.f;
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = tid + i * 128;
        if (c < pitch) {
            uint16_t o;
            float back;
            if (is_bf16) { __nv_bfloat16 h = __float2bfloat16_rn(v[i]); o = *reinterpret_cast<uint16_t*>(&h); back = __bfloat162float(h); }
            else { __half h = __float2half_rn(v[i]); o = *reinterpret_cast<uint16_t*>(&h); back = __half2float(h); }
            out[r * pitch + c] = o;
            acc = fmaf(back, back, acc);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) s_part[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0 && qsrc < nq) qnorm[qsrc] = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
}

// ---------------- host side ----------------
struct UmmaState {
    DevBuf qlow, boot, tau, gmax, coll, coll_cnt, gbar;
    bool boot_clean = false;      // the bootstrap words are all zero (left so by the one-launch search's tail)
    // the last search converted the queries inside the scan kernel and left the merge to a separate kernel: that kernel
    // re-ranks from the ORIGINAL queries (there is no qlow image) and zeroes `zero_words` bootstrap words for the next search
    bool prep_in_scan = false;
    const void* q_orig = nullptr;
    int q_dtype = 0;
    long long zero_words = 0;
    void invalidate() {}
    void release() { qlow.release(); boot.release(); tau.release(); gmax.release(); coll.release(); coll_cnt.release(); gbar.release(); }
};

// what the one-launch search needs to know about the merge it absorbs (filled by flat_index.cu)
struct UmmaTail {
    bool enable = false;
    int out_mode = 0, largest = 1;
    long long id_offset = 0;
    float* D = nullptr;
    long long* I = nullptr;
    const unsigned char* rerank_x = nullptr;      // corpus pointer when the 16-bit L2 re-rank is on
    prs_xchg* xchg = nullptr;                      // row-sharded search: the tail pushes, xchg_pull_kernel finishes
    ScanTimer* timer_merge = nullptr;
    int device = 0;
    bool prep_in_scan = false;                     // without `enable`: the scan still converts the queries itself (two launches)
};

static inline bool umma_eligible(int storage, int d, int pitch, long long nq, int k) {
    (void)d;
    return (storage == PRS_F16 || storage == PRS_BF16) && pitch <= 768 && k <= UMMA_MAX_K && nq >= 1;
}

template <int CL, int NB, int PAIR = 0>
static inline int umma_launch(const UmmaParams& p, int n_clusters, size_t smem, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_clusters * CL), 1, 1);
    cfg.blockDim = dim3(UMMA_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (CL > 1) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    } else {
        attr[0].id = cudaLaunchAttributeCooperative;      // the one-launch search has a grid barrier: all CTAs co-resident
        attr[0].val.cooperative = 1;
    }
    cfg.attrs = attr; cfg.numAttrs = (CL > 1 || p.fuse_merge) ? 1 : 0;
    PRS_CUDA(cudaLaunchKernelEx(&cfg, flat_scan_umma_kernel<CL, NB, PAIR>, p));
    return 0;
}

// how many clusters of CL CTAs (1 CTA per SM at this shared-memory size) the device runs at once.
// Cached per (device, shared-memory size) under a mutex: indices on different devices and threads
// plan concurrently.
template <int CL, int NB, int PAIR = 0>
static inline int umma_max_clusters(size_t smem, int sm_count) {
    static std::mutex mu;
    static std::map<std::pair<int, size_t>, int> cache;
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({device, smem});
    if (it != cache.end()) return it->second;
    if (cudaFuncSetAttribute(flat_scan_umma_kernel<CL, NB, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int n = sm_count / CL;
    if (CL > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(sm_count / CL * CL), 1, 1);
        cfg.blockDim = dim3(UMMA_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int q = 0;
        if (cudaOccupancyMaxActiveClusters(&q, flat_scan_umma_kernel<CL, NB, PAIR>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
        n = q;
    }
    cache[{device, smem}] = n;
    return n;
}

// everything about a search that depends only on shapes
struct UmmaPlan {
    int CL = 1, NB = 1, kbs = 1, stages = 2, n_clusters = 1, dbg = 0, reboot = 0, pair = 0;
    size_t smem = 0;
    long long n_tiles = 0, qblock = 128, nq_pad = 128, boot_words = 0, nblocks = 1;
};

static inline int umma_plan(long long n, int pitch, long long nq, int sm_count, UmmaPlan& pl, bool no_clusters = false, bool allow_pair = true) {
#ifdef PRS_EXPERIMENTS
    static const int dbg_pair = getenv("PRS_UMMA_PAIR") ? atoi(getenv("PRS_UMMA_PAIR")) : 0;
    if (dbg_pair != 1) allow_pair = false;        // the CTA-pair variant is an experiment (PRS_UMMA_PAIR=1): see below
    static const int dbg = getenv("PRS_UMMA_DEBUG") ? atoi(getenv("PRS_UMMA_DEBUG")) : 0;
    static const int dbg_kbs = getenv("PRS_UMMA_KBS") ? atoi(getenv("PRS_UMMA_KBS")) : 0;
    static const int dbg_cl = getenv("PRS_UMMA_CLUSTER") ? atoi(getenv("PRS_UMMA_CLUSTER")) : 0;
    static const int dbg_nb = getenv("PRS_UMMA_NB") ? atoi(getenv("PRS_UMMA_NB")) : 0;
#else
    const int dbg = 0, dbg_kbs = 0, dbg_cl = 0, dbg_nb = 0;
#endif
    if (n > 0x7FFFFFFFll - 2 * BLK_ROWS) { set_error("tcgen05 path: more than 2^31 rows per shard"); return PRS_EUNSUP; }
    pl.dbg = dbg;
    const int kblocks = pitch >> 6;
    pl.NB = pitch <= 512 ? 2 : 1;                   // row blocks per MMA tile (two accumulator buffers must fit TMEM)
    // tensor-bound batches at pitch 768: 128-row tiles with ONE accumulator buffer (half the MMA instructions per row --
    // a TS-mode MMA costs ~54 cycles whatever N <= 64 is -- while the now lean epilogue and the MMAs of a tile no longer
    // overlap): B = 256 / 1024 / 4096 0.440 / 1.715 / 7.28 -> 0.429 / 1.686 / 6.90 ms (profiles/r2_bigbatch_experiments.log)
    if (pitch > 512 && nq > UMMA_M) pl.NB = 2;
    if (dbg_nb == 1) pl.NB = 1;
    if (dbg_nb == 2) pl.NB = 2;                     // experiment: 128-row tiles even when only ONE accumulator buffer fits (pitch > 512)
    // EXPERIMENT (experiments build, PRS_UMMA_PAIR=1): batches above 128 queries as CTA PAIRS (tcgen05 cta_group::2) -- the two
    // CTAs of a cluster hold 128 queries each and HALF of every tile; the even CTA issues one M = 256 MMA over both halves.
    // Against two CTAs that each receive the whole tile by multicast this halves the bytes every SM takes in, and the
    // M = 256, N = 128 instruction runs at the full rate (64 cycles per K = 16: tools/probes/pair_probe.cu).  Correct (the
    // parity tests pass) but NOT faster end to end: the accumulator hand-over (commit -> both epilogues -> remote arrive ->
    // issue thread) costs ~2 600 cycles per tile across two SMs against ~1 250 inside one, and the epilogue's tensor-memory
    // reads slow down when the MMAs run at full rate: 1M x 768, B = 1024: 1.60 ms against 1.53 ms for the multicast
    // clusters; d = 384: 0.90 against 0.82 ms (profiles/r2_bigbatch_experiments.log, fourth series).
#ifdef PRS_EXPERIMENTS
    pl.pair = (allow_pair && nq > UMMA_M && !no_clusters && dbg_cl <= 0) ? 1 : 0;
#else
    pl.pair = 0;
    (void)allow_pair;
#endif
    if (pl.pair) pl.NB = pitch <= 512 ? 2 : 1;      // two accumulator buffers always: 2 x 128 columns beside <= 512-wide queries, else 2 x 64
    if (pl.pair && dbg_nb >= 1 && dbg_nb <= 2) pl.NB = dbg_nb;
    const int half_blocks = pl.pair ? pl.NB : 2 * pl.NB;   // 4 KB half block pieces per k-block in ONE CTA's shared memory
    // k-blocks per pipeline stage: stages of up to 48 KB.  Few, large stages keep the per-stage
    // barrier round trips of the single MMA-issuing thread off the critical path (measured: 8 KB
    // stages 3443 GB/s, 16 KB 4431, 48 KB 4513 -> see profiles/)
    pl.kbs = 1;
    for (int c = 2; c <= 12 / half_blocks; ++c) if (kblocks % c == 0) pl.kbs = c;
    if (dbg_kbs > 0 && kblocks % (dbg_kbs > 0 ? dbg_kbs : 1) == 0) pl.kbs = dbg_kbs;
    const size_t stage_bytes = (size_t)pl.kbs * half_blocks * (KBLOCK_BYTES / 2);
    const size_t fixed = 4 * (size_t)pl.NB * BLK_ROWS * 4 + (2 * UMMA_MAX_STAGES + 5) * 8 + 16;
    pl.stages = (int)((226 * 1024 - 1024 - fixed) / stage_bytes);
    if (pl.stages > UMMA_MAX_STAGES) pl.stages = UMMA_MAX_STAGES;
    pl.smem = 1024 + (size_t)pl.stages * stage_bytes + fixed;
    pl.n_tiles = (n + pl.NB * BLK_ROWS - 1) / (pl.NB * BLK_ROWS);
    // cluster size: query blocks that share one pass over the corpus
    // Clusters of 2 for every batch above 128 queries.  Clusters of 4 halve the passes over HBM again, but only 33 of
    // them fit the 148 SMs (132 SMs busy) and these batches are tensor bound, not HBM bound: measured at 1M x 768,
    // B = 512 / 1024 / 4096: 1.13 / 2.24 / 8.92 ms with clusters of 4 versus 0.97 / 1.93 / 8.62 ms with clusters of 2
    // (profiles/r2_bigbatch_experiments.log).
    pl.CL = nq > UMMA_M ? 2 : 1;
    // tensor-bound batches are limited by the epilogue's insertions: keep the cross-CTA bound fresh (HBM-bound batches have
    // epilogue slack and do not need the extra refreshes)
    pl.reboot = nq > UMMA_M ? 1 : 0;
#ifdef PRS_EXPERIMENTS
    { static const int x = getenv("PRS_UMMA_REBOOT") ? atoi(getenv("PRS_UMMA_REBOOT")) : -1; if (x >= 0) pl.reboot = x; }
#endif
    if (dbg_cl == 1 || dbg_cl == 2 || dbg_cl == 4) pl.CL = dbg_cl;
    if (no_clusters) pl.CL = 1;
    int max_clusters = 0;
#ifdef PRS_EXPERIMENTS
    if (pl.pair) {
        max_clusters = pl.NB == 2 ? umma_max_clusters<2, 2, 1>(pl.smem, sm_count) : umma_max_clusters<2, 1, 1>(pl.smem, sm_count);
        if (max_clusters <= 0) return umma_plan(n, pitch, nq, sm_count, pl, no_clusters, false);   // no pairs on this device: multicast clusters
    }
#endif
    for (; !pl.pair;) {
        const int CL = pl.CL;
        if (pl.NB == 2) max_clusters = CL == 4 ? umma_max_clusters<4, 2>(pl.smem, sm_count) : (CL == 2 ? umma_max_clusters<2, 2>(pl.smem, sm_count) : umma_max_clusters<1, 2>(pl.smem, sm_count));
        else max_clusters = CL == 4 ? umma_max_clusters<4, 1>(pl.smem, sm_count) : (CL == 2 ? umma_max_clusters<2, 1>(pl.smem, sm_count) : umma_max_clusters<1, 1>(pl.smem, sm_count));
        if (max_clusters > 0 || pl.CL == 1) break;
        pl.CL >>= 1;                                // this device cannot co-schedule such clusters
    }
    if (max_clusters <= 0) { set_error("tcgen05 path: kernel does not fit this device"); return PRS_ECUDA; }
    pl.n_clusters = (int)std::min<long long>(max_clusters, pl.n_tiles);
    pl.qblock = (long long)UMMA_M * pl.CL;
    pl.nq_pad = (nq + pl.qblock - 1) / pl.qblock * pl.qblock;
    pl.boot_words = (long long)pl.n_clusters * UMMA_M + 32;
    pl.nblocks = pl.nq_pad / UMMA_M;
    return 0;
}

// queries -> 16-bit slot rows + ||q~||^2; bootstrap arrays zeroed (one launch)
static inline int umma_prep(UmmaState& st, const UmmaPlan& pl, const void* q, int qdtype, long long nq, int d, int pitch, int storage,
                            float* qnorm, cudaStream_t stream, ScanTimer* timer_prep) {
    int rc;
    if ((rc = st.qlow.ensure((size_t)pl.nq_pad * pitch * 2))) return rc;
    if ((rc = st.boot.ensure((size_t)pl.boot_words * pl.nblocks * 4))) return rc;
    const unsigned blocks = (unsigned)pl.nq_pad;
    const int bf = storage == PRS_BF16;
    uint16_t* out = (uint16_t*)st.qlow.p;
    uint32_t* boot = (uint32_t*)st.boot.p;
    const long long bw = pl.boot_words * pl.nblocks;
    if (timer_prep) timer_prep->begin(stream);
    switch (qdtype) {
        case PRS_F32: prep_queries_kernel<float><<<blocks, 128, 0, stream>>>((const float*)q, nq, d, pitch, pl.nq_pad, bf, out, qnorm, boot, bw); break;
        case PRS_F16: prep_queries_kernel<__half><<<blocks, 128, 0, stream>>>((const __half*)q, nq, d, pitch, pl.nq_pad, bf, out, qnorm, boot, bw); break;
        case PRS_BF16: prep_queries_kernel<__nv_bfloat16><<<blocks, 128, 0, stream>>>((const __nv_bfloat16*)q, nq, d, pitch, pl.nq_pad, bf, out, qnorm, boot, bw); break;
        default: set_error("search: unsupported query dtype %d", qdtype); return PRS_EINVAL;
    }
    if (timer_prep) timer_prep->end(stream);
    PRS_LAUNCH_CHECK();
    return 0;
}

// all passes of one scan over the corpus (mode 0: sorted top-k lists of k <= 16; mode 1: collect >= tau)
static inline int umma_scan(UmmaState& st, const UmmaPlan& pl, const void* x, const float* xnorm, long long n, int pitch, int storage,
                            int metric, long long nq, int k, int tile_step, int mode, u64* cand, int* cand_cnt, const float* tau,
                            u64* coll, int* coll_cnt, int coll_cap, cudaStream_t stream, ScanTimer* timer,
                            float* gmax = nullptr, long long gmax_stride = 0, const UmmaParams* fused = nullptr) {
    for (long long q0 = 0; q0 < nq; q0 += pl.qblock) {
        UmmaParams p;
        if (fused) p = *fused;                                  // one-launch search: prologue / tail fields (single pass)
        else { p.q = nullptr; p.qdtype = 0; p.d = 0; p.fuse_prep = 0; p.fuse_merge = 0; p.qnorm = nullptr; p.gbar = nullptr; p.sortn = 0;
               p.out_mode = 0; p.largest = 1; p.id_offset = 0; p.D = nullptr; p.I = nullptr; p.rr = Rerank{}; p.status = nullptr;
               p.use_xchg = 0; p.xgen = 0; }
        p.x = (const unsigned char*)x;
        p.qlow = (const uint16_t*)st.qlow.p + (size_t)q0 * pitch;
        p.xnorm = xnorm; p.n_rows = n; p.pitch = pitch;
        p.nq = (int)std::min<long long>(pl.qblock, nq - q0);
        p.k = k; p.l2 = metric == PRS_METRIC_L2; p.stages = pl.stages; p.is_bf16 = storage == PRS_BF16; p.kbs = pl.kbs; p.dbg = pl.dbg;
        p.reboot = pl.reboot;
        p.nbuf = (pl.NB == 2 && pitch > 512) ? 1 : 2;       // two 128-column accumulators do not fit beside a 768-wide query block
        p.nq_total = (int)nq; p.q0 = (int)q0;
        p.cand = cand; p.cand_cnt = cand_cnt;
        p.tile_step = tile_step; p.mode = mode; p.tau = tau; p.coll = coll; p.coll_cnt = coll_cnt; p.coll_cap = coll_cap;
        p.gmax = gmax; p.gmax_stride = gmax_stride;
        p.boot = (uint32_t*)st.boot.p + (size_t)(q0 / UMMA_M) * pl.boot_words;
        p.boot_stride = pl.boot_words;
        if (timer) timer->begin(stream);
        int rc;
        const int CL = pl.CL, nc = pl.n_clusters;
#ifdef PRS_EXPERIMENTS
        if (pl.pair) rc = pl.NB == 2 ? umma_launch<2, 2, 1>(p, nc, pl.smem, stream) : umma_launch<2, 1, 1>(p, nc, pl.smem, stream);
        else
#endif
        if (pl.NB == 2) rc = CL == 4 ? umma_launch<4, 2>(p, nc, pl.smem, stream) : (CL == 2 ? umma_launch<2, 2>(p, nc, pl.smem, stream) : umma_launch<1, 2>(p, nc, pl.smem, stream));
        else rc = CL == 4 ? umma_launch<4, 1>(p, nc, pl.smem, stream) : (CL == 2 ? umma_launch<2, 1>(p, nc, pl.smem, stream) : umma_launch<1, 1>(p, nc, pl.smem, stream));
        if (rc) return rc;
        if (timer) timer->end(stream);
        PRS_LAUNCH_CHECK();
    }
    return 0;
}

static inline bool umma_local_device_ptr(const void* ptr, int device) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice && a.device == device;
}

// k <= 16.  q: [nq, d] device, dtype qdtype.  qnorm: [nq] device out.  cand: per-part lists out.
// With `tail` (and nq <= 128: one pass, no clusters) the whole search is ONE cooperative launch: the epilogue
// threads convert their own query rows (when q is local device memory and d % 8 == 0; otherwise the preparation
// kernel still runs -- it reads host-mapped or peer queries exactly once), and after a grid barrier the CTAs merge
// the queries among themselves (merge_query of topk_merge.cuh; row-sharded: push here + xchg_pull_kernel).  *fused tells the
// caller that D / I are already on their way.
static inline int search_umma(UmmaState& st, const void* x, const float* xnorm, long long n, int d, int pitch, int storage,
                              int metric, int sm_count, const void* q, int qdtype, long long nq, int k, float* qnorm,
                              DevBuf& cand, DevBuf& cand_cnt, int* parts_out, cudaStream_t stream, ScanTimer* timer = nullptr, ScanTimer* timer_prep = nullptr,
                              const UmmaTail* tail = nullptr, bool* fused_out = nullptr) {
    UmmaPlan pl;
    int rc;
    if (fused_out) *fused_out = false;
    if ((rc = umma_plan(n, pitch, nq, sm_count, pl))) return rc;
    const bool one_pass = tail && pl.CL == 1 && nq <= pl.qblock;
    const bool fuse_merge = one_pass && tail->enable;
    const bool fuse_prep = one_pass && (tail->enable || tail->prep_in_scan) && d % 8 == 0 && ((uintptr_t)q & 15u) == 0 &&
                           umma_local_device_ptr(q, tail->device);
    st.prep_in_scan = fuse_prep && !fuse_merge;
    st.q_orig = q; st.q_dtype = qdtype; st.zero_words = 0;
    if (!fuse_prep) {
        if ((rc = umma_prep(st, pl, q, qdtype, nq, d, pitch, storage, qnorm, stream, timer_prep))) return rc;
    } else {
        const void* before = st.boot.p;
        if ((rc = st.boot.ensure((size_t)pl.boot_words * pl.nblocks * 4))) return rc;
        if (st.boot.p != before) st.boot_clean = false;
        if (!st.boot_clean) PRS_CUDA(cudaMemsetAsync(st.boot.p, 0, st.boot.bytes, stream));
    }
    if ((rc = cand.ensure((size_t)pl.n_clusters * nq * k * 8))) return rc;
    if ((rc = cand_cnt.ensure((size_t)pl.n_clusters * nq * 4))) return rc;
    UmmaParams fp;
    const UmmaParams* fpp = nullptr;
    if (fuse_prep && !fuse_merge) {
        // two launches: the scan prepares its queries, a separate merge (+ exchange) kernel follows and clears the bootstrap words
        fp.q = q; fp.qdtype = qdtype; fp.d = d; fp.fuse_prep = 1; fp.fuse_merge = 0; fp.qnorm = qnorm;
        fp.gbar = nullptr; fp.sortn = 0; fp.out_mode = 0; fp.largest = 1; fp.id_offset = 0; fp.D = nullptr; fp.I = nullptr;
        fp.rr = Rerank{}; fp.status = nullptr; fp.use_xchg = 0; fp.xgen = 0;
        fpp = &fp;
        st.zero_words = (long long)(st.boot.bytes / 4);
    }
    if (fuse_merge) {
        if (!st.gbar.p) {
            if ((rc = st.gbar.ensure(256))) return rc;
            PRS_CUDA(cudaMemsetAsync(st.gbar.p, 0, 256, stream));
        }
        fp.q = q; fp.qdtype = qdtype; fp.d = d; fp.fuse_prep = fuse_prep ? 1 : 0; fp.fuse_merge = 1; fp.qnorm = qnorm;
        fp.gbar = (unsigned int*)st.gbar.p;
        fp.sortn = next_pow2((int)std::max<long long>(k + UMMA_THREADS, (long long)pl.n_clusters * k));
        fp.out_mode = tail->out_mode; fp.largest = tail->largest; fp.id_offset = tail->id_offset; fp.D = tail->D; fp.I = tail->I;
        fp.rr = Rerank{tail->rerank_x, fuse_prep ? nullptr : (const uint16_t*)st.qlow.p, q, qdtype, d, pitch, storage == PRS_BF16 ? 1 : 0};
        fp.status = nullptr;
        fp.use_xchg = 0; fp.xgen = 0;
        if (tail->xchg) {
            // row-sharded: the tail pushes this rank's lists into the peers' buffers; xchg_pull_kernel (below) waits for
            // theirs and merges.  Searches on one exchange context are ordered through its event.
            prs_xchg* xc = tail->xchg;
            ++xc->gen;
            if (xc->used) PRS_CUDA(cudaStreamWaitEvent(stream, xc->event, 0));
            fp.use_xchg = 1; fp.xgen = xc->gen; fp.xv = xc->view;
        }
        fpp = &fp;
    }
    if ((rc = umma_scan(st, pl, x, xnorm, n, pitch, storage, metric, nq, k, 1, 0, (u64*)cand.p, (int*)cand_cnt.p, nullptr, nullptr, nullptr, 0,
                        stream, timer, nullptr, 0, fpp))) return rc;
    st.boot_clean = fuse_merge || st.prep_in_scan;               // the fused tail / the separate merge kernel leaves the bootstrap words zeroed
    if (fuse_merge && tail->xchg) {
        prs_xchg* xc = tail->xchg;
        const int sortn2 = next_pow2(std::max(k + XCHG_PULL_THREADS, xc->G * k));
        const size_t smem2 = merge_smem_bytes(sortn2, XCHG_PULL_THREADS, k);
        if (tail->timer_merge) tail->timer_merge->begin(stream);
        xchg_pull_kernel<<<(unsigned)nq, XCHG_PULL_THREADS, smem2, stream>>>(k, sortn2, tail->out_mode, tail->largest, tail->rerank_x ? 1 : 0, xc->view,
                                                                            xc->gen, xc->timeout_ns, tail->D, tail->I, xc->d_status);
        if (tail->timer_merge) tail->timer_merge->end(stream);
        PRS_LAUNCH_CHECK();
        PRS_CUDA(cudaEventRecord(xc->event, stream));
        xc->used = true;
    }
    if (fused_out) *fused_out = fuse_merge;
    *parts_out = pl.n_clusters;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Wide k (16 < k <= 1024): sample -> threshold -> collect -> select.
//   1. the scan kernel visits every S-th tile (S <= 16) and records only the BEST score of each visited
//      tile per query (branch free: no lists, no insertions);
//   2. tau[q] = k-th largest of those M tile maxima -- scores of k distinct rows, hence a lower bound of
//      the true k-th best (with few qualifying rows per tile this is as tight as the sample's own k-th);
//   3. the full scan runs in collect mode: every score >= tau[q] is appended, unsorted, to the
//      (part, query) slice of a collection buffer -- about k * S candidates per query in total.  A slice
//      holds >= 2k entries; if it ever fills (thousands of rows tying with tau) its owner thread keeps
//      the slice's k best and raises its threshold, so the result stays exact without any fallback;
//   4. one CTA per query selects the k best of its slices into a single sorted list (then the
//      ordinary merge / merge+exchange kernel finishes it).
// Cost: (1 + 1/S) scans.  Fully asynchronous on `stream` (no host synchronisation).
// ---------------------------------------------------------------------------------------------------
constexpr int UMMA_WIDE_MAX_K = PRS_MAX_K;
static inline bool umma_wide_eligible(int storage, int pitch, long long nq, int k, long long n) {
    return (storage == PRS_F16 || storage == PRS_BF16) && pitch <= 768 && k > UMMA_MAX_K && k <= UMMA_WIDE_MAX_K && nq >= 1 &&
           n >= 256ll * k && n >= 16384;                      // >= 2k tiles of 128 rows to sample a threshold from
}
// queries per call of search_umma_wide such that the collection buffer stays below ~2 GB whatever the
// shard looks like (depends on k only: every rank of a sharded search must chunk identically)
static inline long long umma_wide_chunk(int k) {
    const long long cap = next_pow2(std::max(64, 2 * k));
    const long long per_query = 160ll * cap * 8;
    long long c = (2ll << 30) / per_query;
    c = c / UMMA_M * UMMA_M;
    return std::max<long long>(UMMA_M, c);
}

static inline int search_umma_wide(UmmaState& st, const void* x, const float* xnorm, long long n, int d, int pitch, int storage,
                                   int metric, int sm_count, const void* q, int qdtype, long long nq, int k, float* qnorm,
                                   DevBuf& cand, DevBuf& cand_cnt, cudaStream_t stream,
                                   ScanTimer* timer = nullptr, ScanTimer* timer_prep = nullptr) {
    UmmaPlan pl;
    int rc;
    // no clusters: 148 independent parts keep the collection slices short
    if ((rc = umma_plan(n, pitch, nq, sm_count, pl, true))) return rc;
    const int parts = pl.n_clusters;
    // sampling step S: the sample must hold M >= 2k tiles (each contributes its best score = one
    // distinct row); at most every 16th tile so that the collecting pass sees <= ~16 k candidates per query
    long long S = std::min<long long>(16, pl.n_tiles / (2ll * k));
    if (S < 1) S = 1;                                                // (eligibility guarantees n_tiles >= 2k)
    const long long M = (pl.n_tiles + S - 1) / S;                    // sampled tiles
    const double expect = (double)k * (double)S / parts;             // survivors per (part, query)
    const int cap = next_pow2((int)std::min<double>(1 << 20, std::max<double>(std::max(64, 2 * k), 6.0 * expect + 32.0)));
    const size_t coll_bytes = (size_t)parts * nq * cap * 8;
    if ((rc = umma_prep(st, pl, q, qdtype, nq, d, pitch, storage, qnorm, stream, timer_prep))) return rc;
    if ((rc = cand.ensure((size_t)nq * k * 8))) return rc;
    if ((rc = st.tau.ensure((size_t)nq * 4 + 16))) return rc;
    if ((rc = st.gmax.ensure((size_t)nq * M * 4))) return rc;
    if ((rc = st.coll.ensure(coll_bytes))) return rc;
    if ((rc = st.coll_cnt.ensure((size_t)parts * nq * 4))) return rc;
    (void)cand_cnt;
    // 1. sampling pass: best score of every S-th tile
    if ((rc = umma_scan(st, pl, x, xnorm, n, pitch, storage, metric, nq, k, (int)S, 2, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                        stream, nullptr, (float*)st.gmax.p, M))) return rc;
    // 2. thresholds: k-th largest of the M tile maxima
    {
        const int sortn = next_pow2(k + MERGE_THREADS);
        const size_t smem = (size_t)sortn * 8 + 16;
        PRS_CUDA(cudaFuncSetAttribute(tau_from_maxima_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tau_from_maxima_kernel<<<(unsigned)nq, MERGE_THREADS, smem, stream>>>((const float*)st.gmax.p, M, k, sortn, (float*)st.tau.p);
        PRS_LAUNCH_CHECK();
    }
    // 3. collecting pass (the dominant kernel: timed)
    if ((rc = umma_scan(st, pl, x, xnorm, n, pitch, storage, metric, nq, k, 1, 1, nullptr, nullptr, (const float*)st.tau.p, (u64*)st.coll.p,
                        (int*)st.coll_cnt.p, cap, stream, timer))) return rc;
    // 4. select into one sorted list per query: cand[1][nq][k]
    {
        const int sortn = next_pow2(k + MERGE_THREADS);
        const size_t smem = (size_t)sortn * 8 + (size_t)(parts + 2) * 4 + 16;
        PRS_CUDA(cudaFuncSetAttribute(select_collected_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        select_collected_kernel<<<(unsigned)nq, MERGE_THREADS, smem, stream>>>((const u64*)st.coll.p, (const int*)st.coll_cnt.p, parts, (int)nq, cap, k,
                                                                             sortn, (u64*)cand.p);
        PRS_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace prs
