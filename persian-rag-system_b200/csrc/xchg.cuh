// xchg.cuh -- row-sharded search: local merge + exchange over NVLink peer memory + global merge
// in ONE kernel (replaces local merge -> 2 x ncclAllGather -> merge).
//
// Every rank owns an exchange buffer (cudaMalloc, shared with the other ranks of the box through
// CUDA IPC, i.e. mapped peer memory over NVLink / NVSwitch):
//     vals, vals2 [2 slots][G ranks][cap] float      ids [2][G][cap] int64      flags [2][G][nq_cap] uint32
// CTA q of rank r merges the per-CTA candidate lists of its local scan into the local top-k of
// query q, STORES that list into slot[.][r] of every rank's buffer (plain st.global on mapped peer
// pointers, 12*k bytes per peer), fences at system scope and publishes flag = generation; then it
// waits (bounded spin, ld.acquire.sys on its OWN memory) for the G-1 other ranks' lists of the same
// query and merges the G lists.  Only 12*nq*k bytes per rank pair cross NVLink and there is no
// separate collective launch.  Two slots alternate by generation: a rank can only be two searches
// ahead of another after that rank has finished reading the older slot (its next exchange kernel
// is stream-ordered behind the read), so a slot is never overwritten while it is being read.
#pragma once
#include "common.cuh"
#include "topk_merge.cuh"

namespace prs {

constexpr int XCHG_MAX_RANKS = 16;

struct XchgView {
    float* vals[XCHG_MAX_RANKS];          // base of rank p's vals array (peer-mapped; [rank] is local)
    float* vals2[XCHG_MAX_RANKS];         // second value per entry: the direct-form L2 distance (16-bit L2 re-rank)
    long long* ids[XCHG_MAX_RANKS];
    uint32_t* flags[XCHG_MAX_RANKS];
    long long cap;                        // entries per (slot, rank)
    int nq_cap, G, rank;
#ifdef PRS_EXPERIMENTS
    int dbg;                              // PRS_XCHG_DBG: 1 do not wait for the peers' flags, 2 do not store into the peers (results are WRONG)
#endif
};
#ifdef PRS_EXPERIMENTS
#define XCHG_DBG(xv) ((xv).dbg)
#else
#define XCHG_DBG(xv) 0
#endif

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// out_mode as in merge_cand_kernel.  largest: 1 for IP, 0 for L2 (order of the exchanged values).
// rr.x != nullptr (16-bit storage, L2): the global SELECTION still runs on the expanded-form values
// -- so the selected set is the unsharded index's, whatever the partition -- while each entry also
// carries its direct-form distance (vals2), which becomes the output value and the final order.
// status: host-mapped word; set to 1 when a peer's lists did not arrive within timeout_ns -- the
// query's output is then the empty-result sentinel (ids -1), never a merge of stale lists.
// The exchange of one query has two halves, usable from one kernel (merge_xchg_kernel) or from two (the one-launch scan
// pushes in its tail, xchg_pull_kernel finishes): all NT threads of the CTA call them.
//
// PUSH: local merge over this rank's CTAs (+ direct-form distances), the local list stored into every rank's buffer,
// the flag of (slot, this rank, query) released at system scope.
template <int NT, bool STREAM = true, int ONESHOT = MERGE_ONESHOT>
__device__ __forceinline__ void xchg_push_query(
    const u64* __restrict__ cand, int parts, int nq, int k, int sortn,
    int out_mode, const float* __restrict__ qnorm, long long id_offset, const Rerank& rr,
    const XchgView& xv, uint32_t gen, int q, int tid, unsigned char* msm) {
    u64* buf = reinterpret_cast<u64*>(msm);
    u64* heads = buf + sortn;
    int* s_n = reinterpret_cast<int*>(heads + NT);
    float* dd = reinterpret_cast<float*>(s_n + 4);                 // [k]
    const int G = xv.G, slot = (int)(gen & 1u);
    const bool rerank = out_mode == 2 && rr.x != nullptr;

    // ---- 1. local merge over this rank's CTAs ----
    auto fetch = [&](long long i) -> u64 {
        const int part = (int)((unsigned)i / (unsigned)k), j = (int)i - part * k;
        return __ldcg(cand + ((size_t)part * nq + q) * k + j);
    };
    const int n = block_topk_lists<NT, STREAM, ONESHOT>(fetch, parts, k, k, buf, sortn, heads, s_n, tid);
    if (rerank) {
        __syncthreads();
        direct_l2_of_keys<NT>(buf, n, rr, q, dd, tid);
    }
    __syncthreads();
    // ---- 2. push the local list into every rank's buffer (slot, my rank, query q) ----
    const size_t ebase = ((size_t)slot * G + xv.rank) * (size_t)xv.cap + (size_t)q * k;
    for (int j = tid; j < k; j += NT) {
        float dv, d2 = 0.f;
        long long iv;
        if (j < n) {
            const u64 key = buf[j];
            const float s = key_score(key);
            dv = out_mode == 0 ? s : (out_mode == 1 ? -s : fmaxf(0.f, qnorm[q] - s));
            iv = (long long)key_id<PRS_TIE_LOW_ID>(key) + id_offset;
            if (rerank) d2 = dd[j];
        } else {
            dv = out_mode == 0 ? -3.402823466e+38f : 3.402823466e+38f;
            iv = -1;
        }
        for (int p = 0; p < G; ++p) {
            if ((XCHG_DBG(xv) & 2) && p != xv.rank) continue;
            xv.vals[p][ebase + j] = dv;
            xv.ids[p][ebase + j] = iv;
            if (rerank) xv.vals2[p][ebase + j] = d2;
        }
    }
    // bar.sync orders the CTA's stores before the flag writers; st.release.sys is cumulative, so the
    // peer that acquires the flag sees the whole list (no per-thread system fence needed)
    __syncthreads();
    if (tid < G) st_release_sys(xv.flags[tid] + ((size_t)slot * G + xv.rank) * xv.nq_cap + q, gen);
}

// PULL: bounded wait for the other ranks' lists of this query (they arrive in MY memory), then the G-way merge.
// smem: buf [sortn] | heads [NT] | s_n [4] | dd [k] | ids [k]
template <int NT, bool STREAM = true, int ONESHOT2 = MERGE_ONESHOT>
__device__ __forceinline__ void xchg_pull_query(
    int k, int sortn, int out_mode, int largest, bool rerank, const XchgView& xv, uint32_t gen, unsigned long long timeout_ns,
    float* __restrict__ D, long long* __restrict__ I, volatile int* status, int q, int tid, unsigned char* msm) {
    u64* buf = reinterpret_cast<u64*>(msm);
    u64* heads = buf + sortn;
    int* s_n = reinterpret_cast<int*>(heads + NT);                 // [0..1] block_topk_lists, [2] timeout flag
    float* dd = reinterpret_cast<float*>(s_n + 4);                 // [k]
    long long* ids_s = reinterpret_cast<long long*>(dd + ((k + 1) & ~1));   // [k]
    const int G = xv.G, slot = (int)(gen & 1u);
    if (tid == 0) s_n[2] = 0;
    __syncthreads();
    // ---- 3. wait for the other ranks' lists of this query ----
    if (tid < G) {
        const uint32_t* f = xv.flags[xv.rank] + ((size_t)slot * G + tid) * xv.nq_cap + q;
        unsigned long long t0 = 0;
        int spins = 0;
        while (ld_relaxed_sys(f) != gen) {
            if (XCHG_DBG(xv) & 1) break;
            __nanosleep(32);
            if ((++spins & 1023) == 0) {                                      // look at the clock every ~50 us
                const unsigned long long now = global_timer_ns();
                if (!t0) t0 = now;
                else if (now - t0 > timeout_ns) { s_n[2] = 1; break; }        // a peer died or never searched
            }
        }
        (void)ld_acquire_sys(f);                                              // order the list reads after the flag
    }
    __syncthreads();
    if (s_n[2]) {
        // never merge stale or partial lists: this query answers "nothing found" and the host is told
        for (int j = tid; j < k; j += NT) {
            D[(size_t)q * k + j] = largest ? -3.402823466e+38f : 3.402823466e+38f;
            I[(size_t)q * k + j] = -1;
        }
        if (tid == 0) { *status = 1; __threadfence_system(); }
        return;
    }

    // ---- 4. merge the G lists (positions order ties like the global id does: shards hold ascending blocks) ----
    const float* Dp = xv.vals[xv.rank] + (size_t)slot * G * (size_t)xv.cap;
    const float* Dp2 = xv.vals2[xv.rank] + (size_t)slot * G * (size_t)xv.cap;
    const long long* Ip = xv.ids[xv.rank] + (size_t)slot * G * (size_t)xv.cap;
    auto fetch2 = [&](long long i) -> u64 {
        const int part = (int)(i / k), j = (int)(i - (long long)part * k);
        const size_t o = (size_t)part * (size_t)xv.cap + (size_t)q * k + j;
        if (__ldcg(Ip + o) < 0) return 0ull;
        const float v = __ldcg(Dp + o);
        const float s = sanitize(largest ? v : -v);
        return ((u64)f2ord(s) << 32) | (u64)(~(uint32_t)(part * k + j));
    };
    int n2;
    if (G * k <= 128) {
        // a handful of keys (G lists of k): one warp fetches them all and sorts them in registers
        if (tid < 32) {
            int valid = 0;
            for (int e = tid; e < 128; e += 32) {
                const u64 key = e < G * k ? fetch2((long long)e) : 0ull;
                buf[e] = key;
                valid += key != 0ull;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
            __syncwarp();
            warp_sort_buf_desc(buf, G * k, k, tid);
            if (tid == 0) s_n[0] = valid < k ? valid : k;
        }
        __syncthreads();
        n2 = s_n[0];
    } else {
        n2 = block_topk_lists<NT, STREAM, ONESHOT2>(fetch2, G, k, k, buf, sortn, heads, s_n, tid);
    }
    if (rerank) {
        // selected by the expanded form; output the direct-form distances ordered by (distance, global id)
        __syncthreads();
        for (int j = tid; j < n2; j += NT) {
            const uint32_t pos = ~(uint32_t)buf[j];
            const int part = (int)(pos / k), jj = (int)(pos - (uint32_t)part * k);
            const size_t o = (size_t)part * (size_t)xv.cap + (size_t)q * k + jj;
            dd[j] = __ldcg(Dp2 + o);
            ids_s[j] = __ldcg(Ip + o);
        }
        __syncthreads();
        for (int j = tid; j < k; j += NT) {
            if (j < n2) {
                const float dj = dd[j];
                const long long idj = ids_s[j];
                int rank = 0;
                for (int i = 0; i < n2; ++i) rank += (dd[i] < dj) || (dd[i] == dj && ids_s[i] < idj);
                D[(size_t)q * k + rank] = dj;
                I[(size_t)q * k + rank] = idj;
            } else {
                D[(size_t)q * k + j] = 3.402823466e+38f;
                I[(size_t)q * k + j] = -1;
            }
        }
        return;
    }
    for (int j = tid; j < k; j += NT) {
        if (j < n2) {
            const uint32_t pos = ~(uint32_t)buf[j];
            const int part = (int)(pos / k), jj = (int)(pos - (uint32_t)part * k);
            const size_t o = (size_t)part * (size_t)xv.cap + (size_t)q * k + jj;
            D[(size_t)q * k + j] = __ldcg(Dp + o);
            I[(size_t)q * k + j] = __ldcg(Ip + o);
        } else {
            D[(size_t)q * k + j] = largest ? -3.402823466e+38f : 3.402823466e+38f;
            I[(size_t)q * k + j] = -1;
        }
    }
}

// both halves in one kernel (searches that are not one launch: clusters, wide k, the fp32 scan)
template <int NT>
__device__ __forceinline__ void merge_xchg_query(
    const u64* __restrict__ cand, int parts, int nq, int k, int sortn,
    int out_mode, const float* __restrict__ qnorm, long long id_offset, int largest, const Rerank& rr,
    const XchgView& xv, uint32_t gen, unsigned long long timeout_ns, float* __restrict__ D, long long* __restrict__ I,
    volatile int* status, int q, int tid, unsigned char* msm) {
    xchg_push_query<NT>(cand, parts, nq, k, sortn, out_mode, qnorm, id_offset, rr, xv, gen, q, tid, msm);
    __syncthreads();
    xchg_pull_query<NT>(k, sortn, out_mode, largest, out_mode == 2 && rr.x != nullptr, xv, gen, timeout_ns, D, I, status, q, tid, msm);
}

__global__ void __launch_bounds__(MERGE_THREADS) merge_xchg_kernel(
    const u64* __restrict__ cand, int parts, int nq, int k, int sortn,
    int out_mode, const float* __restrict__ qnorm, long long id_offset, int largest, const Rerank rr,
    const XchgView xv, uint32_t gen, unsigned long long timeout_ns, float* __restrict__ D, long long* __restrict__ I,
    volatile int* status, uint32_t* __restrict__ zero, long long zero_words) {
    extern __shared__ __align__(16) unsigned char msm[];
    // the scan kernel's bootstrap words, cleared for the next search when no preparation kernel will do it
    for (long long i = (long long)blockIdx.x * MERGE_THREADS + threadIdx.x; i < zero_words; i += (long long)gridDim.x * MERGE_THREADS) zero[i] = 0u;
    merge_xchg_query<MERGE_THREADS>(cand, parts, nq, k, sortn, out_mode, qnorm, id_offset, largest, rr, xv, gen, timeout_ns, D, I, status,
                                    (int)blockIdx.x, (int)threadIdx.x, msm);
}

// second launch of a row-sharded one-launch search: the scan kernel's tail has already pushed this rank's lists
constexpr int XCHG_PULL_THREADS = 64;
__global__ void __launch_bounds__(XCHG_PULL_THREADS) xchg_pull_kernel(int k, int sortn, int out_mode, int largest, int rerank, const XchgView xv,
                                                                      uint32_t gen, unsigned long long timeout_ns, float* __restrict__ D,
                                                                      long long* __restrict__ I, volatile int* status) {
    extern __shared__ __align__(16) unsigned char msm[];
    xchg_pull_query<XCHG_PULL_THREADS, true, 256>(k, sortn, out_mode, largest, rerank != 0, xv, gen, timeout_ns, D, I, status, (int)blockIdx.x,
                                                  (int)threadIdx.x, msm);
}

}  // namespace prs

// ---- host side of one rank's exchange buffer (used by flat_index.cu and by the one-launch search in flat_umma.cuh) ----
// peer-memory exchange buffer of one rank (see xchg.cuh)
struct prs_xchg {
    int device = 0, G = 1, rank = 0, nq_cap = 0;
    long long cap = 0;
    void* base = nullptr;                 // local allocation: vals | ids | flags | status
    size_t bytes = 0;
    void* peer_base[prs::XCHG_MAX_RANKS] = {};
    bool opened[prs::XCHG_MAX_RANKS] = {};
    uint32_t gen = 0;
    prs::XchgView view{};
    int* h_status = nullptr;              // page-locked, device-mapped: the kernel's timeout report, readable without a sync
    int* d_status = nullptr;              // device alias of h_status
    unsigned long long timeout_ns = 2000000000ull;
    // consecutive searches on ONE exchange context must be stream-ordered on every rank (a peer may
    // only overwrite a slot after this rank's read of it has finished): each search waits for the
    // previous one's merge kernel through this event, whatever streams the caller uses
    cudaEvent_t event = nullptr;
    bool used = false;
};

static inline size_t xchg_vals_bytes(const prs_xchg* x) { return (size_t)2 * x->G * x->cap * 4; }
static inline size_t xchg_ids_bytes(const prs_xchg* x) { return (size_t)2 * x->G * x->cap * 8; }
static inline size_t xchg_flags_bytes(const prs_xchg* x) { return (((size_t)2 * x->G * x->nq_cap * 4) + 255) & ~(size_t)255; }
static inline void xchg_fill_view(prs_xchg* x) {
    x->view.cap = x->cap; x->view.nq_cap = x->nq_cap; x->view.G = x->G; x->view.rank = x->rank;
#ifdef PRS_EXPERIMENTS
    x->view.dbg = getenv("PRS_XCHG_DBG") ? atoi(getenv("PRS_XCHG_DBG")) : 0;
#endif
    for (int p = 0; p < x->G; ++p) {
        unsigned char* b = (unsigned char*)x->peer_base[p];
        x->view.ids[p] = (long long*)b;                                  // 8-byte aligned first
        x->view.vals[p] = (float*)(b + xchg_ids_bytes(x));
        x->view.vals2[p] = (float*)(b + xchg_ids_bytes(x) + xchg_vals_bytes(x));
        x->view.flags[p] = (uint32_t*)(b + xchg_ids_bytes(x) + 2 * xchg_vals_bytes(x));
    }
}


