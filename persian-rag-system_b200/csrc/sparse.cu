// sparse.cu -- BM25 / TF-IDF scoring over an inverted index with a fused top-k.
//
// Replaces (reference file:line):
//   BM25Okapi.get_scores(query_tokens)        src/retrieval.py:127   (rank_bm25 0.2.2)
//   cosine_similarity(query_vector, matrix)   src/retrieval.py:156   (scikit-learn / scipy sparse)
//   np.argsort(scores)[::-1][:top_k]          src/retrieval.py:130,159
//
// Layout in HBM: term-major postings (CSC of the doc-by-term matrix): tptr[n_terms+1] (int64),
// pdoc[nnz] (int32 doc id, ascending inside a term), pval[nnz] (fp32 or fp64 weight).
// One CTA owns (a strided set of doc tiles) x (one query): it zeroes a float64 accumulator tile
// in shared memory, walks the query's entries IN ORDER (so the floating-point sum order per doc
// is the reference's), adds each entry's postings that fall inside the tile (coalesced reads of
// doc ids and weights, no atomics: a term lists a doc once), then streams the tile's scores --
// zero scores included, like the reference -- through a block-level top-k.  The dense score
// vector never reaches HBM.  Ties: (score desc, doc id DESC) == np.argsort(kind="stable")[::-1].
//
// Two scoring kernels share the index, the merge and the exact re-score:
//   sparse_score_kernel          "exact" mode: float64 accumulation in query-entry order (rank_bm25 /
//                                scipy summation order), one query per CTA;
//   sparse_score_batched_kernel  "throughput" mode: 8 queries per CTA share every posting they have in
//                                common (one load feeds all of them), accumulation in 32-bit FIXED POINT
//                                with native shared-memory integer atomics (order independent, hence
//                                deterministic, no barriers between query entries), per-term tile offsets
//                                precomputed at build time instead of searched.
// Both select k + SP_MARGIN candidates per query on their (rounded) selection keys; the candidates are
// then re-scored exactly in float64, ordered by (score desc, id desc) and cut to k -- so the returned
// scores are bit-exact in both modes and the rounding of the selection key can only matter when more
// than SP_MARGIN docs tie with the k-th score to ~1e-7 relative (inside the 1e-5 tolerance).
#include <algorithm>
#include <type_traits>
#include <cerrno>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"
#include "host_common.h"

namespace prs {

constexpr int SP_THREADS = 256;
constexpr int SP_TILE = 8192;        // docs per accumulator tile (64 KB of float64)
constexpr int SP_QCHUNK = 64;        // query entries whose posting ranges are resolved at once

template <typename VT>
__global__ void __launch_bounds__(SP_THREADS) sparse_score_kernel(
    const long long* __restrict__ tptr, const int* __restrict__ pdoc, const VT* __restrict__ pval,
    long long n_docs, int n_terms, const long long* __restrict__ q_indptr, const int* __restrict__ q_terms,
    const double* __restrict__ q_weights, int nq, int k, int sortn, u64* __restrict__ cand, int* __restrict__ cand_cnt) {
    extern __shared__ __align__(16) unsigned char ssm[];
    double* acc = reinterpret_cast<double*>(ssm);                       // [SP_TILE]
    u64* buf = reinterpret_cast<u64*>(ssm + (size_t)SP_TILE * 8);       // [sortn]
    long long* r_lo = reinterpret_cast<long long*>(buf + sortn);        // [SP_QCHUNK]
    long long* r_hi = r_lo + SP_QCHUNK;
    double* r_w = reinterpret_cast<double*>(r_hi + SP_QCHUNK);
    long long* r_end = reinterpret_cast<long long*>(r_w + SP_QCHUNK);
    int* s_n = reinterpret_cast<int*>(r_end + SP_QCHUNK);        // [0] buffered keys, [1] scratch

    const int tid = threadIdx.x;
    const int q = blockIdx.y;
    const long long e0 = q_indptr[q], e1 = q_indptr[q + 1];
    const long long n_tiles = (n_docs + SP_TILE - 1) / SP_TILE;

    if (tid == 0) *s_n = 0;
    u64 thr = 0;
    __syncthreads();

    // This CTA owns the CONTIGUOUS tile range [tile0, tile1): posting cursors only move forward, so
    // a tile's posting range per query entry is found with a bounded 3-probe warp search instead of
    // two full binary searches (a tile holds <= SP_TILE postings of a term: a doc lists a term once).
    const long long tiles_per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long tile0 = (long long)blockIdx.x * tiles_per;
    const long long tile1 = (tile0 + tiles_per < n_tiles) ? tile0 + tiles_per : n_tiles;
    const int warp = tid >> 5, lane = tid & 31;
    const bool fast = (e1 - e0) <= SP_QCHUNK;        // all entries of the query fit the cursor table
    if (fast) {
        const int ne = (int)(e1 - e0);
        if (tid < ne) {
            const int t = q_terms[e0 + tid];
            long long l = 0, p1 = 0;
            if (t >= 0 && t < n_terms) {
                l = tptr[t]; p1 = tptr[t + 1];
                long long r = p1;
                const long long lo0 = tile0 * SP_TILE;                  // first posting with doc >= lo0
                while (l < r) { const long long mid = (l + r) >> 1; if (pdoc[mid] < lo0) l = mid + 1; else r = mid; }
            }
            r_lo[tid] = l; r_hi[tid] = p1; r_w[tid] = q_weights[e0 + tid];   // r_lo = cursor, r_hi = end of the term's postings
        }
        __syncthreads();
    }

    for (long long tile = tile0; tile < tile1; ++tile) {
        const long long lo = tile * SP_TILE;
        const long long hi = (lo + SP_TILE < n_docs) ? lo + SP_TILE : n_docs;
        const int nd = (int)(hi - lo);
        for (int i = tid; i < nd; i += SP_THREADS) acc[i] = 0.0;
        if (fast) {
            const int ne = (int)(e1 - e0);
            // warp-cooperative bounded search: first posting in [cur, min(p1, cur + SP_TILE)] with doc >= hi
            for (int e = warp; e < ne; e += SP_THREADS / 32) {
                const long long cur = r_lo[e], p1 = r_hi[e];
                auto below = [&](long long pos) -> bool { return pos < p1 && pdoc[pos] < (int)hi; };
                long long bnd;
                const int c1 = __popc(__ballot_sync(0xffffffffu, below(cur + (long long)lane * 256)));
                if (c1 == 0) bnd = cur;
                else {
                    const long long b1 = cur + (long long)(c1 - 1) * 256;
                    const int c2 = __popc(__ballot_sync(0xffffffffu, below(b1 + 1 + (long long)lane * 8)));
                    if (c2 == 0) bnd = b1 + 1;
                    else {
                        const long long b2 = b1 + 1 + (long long)(c2 - 1) * 8;
                        const int c3 = __popc(__ballot_sync(0xffffffffu, lane < 8 && below(b2 + 1 + lane)));
                        bnd = b2 + 1 + c3;
                    }
                }
                if (lane == 0) r_end[e] = bnd;
            }
            __syncthreads();
            // Posting phase, software pipelined ACROSS entries: the loads of the next batch (possibly
            // of the next entry) are in flight while this batch is added into the accumulator tile.
            // Only the read-modify-writes of different entries are ordered (bar.sync between them),
            // which keeps every doc's float64 sum in query-entry order.  (Shared-memory float64
            // atomics without the barriers were measured 15 % slower.)
            {
                constexpr int U = 8;
                int dl[2][U];
                double v[2][U];
                int e_cur = -1, e_nxt = 0;
                long long base_nxt = 0;
                auto seek = [&]() {                      // first non-empty batch at or after (e_nxt, base_nxt)
                    while (e_nxt < ne) {
                        if (base_nxt < r_lo[e_nxt]) base_nxt = r_lo[e_nxt];
                        if (base_nxt < r_end[e_nxt]) return;
                        ++e_nxt; base_nxt = 0;
                    }
                };
                // slot is a compile-time constant so dl/v stay in registers (a runtime slot index
                // sent them to local memory: 32 registers + 192 B stack, measured 40 % slower)
                auto load = [&](auto SLOT, int e, long long base) {
                    constexpr int S = decltype(SLOT)::value;
                    const long long b = r_end[e];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const long long pp = base + (long long)u * SP_THREADS + tid;
                        dl[S][u] = pp < b ? __ldg(pdoc + pp) - (int)lo : -1;
                        v[S][u] = pp < b ? (double)__ldg(pval + pp) : 0.0;
                    }
                };
                // one pipeline step: issue the next batch into the other slot, add this slot's batch
                auto step = [&](auto SLOT) {
                    constexpr int S = decltype(SLOT)::value;
                    seek();
                    const int e_next = e_nxt < ne ? e_nxt : -1;
                    if (e_next >= 0) { load(std::integral_constant<int, S ^ 1>{}, e_next, base_nxt); base_nxt += (long long)SP_THREADS * U; }
                    const double w = r_w[e_cur];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (dl[S][u] >= 0) acc[dl[S][u]] = __dadd_rn(acc[dl[S][u]], __dmul_rn(w, v[S][u]));
                    if (e_next != e_cur) __syncthreads();   // entry boundary (uniform: every thread walks the same batches)
                    e_cur = e_next;
                };
                seek();
                if (e_nxt < ne) { load(std::integral_constant<int, 0>{}, e_nxt, base_nxt); e_cur = e_nxt; base_nxt += (long long)SP_THREADS * U; }
                while (e_cur >= 0) {
                    step(std::integral_constant<int, 0>{});
                    if (e_cur < 0) break;
                    step(std::integral_constant<int, 1>{});
                }
            }
            __syncthreads();
            if (tid < ne) r_lo[tid] = r_end[tid];           // advance the cursors
            __syncthreads();
        } else {
        __syncthreads();
        for (long long eb = e0; eb < e1; eb += SP_QCHUNK) {
            const int ne = (int)((e1 - eb < SP_QCHUNK) ? (e1 - eb) : SP_QCHUNK);
            if (tid < ne) {
                const int t = q_terms[eb + tid];
                long long a = 0, b = 0;
                if (t >= 0 && t < n_terms) {
                    const long long p0 = tptr[t], p1 = tptr[t + 1];
                    long long l = p0, r = p1;              // first posting with doc >= lo
                    while (l < r) { const long long mid = (l + r) >> 1; if (pdoc[mid] < lo) l = mid + 1; else r = mid; }
                    a = l; r = p1;                          // first posting with doc >= hi
                    while (l < r) { const long long mid = (l + r) >> 1; if (pdoc[mid] < hi) l = mid + 1; else r = mid; }
                    b = l;
                }
                r_lo[tid] = a; r_hi[tid] = b; r_w[tid] = q_weights[eb + tid];
            }
            __syncthreads();
            for (int e = 0; e < ne; ++e) {
                const long long a = r_lo[e], b = r_hi[e];
                const double w = r_w[e];
                for (long long pp = a + tid; pp < b; pp += SP_THREADS) {
                    const int dl = pdoc[pp] - (int)lo;
                    acc[dl] = __dadd_rn(acc[dl], __dmul_rn(w, (double)pval[pp]));
                }
                if (a < b) __syncthreads();                 // keep the per-doc sum order = entry order
            }
            __syncthreads();
        }
        }
        // stream this tile's scores (all docs, zero scores included) through the top-k buffer.
        // Common case (threshold already tight): count the survivors first and append them all in
        // one pass; only a tile with more survivors than the buffer holds takes the round-by-round path.
        int mine = 0;
        for (int i = tid; i < nd; i += SP_THREADS)
            mine += make_key_rt(sanitize(__double2float_rn(acc[i])), (uint32_t)(lo + i), 1) > thr;
        if (__syncthreads_count(mine > 0)) {
            if (tid == 0) s_n[1] = 0;
            __syncthreads();
            if (mine) atomicAdd(&s_n[1], mine);
            __syncthreads();
            const int total = s_n[1];
            if (*s_n + total <= sortn) {
                for (int i = tid; i < nd; i += SP_THREADS) {
                    const u64 key = make_key_rt(sanitize(__double2float_rn(acc[i])), (uint32_t)(lo + i), 1);
                    if (key > thr) buf[atomicAdd(s_n, 1)] = key;
                }
                __syncthreads();
                const int cnt = *s_n;
                if (cnt > sortn - SP_THREADS) {
                    for (int pz = cnt + tid; pz < sortn; pz += SP_THREADS) buf[pz] = 0ull;
                    __syncthreads();
                    block_sort_desc(buf, sortn, tid, SP_THREADS, 1);
                    const int keep = cnt < k ? cnt : k;
                    thr = (keep == k) ? buf[k - 1] : 0ull;
                    if (tid == 0) *s_n = keep;
                    __syncthreads();
                }
            } else {
                for (int base = 0; base < nd; base += SP_THREADS) {
                    const int i = base + tid;
                    if (i < nd) {
                        const u64 key = make_key_rt(sanitize(__double2float_rn(acc[i])), (uint32_t)(lo + i), 1);
                        if (key > thr) { const int pos = atomicAdd(s_n, 1); buf[pos] = key; }
                    }
                    __syncthreads();
                    const int cnt = *s_n;
                    if (cnt > sortn - SP_THREADS) {
                        for (int pz = cnt + tid; pz < sortn; pz += SP_THREADS) buf[pz] = 0ull;
                        __syncthreads();
                        block_sort_desc(buf, sortn, tid, SP_THREADS, 1);
                        const int keep = cnt < k ? cnt : k;
                        thr = (keep == k) ? buf[k - 1] : 0ull;
                        if (tid == 0) *s_n = keep;
                        __syncthreads();
                    }
                }
            }
        }
        __syncthreads();
    }
    int cnt = *s_n;
    __syncthreads();
    int n2 = 2;
    while (n2 < cnt) n2 <<= 1;
    for (int pz = cnt + tid; pz < n2; pz += SP_THREADS) buf[pz] = 0ull;
    __syncthreads();
    block_sort_desc(buf, n2, tid, SP_THREADS, 1);
    const int n = cnt < k ? cnt : k;
    const size_t o = (size_t)blockIdx.x * nq + q;
    for (int j = tid; j < n; j += SP_THREADS) cand[o * k + j] = buf[j];
    if (tid == 0) cand_cnt[o] = n;
}

// merge parts -> doc ids (tie: higher id wins), one CTA per query
__global__ void __launch_bounds__(SP_THREADS) sparse_merge_kernel(const u64* __restrict__ cand, const int* __restrict__ cand_cnt,
                                                                 int parts, int nq, int k, int sortn, long long* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char msm[];
    u64* buf = reinterpret_cast<u64*>(msm);
    int* s_n = reinterpret_cast<int*>(msm + (size_t)sortn * 8);
    const int q = blockIdx.x, tid = threadIdx.x;
    auto fetch = [&](long long i) -> u64 {
        const int part = (int)(i / k), j = (int)(i - (long long)part * k);
        const size_t o = (size_t)part * nq + q;
        return (j < cand_cnt[o]) ? cand[o * k + j] : 0ull;
    };
    const int n = block_topk_stream(fetch, (long long)parts * k, k, buf, sortn, s_n, tid, SP_THREADS, 1);
    for (int j = tid; j < k; j += SP_THREADS) I[(size_t)q * k + j] = (j < n) ? (long long)(uint32_t)buf[j] : -1ll;
}

// exact float64 score of each selected doc (same entry order as the scan), then order the k
// candidates of a query by (score desc, id desc) on the float64 values and keep the k_out best.
template <typename VT>
__global__ void sparse_rescore_kernel(const long long* __restrict__ tptr, const int* __restrict__ pdoc, const VT* __restrict__ pval,
                                      int n_terms, const long long* __restrict__ q_indptr, const int* __restrict__ q_terms,
                                      const double* __restrict__ q_weights, int k, int k_out, const long long* __restrict__ I_in,
                                      long long id_offset, double* __restrict__ S, long long* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char rsm[];
    double* sc = reinterpret_cast<double*>(rsm);                 // [k]
    long long* ids = reinterpret_cast<long long*>(sc + k);       // [k]
    const int q = blockIdx.x;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const long long doc = I_in[(size_t)q * k + j];
        double s = 0.0;
        if (doc >= 0) {
            for (long long e = q_indptr[q]; e < q_indptr[q + 1]; ++e) {
                const int t = q_terms[e];
                if (t < 0 || t >= n_terms) continue;
                long long l = tptr[t], r = tptr[t + 1];
                const long long end = r;
                while (l < r) { const long long mid = (l + r) >> 1; if (pdoc[mid] < doc) l = mid + 1; else r = mid; }
                if (l < end && pdoc[l] == doc) s = __dadd_rn(s, __dmul_rn(q_weights[e], (double)pval[l]));
            }
        }
        sc[j] = s; ids[j] = doc;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const double s = sc[j];
        const long long id = ids[j];
        int rank = 0;
        if (id < 0) {
            // padding keeps its place at the end
            rank = j;
        } else {
            for (int i = 0; i < k; ++i) {
                if (ids[i] < 0) continue;
                if (sc[i] > s || (sc[i] == s && ids[i] > id)) ++rank;
            }
        }
        if (rank < k_out) {                                  // k candidates (k_out + margin) in, the k_out best out
            S[(size_t)q * k_out + rank] = (id < 0) ? 0.0 : s;
            I[(size_t)q * k_out + rank] = (id < 0) ? -1ll : id + id_offset;
        }
    }
}

// ==============================================================================================
// Throughput mode
// ==============================================================================================
constexpr int SP_MARGIN = 16;          // extra candidates selected per query before the exact re-score
constexpr int SB_QG = 8;               // queries per CTA
constexpr int SB_TD = 2048;            // docs per accumulator tile
constexpr int SB_THREADS = 512;
constexpr int SB_NW = SB_THREADS / 32;
constexpr int SB_QLEN = 64;            // query entries per slot resolved per round
constexpr int SB_MAXE = SB_QG * SB_QLEN;
constexpr int SB_CAP = 128;            // candidate keys per (slot, tile half): 32 lanes x 4, compacted by a warp sort
constexpr int SB_MAXKK = 64;           // k + SP_MARGIN this kernel supports (compaction must free >= 32 places)
constexpr int SB_CHUNK = 128;          // postings per warp step (4 per lane in flight)

struct SbParams {
    const long long* tptr; const int* pdoc; const float* pval; const float* maxw;
    const int* skip_row; const uint32_t* skip; long long n_tiles;      // tiles of SB_TD docs
    long long n_docs; int n_terms;
    const long long* q_indptr; const int* q_terms; const double* q_weights;
    int nq, kk;
    u64* cand; int* cand_cnt;          // [parts][nq][kk]
};

constexpr size_t SB_SMEM = (size_t)SB_QG * SB_TD * 4 + (size_t)SB_NW * SB_CAP * 8 + (size_t)SB_MAXE * 8 /*skey*/ +
                           (size_t)SB_MAXE * 8 /*u_base*/ + (size_t)SB_MAXE * 4 * 5 /*u_len,u_cur,u_end,u_row,w0*/ +
                           (size_t)(SB_MAXE + 1) * 4 * 2 /*u_seg,pre*/ + (size_t)SB_MAXE * 4 /*e_scale*/ + SB_MAXE /*e_slot*/ + 256 /*misc*/ + (size_t)(SB_MAXE + 1) * 4 /*nz_pre*/ + SB_MAXE * 2 /*nz_u*/ + 64;

__global__ void __launch_bounds__(SB_THREADS, 2) sparse_score_batched_kernel(const SbParams p) {
    extern __shared__ __align__(16) unsigned char sbm[];
    int* acc = reinterpret_cast<int*>(sbm);                                  // [QG][TD] fixed-point scores
    u64* cbuf = reinterpret_cast<u64*>(acc + SB_QG * SB_TD);                 // [NW][CAP]
    u64* skey = cbuf + SB_NW * SB_CAP;                                       // [MAXE]
    long long* u_base = reinterpret_cast<long long*>(skey + SB_MAXE);        // [MAXE] tptr[t]
    uint32_t* u_len = reinterpret_cast<uint32_t*>(u_base + SB_MAXE);         // [MAXE] df(t)
    uint32_t* u_cur = u_len + SB_MAXE;                                       // [MAXE] tile range (relative to u_base)
    uint32_t* u_end = u_cur + SB_MAXE;
    int* u_row = reinterpret_cast<int*>(u_end + SB_MAXE);                    // [MAXE] row of the tile-offset table or -1
    float* w0 = reinterpret_cast<float*>(u_row + SB_MAXE);                   // [MAXE] entry weights before the sort
    int* u_seg = reinterpret_cast<int*>(w0 + SB_MAXE);                       // [MAXE+1] first sorted entry of unique term u
    int* pre = u_seg + SB_MAXE + 1;                                          // [MAXE+1] postings of this tile before term u
    float* e_scale = reinterpret_cast<float*>(pre + SB_MAXE + 1);            // [MAXE] sorted entries: weight * slot scale
    unsigned char* e_slot = reinterpret_cast<unsigned char*>(e_scale + SB_MAXE);   // [MAXE]
    int* misc = reinterpret_cast<int*>(e_slot + SB_MAXE);                    // [0] U, [1] E, [2..17] warp totals, [18] P, [19] chunk, [20..27] slot scale (float), [28] K
    float* s_scale = reinterpret_cast<float*>(misc + 20);
    int* s_K = misc + 28;
    int* s_dummy = misc + 32;                                                // [32]
    int* nz_pre = misc + 64;                                                 // [MAXE+1] postings before the k-th non-empty term
    unsigned short* nz_u = reinterpret_cast<unsigned short*>(nz_pre + SB_MAXE + 1);   // [MAXE] its index u

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = (int)blockIdx.x * SB_QG;                                  // first query of this group
    const int part = blockIdx.y, nparts = gridDim.y;
    const long long tiles_per = (p.n_tiles + nparts - 1) / nparts;
    const long long tile0 = (long long)part * tiles_per;
    const long long tile1 = (tile0 + tiles_per < p.n_tiles) ? tile0 + tiles_per : p.n_tiles;

    // ---- per slot: entry count, fixed-point scale 2^30 / (upper bound of |score|) ----
    int my_len = 0;                                                          // thread s < QG: entries of slot s
    if (tid < SB_QG) {
        const int q = q0 + tid;
        float scale = 0.f;
        if (q < p.nq) {
            const long long a = p.q_indptr[q], b = p.q_indptr[q + 1];
            my_len = (int)(b - a);
            double bound = 0.0;
            for (long long e = a; e < b; ++e) {
                const int t = p.q_terms[e];
                if (t >= 0 && t < p.n_terms) bound += fabs(p.q_weights[e]) * (double)p.maxw[t];
            }
            if (bound > 0.0) scale = (float)(1073741824.0 / bound);
        }
        s_scale[tid] = scale;
        misc[2 + tid] = my_len;
    }
    __syncthreads();
    int max_len = 0;
#pragma unroll
    for (int s = 0; s < SB_QG; ++s) max_len = max(max_len, misc[2 + s]);
    const int rounds = (max_len + SB_QLEN - 1) / SB_QLEN;
    const bool persistent = rounds <= 1;                                     // the usual case: cursors live across tiles
    __syncthreads();

    // ---- table of the unique terms of round r: sorted entries, segments, posting-list geometry ----
    auto build_table = [&](int r) {
        {
            const int slot = tid / SB_QLEN, j = tid % SB_QLEN + r * SB_QLEN;
            const int q = q0 + slot;
            u64 key = 0ull;
            float w = 0.f;
            if (q < p.nq) {
                const long long a = p.q_indptr[q], b = p.q_indptr[q + 1];
                if (a + j < b) {
                    const int t = p.q_terms[a + j];
                    if (t >= 0 && t < p.n_terms && p.tptr[t + 1] > p.tptr[t]) {
                        key = ((u64)(uint32_t)(t + 1) << 32) | (u64)(uint32_t)tid;
                        w = (float)p.q_weights[a + j];
                    }
                }
            }
            skey[tid] = key;
            w0[tid] = w;
        }
        __syncthreads();
        block_sort_desc(skey, SB_MAXE, tid, SB_THREADS, 1);
        // heads of the unique-term segments (sorted descending by term; empty keys last)
        const u64 key = skey[tid];
        const bool valid = key != 0ull;
        const bool head = valid && (tid == 0 || (uint32_t)(skey[tid - 1] >> 32) != (uint32_t)(key >> 32));
        const unsigned hb = __ballot_sync(0xffffffffu, head), vb = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) { misc[2 + warp] = __popc(hb); }
        if (tid == 0) { misc[1] = 0; }
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SB_NW; ++w) { const int c = misc[2 + w]; before += (w < warp) ? c : 0; total += c; }
        const int u = before + __popc(hb & ((1u << lane) - 1u));
        if (valid) {
            const int orig = (int)(uint32_t)key;
            const int slot = orig / SB_QLEN;
            e_slot[tid] = (unsigned char)slot;
            e_scale[tid] = w0[orig] * s_scale[slot];
        }
        if (lane == 0 && vb) atomicAdd(&misc[1], __popc(vb));
        if (head) {
            const int t = (int)(uint32_t)(key >> 32) - 1;
            const long long b0 = p.tptr[t];
            u_seg[u] = tid;
            u_base[u] = b0;
            u_len[u] = (uint32_t)(p.tptr[t + 1] - b0);
            u_row[u] = p.skip_row[t];
        }
        if (tid == 0) misc[0] = total;
        __syncthreads();
        if (tid == 0) u_seg[misc[0]] = misc[1];
        __syncthreads();
    };
    // first posting of term u (relative) whose doc is >= bound
    auto lower_bound_rel = [&](int u, long long bound, uint32_t l, uint32_t r) -> uint32_t {
        const int* d = p.pdoc + u_base[u];
        while (l < r) { const uint32_t mid = (l + r) >> 1; if ((long long)d[mid] < bound) l = mid + 1; else r = mid; }
        return l;
    };

    // candidate list of (slot = warp % QG, tile half = warp / QG): every warp owns one list exclusively --
    // count and admission key live in its registers, no atomics.  The two halves of a slot keep their
    // own k' best; the merge kernel unites them like two more doc-range parts.
    const int my_slot = warp % SB_QG, my_half = warp / SB_QG;
    const bool my_valid = q0 + my_slot < p.nq;
    int ccount = 0;
    u64 cthr = 0ull;
    u64* myc = cbuf + (size_t)warp * SB_CAP;
    auto compact = [&]() {                                 // warp-level: keep the kk best of this warp's list
        u64 v[SB_CAP / 32];
#pragma unroll
        for (int i = 0; i < SB_CAP / 32; ++i) { const int e = lane * (SB_CAP / 32) + i; v[i] = e < ccount ? myc[e] : 0ull; }
        __syncwarp();
        warp_sort_desc<SB_CAP / 32>(v, lane);
#pragma unroll
        for (int i = 0; i < SB_CAP / 32; ++i) myc[lane * (SB_CAP / 32) + i] = v[i];
        __syncwarp();
        ccount = ccount < p.kk ? ccount : p.kk;
        cthr = ccount == p.kk ? myc[p.kk - 1] : 0ull;
    };
    int* s_P = misc + 18;            // postings of the current tile
    int* s_chunk = misc + 19;        // next chunk to hand out

    // ---- phase A: this tile's posting range of every unique term (thread u).  Split in two so that the
    // loads (tile-offset rows / posting probes) of the NEXT tile are in flight during the posting phase
    // of the current one: range_load() returns (start, end) in registers, range_store() publishes them.
    auto range_load = [&](long long tile, long long lo, long long hi, uint32_t& c, uint32_t& e) {
        c = 0u; e = 0u;
        if (tid < misc[0]) {
            const int row = u_row[tid];
            if (row >= 0) {
                const uint32_t* sk = p.skip + (size_t)row * (size_t)(p.n_tiles + 1) + tile;
                c = __ldg(sk); e = __ldg(sk + 1);
            } else if (persistent) {
                e = u_end[tid];                                                  // descending tiles: the previous tile's start
                if (e == 0u || (long long)p.pdoc[u_base[tid] + e - 1] < lo) c = e;               // nothing in this tile
                else c = lower_bound_rel(tid, lo, e > (uint32_t)SB_TD ? e - SB_TD : 0u, e);
            } else {
                c = lower_bound_rel(tid, lo, 0u, u_len[tid]);
                e = lower_bound_rel(tid, hi, c, u_len[tid]);
            }
        }
    };
    auto range_store = [&](uint32_t c, uint32_t e) {
        if (tid < misc[0]) {
            u_cur[tid] = c;
            u_end[tid] = persistent ? c : e;                                     // persistent: becomes the next tile's end
            pre[tid + 1] = (int)(e - c);
        }
    };
    // ---- phase B: warp 0 compacts the NON-EMPTY terms of the tile (most tail terms have no posting in a
    // given 2 048-doc tile) and turns their counts into a prefix: nz_u[k] = term, pre[k] = postings before it ----
    auto prefix = [&]() {
        if (warp == 0) {
            const int U = misc[0];
            int carry = 0, kbase = 0;
            for (int b = 0; b < U; b += 32) {
                const int cnt = (b + lane < U) ? pre[b + lane + 1] : 0;
                const unsigned nzm = __ballot_sync(0xffffffffu, cnt > 0);
                int v = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
                __syncwarp();
                if (cnt > 0) {
                    const int k = kbase + __popc(nzm & ((1u << lane) - 1u));
                    nz_u[k] = (unsigned short)(b + lane);
                    nz_pre[k] = carry + v - cnt;                                 // exclusive
                }
                carry += __shfl_sync(0xffffffffu, v, 31);
                kbase += __popc(nzm);
            }
            if (lane == 0) { nz_pre[kbase] = carry; *s_P = carry; *s_chunk = 0; *s_K = kbase; }
        }
    };
    // ---- posting phase: warps grab chunks of the tile's flattened postings; one load feeds every query
    // of the group that has the term (shared-memory integer atomics, order independent) ----
    auto postings = [&](long long lo) {
        const int P = *s_P, K = *s_K;
        const int nchunks = (P + SB_CHUNK - 1) / SB_CHUNK;
        const int s1 = (K + 31) >> 5;                                            // first-level stride of the 32-ary search
        int* const dummy = s_dummy + lane;                                        // lanes without a posting add 0 here (no branch)
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(s_chunk, 1);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= nchunks) break;
            const int i0 = c * SB_CHUNK;
            const int i1 = (i0 + SB_CHUNK < P) ? i0 + SB_CHUNK : P;
            // last k with nz_pre[k] <= i0, two 32-wide probes
            int idx = lane * s1;
            const int klo = (__popc(__ballot_sync(0xffffffffu, idx < K && nz_pre[idx] <= i0)) - 1) * s1;
            idx = klo + lane;
            int k = klo + __popc(__ballot_sync(0xffffffffu, lane < s1 && idx < K && nz_pre[idx] <= i0)) - 1;
            int i = i0;
            while (i < i1) {
                const int pk = nz_pre[k], pn = nz_pre[k + 1];
                const int u = nz_u[k];
                const int segend = (i1 < pn) ? i1 : pn;
                const int n = segend - i;
                const long long g = u_base[u] + (long long)u_cur[u] + (i - pk);
                const int* dp = p.pdoc + g + lane;
                const float* vp = p.pval + g + lane;
                int dl[SB_CHUNK / 32];
                float pv[SB_CHUNK / 32];
#pragma unroll
                for (int j = 0; j < SB_CHUNK / 32; ++j) {
                    const bool in = lane + 32 * j < n;
                    dl[j] = in ? __ldg(dp + 32 * j) : -1;
                    pv[j] = in ? __ldg(vp + 32 * j) : 0.f;
                }
                const int ilo = (int)lo;
                const int e1 = u_seg[u + 1];
                for (int e = u_seg[u]; e < e1; ++e) {
                    const float sc = e_scale[e];
                    int* row = acc + (int)e_slot[e] * SB_TD - ilo;
#pragma unroll
                    for (int j = 0; j < SB_CHUNK / 32; ++j)
                        atomicAdd(dl[j] >= 0 ? row + dl[j] : dummy, __float2int_rn(pv[j] * sc));
                }
                i = segend;
                ++k;
            }
        }
    };
    // ---- selection: warp (slot, half) scans its 1024 scores, 4 per lane and step, clearing them for the
    // next tile as it goes; survivors of the integer threshold test are rare ----
    auto scan_clear = [&](long long lo, long long hi) {
        constexpr int HALF = SB_TD / 2;
        int4* row4 = reinterpret_cast<int4*>(acc + my_slot * SB_TD + my_half * HALF);
        const long long base = lo + my_half * HALF;
#pragma unroll 2
        for (int it = 0; it < HALF / 128; ++it) {
            const int i4 = it * 32 + lane;
            const int4 v = row4[i4];
            row4[i4] = make_int4(0, 0, 0, 0);
            if (!my_valid) continue;
            const int tv = cthr ? (int)((uint32_t)(cthr >> 32) ^ 0x80000000u) : (int)0x80000000;
            const bool any = v.x >= tv || v.y >= tv || v.z >= tv || v.w >= tv;
            if (__ballot_sync(0xffffffffu, any) == 0u) continue;
            const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long doc = base + (long long)i4 * 4 + j;
                const u64 key = ((u64)((uint32_t)vv[j] ^ 0x80000000u) << 32) | (u64)(uint32_t)doc;
                const bool pred = doc < hi && key > cthr;
                const unsigned m = __ballot_sync(0xffffffffu, pred);
                if (m) {
                    if (pred) myc[ccount + __popc(m & ((1u << lane) - 1u))] = key;
                    ccount += __popc(m);
                    __syncwarp();
                    if (ccount > SB_CAP - 32) compact();
                }
            }
        }
    };

    {
        int4* z = reinterpret_cast<int4*>(acc);
        for (int i = tid; i < SB_QG * SB_TD / 4; i += SB_THREADS) z[i] = make_int4(0, 0, 0, 0);
    }
    if (tid == 0) { misc[0] = 0; *s_P = 0; *s_chunk = 0; *s_K = 0; }
    if (tid < 32) s_dummy[tid] = 0;
    __syncthreads();
    auto tile_lo = [&](long long tile) { return tile * SB_TD; };
    auto tile_hi = [&](long long tile) { const long long h = tile * SB_TD + SB_TD; return h < p.n_docs ? h : p.n_docs; };
    if (rounds == 1) {
        build_table(0);
        const int U = misc[0];
        if (tid < U && u_row[tid] < 0) u_end[tid] = lower_bound_rel(tid, tile1 * SB_TD, 0u, u_len[tid]);
        __syncthreads();
        if (tile1 > tile0) {
            uint32_t c, e;
            range_load(tile1 - 1, tile_lo(tile1 - 1), tile_hi(tile1 - 1), c, e);
            range_store(c, e);
        }
        __syncthreads();
    }

    // Tiles are visited in DESCENDING doc order: a later doc then has a lower id and loses every tie
    // (score desc, id DESC), so floods of equal scores (all-zero queries) stop entering the lists once
    // they hold kk entries.
    for (long long tile = tile1 - 1; tile >= tile0; --tile) {
        const long long lo = tile_lo(tile), hi = tile_hi(tile);
        if (persistent) {
            // ranges of this tile are already in shared memory (published during the previous tile)
            uint32_t nc = 0u, ne = 0u;
            prefix();
            __syncthreads();
            if (rounds == 1) {
                if (tile > tile0) range_load(tile - 1, tile_lo(tile - 1), tile_hi(tile - 1), nc, ne);   // in flight during the posting phase
                postings(lo);
            }
            __syncthreads();
            if (rounds == 1 && tile > tile0) range_store(nc, ne);
        } else {
            for (int r = 0; r < rounds; ++r) {
                build_table(r);
                uint32_t c, e;
                range_load(tile, lo, hi, c, e);
                range_store(c, e);
                __syncthreads();
                prefix();
                __syncthreads();
                postings(lo);
                __syncthreads();
            }
        }
        scan_clear(lo, hi);
        __syncthreads();
    }
    if (my_valid) {
        compact();
        const size_t o = (size_t)(part * 2 + my_half) * p.nq + (q0 + my_slot);
        for (int j = lane; j < p.kk; j += 32) p.cand[o * p.kk + j] = j < ccount ? myc[j] : 0ull;
        if (lane == 0) p.cand_cnt[o] = ccount;
    }
}

// ---- build-time helpers of the throughput mode ----
// largest |weight| of every term (one warp per term): bounds a query's score for the fixed-point scale
template <typename VT>
__global__ void term_maxw_kernel(const long long* __restrict__ tptr, const VT* __restrict__ pval, int n_terms, float* __restrict__ maxw) {
    const int t = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= n_terms) return;
    float m = 0.f;
    for (long long i = tptr[t] + lane; i < tptr[t + 1]; i += 32) m = fmaxf(m, fabsf((float)pval[i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // round up so that float rounding of the weights can never exceed the bound
    if (lane == 0) maxw[t] = m * 1.0000002f;
}
// tile-offset rows of the frequent terms: skip[row][j] = first posting (relative) with doc >= j * SB_TD
__global__ void skip_fill_kernel(const long long* __restrict__ tptr, const int* __restrict__ pdoc, const int* __restrict__ row_term,
                                 int n_rows, long long n_tiles, uint32_t* __restrict__ skip) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_rows * (n_tiles + 1)) return;
    const int row = (int)(i / (n_tiles + 1));
    const long long j = i - (long long)row * (n_tiles + 1);
    const int t = row_term[row];
    const long long b0 = tptr[t];
    long long l = 0, r = tptr[t + 1] - b0;
    const long long bound = j * SB_TD;
    while (l < r) { const long long mid = (l + r) >> 1; if ((long long)pdoc[b0 + mid] < bound) l = mid + 1; else r = mid; }
    skip[i] = (uint32_t)l;
}
__global__ void f64_to_f32_kernel(const double* __restrict__ in, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
// postings a batch of queries touches (the algorithmic bytes of the search / 8): sum of df over its entries
__global__ void count_postings_kernel(const long long* __restrict__ tptr, int n_terms, const long long* __restrict__ q_indptr,
                                      const int* __restrict__ q_terms, int nq, unsigned long long* __restrict__ out) {
    const long long ne = q_indptr[nq];
    unsigned long long s = 0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (long long)gridDim.x * blockDim.x) {
        const int t = q_terms[e];
        if (t >= 0 && t < n_terms) s += (unsigned long long)(tptr[t + 1] - tptr[t]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, s);
}

}  // namespace prs

using namespace prs;

struct prs_sparse {
    int device = 0, sm_count = 0, vdtype = PRS_F32, mode = 0;
    long long n_docs = 0, nnz = 0, id_offset = 0, last_postings = 0;
    int n_terms = 0;
    long long* tptr = nullptr;
    int* pdoc = nullptr;
    void* pval = nullptr;
    // throughput mode (built on first use): fp32 weights, per-term max |weight|, tile-offset rows
    float* pval32 = nullptr;
    float* maxw = nullptr;
    int* skip_row = nullptr;
    uint32_t* skip = nullptr;
    long long n_tiles_td = 0;
    bool fast_ready = false;
    std::vector<long long> h_tptr;
    std::mutex mu;
    DevBuf cand, cand_cnt, qptr, qterms, qw, dI, dI2, dS;
    unsigned long long* h_count = nullptr;     // mapped host word: postings touched by the last device-side search
    unsigned long long* d_count = nullptr;
    bool count_on_device = false;
    cudaEvent_t event = nullptr;               // searches share the workspace: each waits for the previous one
    bool used = false;
};

// throughput-mode structures (needs a CUDA context on sp->device; called under sp->mu)
static int sparse_prepare_fast(prs_sparse* sp) {
    if (sp->fast_ready) return 0;
    const long long nnz = std::max<long long>(sp->nnz, 1);
    if (sp->vdtype == PRS_F64) {
        PRS_CUDA(cudaMalloc(&sp->pval32, (size_t)nnz * 4));
        f64_to_f32_kernel<<<(unsigned)((nnz + 255) / 256), 256>>>((const double*)sp->pval, sp->nnz, sp->pval32);
        PRS_LAUNCH_CHECK();
    }
    PRS_CUDA(cudaMalloc(&sp->maxw, ((size_t)sp->n_terms + 1) * 4));
    if (sp->n_terms > 0) {
        const unsigned blocks = (unsigned)(((long long)sp->n_terms * 32 + 255) / 256);
        if (sp->vdtype == PRS_F64) term_maxw_kernel<double><<<blocks, 256>>>(sp->tptr, (const double*)sp->pval, sp->n_terms, sp->maxw);
        else term_maxw_kernel<float><<<blocks, 256>>>(sp->tptr, (const float*)sp->pval, sp->n_terms, sp->maxw);
        PRS_LAUNCH_CHECK();
    }
    // tile-offset rows for terms frequent enough that a row (4 bytes per tile) is small next to their
    // postings; the threshold doubles until the table fits a quarter of the index (at least 64 MB)
    sp->n_tiles_td = (sp->n_docs + SB_TD - 1) / SB_TD;
    const size_t row_bytes = (size_t)(sp->n_tiles_td + 1) * 4;
    const size_t budget = std::max<size_t>((size_t)64 << 20, (size_t)sp->nnz * 2);
    long long min_df = std::max<long long>(4096, 2 * sp->n_tiles_td);
    std::vector<int> rows;
    for (;;) {
        rows.clear();
        for (int t = 0; t < sp->n_terms; ++t) if (sp->h_tptr[(size_t)t + 1] - sp->h_tptr[t] >= min_df) rows.push_back(t);
        if (rows.size() * row_bytes <= budget) break;
        min_df *= 2;
    }
    std::vector<int> skip_row((size_t)sp->n_terms + 1, -1);
    for (size_t r = 0; r < rows.size(); ++r) skip_row[rows[r]] = (int)r;
    PRS_CUDA(cudaMalloc(&sp->skip_row, ((size_t)sp->n_terms + 1) * 4));
    PRS_CUDA(cudaMemcpy(sp->skip_row, skip_row.data(), ((size_t)sp->n_terms + 1) * 4, cudaMemcpyHostToDevice));
    PRS_CUDA(cudaMalloc(&sp->skip, std::max<size_t>(rows.size() * row_bytes, 256)));
    if (!rows.empty()) {
        int* d_rows = nullptr;
        PRS_CUDA(cudaMalloc(&d_rows, rows.size() * 4));
        PRS_CUDA(cudaMemcpy(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice));
        const long long tot = (long long)rows.size() * (sp->n_tiles_td + 1);
        skip_fill_kernel<<<(unsigned)((tot + 255) / 256), 256>>>(sp->tptr, sp->pdoc, d_rows, (int)rows.size(), sp->n_tiles_td, sp->skip);
        PRS_LAUNCH_CHECK();
        PRS_CUDA(cudaDeviceSynchronize());
        cudaFree(d_rows);
    }
    PRS_CUDA(cudaDeviceSynchronize());
    sp->fast_ready = true;
    return 0;
}

// scoring + merge + exact re-score; all pointers on the device; asynchronous on `st` (called under sp->mu)
static int sparse_search_core(prs_sparse* sp, const long long* d_qptr, const int* d_qterms, const double* d_qw, long long nq, int k,
                              double* dS, long long* dI, cudaStream_t st) {
    int rc;
    const int kk = std::min(k + SP_MARGIN, PRS_MAX_K + SP_MARGIN);
    const bool fast = (sp->mode == 1 || (sp->mode == 2 && nq >= 2 * SB_QG)) && kk <= SB_MAXKK;     // mode 2 = by batch size
    if (fast && (rc = sparse_prepare_fast(sp))) return rc;
    if (!sp->event) PRS_CUDA(cudaEventCreateWithFlags(&sp->event, cudaEventDisableTiming));
    if (sp->used) PRS_CUDA(cudaStreamWaitEvent(st, sp->event, 0));
    int parts;
    const int sortn = next_pow2(kk + SP_THREADS);
    if (fast) {
        const long long groups = (nq + SB_QG - 1) / SB_QG;
        long long want = ((long long)sp->sm_count * 2 * 8 + groups - 1) / groups;      // ~8 waves of 2 CTAs per SM
        parts = (int)std::max<long long>(1, std::min<long long>(sp->n_tiles_td, want));
    } else {
        const long long n_tiles = (sp->n_docs + SP_TILE - 1) / SP_TILE;
        long long want = ((long long)sp->sm_count * 3 * 4 + nq - 1) / nq;              // ~4 waves of 3 CTAs per SM
        parts = (int)std::max<long long>(1, std::min<long long>(n_tiles, want));
    }
    const int lists = fast ? parts * 2 : parts;          // the batched kernel keeps one list per (part, tile half)
    if ((rc = sp->cand.ensure((size_t)lists * nq * kk * 8))) return rc;
    if ((rc = sp->cand_cnt.ensure((size_t)lists * nq * 4))) return rc;
    if ((rc = sp->dI.ensure((size_t)nq * kk * 8))) return rc;
    if (fast) {
        SbParams p;
        p.tptr = sp->tptr; p.pdoc = sp->pdoc; p.pval = sp->vdtype == PRS_F64 ? sp->pval32 : (const float*)sp->pval; p.maxw = sp->maxw;
        p.skip_row = sp->skip_row; p.skip = sp->skip; p.n_tiles = sp->n_tiles_td; p.n_docs = sp->n_docs; p.n_terms = sp->n_terms;
        p.q_indptr = d_qptr; p.q_terms = d_qterms; p.q_weights = d_qw; p.nq = (int)nq; p.kk = kk;
        p.cand = (u64*)sp->cand.p; p.cand_cnt = (int*)sp->cand_cnt.p;
        PRS_CUDA(cudaFuncSetAttribute(sparse_score_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB_SMEM));
        dim3 grid((unsigned)((nq + SB_QG - 1) / SB_QG), (unsigned)parts);
        sparse_score_batched_kernel<<<grid, SB_THREADS, SB_SMEM, st>>>(p);
        PRS_LAUNCH_CHECK();
    } else {
        const size_t smem = (size_t)SP_TILE * 8 + (size_t)sortn * 8 + SP_QCHUNK * 32 + 16;
        dim3 grid((unsigned)parts, (unsigned)nq);
        if (sp->vdtype == PRS_F64) {
            PRS_CUDA(cudaFuncSetAttribute(sparse_score_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sparse_score_kernel<double><<<grid, SP_THREADS, smem, st>>>(sp->tptr, sp->pdoc, (const double*)sp->pval, sp->n_docs, sp->n_terms,
                                                                        d_qptr, d_qterms, d_qw, (int)nq, kk, sortn, (u64*)sp->cand.p, (int*)sp->cand_cnt.p);
        } else {
            PRS_CUDA(cudaFuncSetAttribute(sparse_score_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sparse_score_kernel<float><<<grid, SP_THREADS, smem, st>>>(sp->tptr, sp->pdoc, (const float*)sp->pval, sp->n_docs, sp->n_terms,
                                                                       d_qptr, d_qterms, d_qw, (int)nq, kk, sortn, (u64*)sp->cand.p, (int*)sp->cand_cnt.p);
        }
        PRS_LAUNCH_CHECK();
    }
    {
        const size_t msmem = (size_t)sortn * 8 + 16;
        PRS_CUDA(cudaFuncSetAttribute(sparse_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
        sparse_merge_kernel<<<(unsigned)nq, SP_THREADS, msmem, st>>>((const u64*)sp->cand.p, (const int*)sp->cand_cnt.p, lists, (int)nq, kk,
                                                                     sortn, (long long*)sp->dI.p);
        PRS_LAUNCH_CHECK();
    }
    {
        const size_t rsmem = (size_t)kk * 16;
        const int nt = kk < 32 ? 32 : (kk > 256 ? 256 : (kk + 31) / 32 * 32);
        if (sp->vdtype == PRS_F64)
            sparse_rescore_kernel<double><<<(unsigned)nq, nt, rsmem, st>>>(sp->tptr, sp->pdoc, (const double*)sp->pval, sp->n_terms, d_qptr, d_qterms,
                                                                           d_qw, kk, k, (const long long*)sp->dI.p, sp->id_offset, dS, dI);
        else
            sparse_rescore_kernel<float><<<(unsigned)nq, nt, rsmem, st>>>(sp->tptr, sp->pdoc, (const float*)sp->pval, sp->n_terms, d_qptr, d_qterms,
                                                                          d_qw, kk, k, (const long long*)sp->dI.p, sp->id_offset, dS, dI);
        PRS_LAUNCH_CHECK();
    }
    PRS_CUDA(cudaEventRecord(sp->event, st));
    sp->used = true;
    return 0;
}

// doc-range shards: [nparts, nq, k] float64 (score, global id) lists -> [nq, k], ordered (score desc, then id
// DESC when tie_high, ASC otherwise).  One CTA per query, rank counting over the nparts*k candidates.
__global__ void __launch_bounds__(256) merge_parts_f64_kernel(const double* __restrict__ Sp, const long long* __restrict__ Ip, int nparts,
                                                              long long nq, int k, int tie_high, double* __restrict__ S, long long* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char fsm[];
    const int n = nparts * k;
    double* sc = reinterpret_cast<double*>(fsm);                 // [n] when it fits, else the lists are re-read from L2
    long long* id = reinterpret_cast<long long*>(sc + n);
    __shared__ int s_valid;
    const long long q = blockIdx.x;
    const int tid = threadIdx.x;
    const bool staged = n <= 4096;
    auto at = [&](int i, double& s, long long& d) {
        const int part = i / k, j = i - part * k;
        const size_t o = ((size_t)part * nq + q) * k + j;
        s = Sp[o]; d = Ip[o];
    };
    if (tid == 0) s_valid = 0;
    if (staged) for (int i = tid; i < n; i += 256) at(i, sc[i], id[i]);
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < n; i += 256) {
        double s; long long d;
        if (staged) { s = sc[i]; d = id[i]; } else at(i, s, d);
        if (d < 0) continue;
        ++mine;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            double sj; long long dj;
            if (staged) { sj = sc[j]; dj = id[j]; } else at(j, sj, dj);
            rank += dj >= 0 && (sj > s || (sj == s && (tie_high ? dj > d : dj < d)));
        }
        if (rank < k) { S[q * k + rank] = s; I[q * k + rank] = d; }
    }
    if (mine) atomicAdd(&s_valid, mine);
    __syncthreads();
    for (int j = s_valid + tid; j < k; j += 256) { S[q * k + j] = 0.0; I[q * k + j] = -1; }
}

__global__ void sparse_fill_empty_kernel(double* S, long long* I, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { S[i] = 0.0; I[i] = -1; }
}

extern "C" {

int prs_sparse_build(const int64_t* indptr, const int32_t* indices, const void* values, int vdtype, int64_t n_docs,
                     int32_t n_terms, int device, prs_sparse** out) {
    if (!out) { set_error("sparse_build: out is null"); return PRS_EINVAL; }
    *out = nullptr;
    if (!indptr || n_docs < 0 || n_terms < 0 || (vdtype != PRS_F32 && vdtype != PRS_F64)) { set_error("sparse_build: bad arguments"); return PRS_EINVAL; }
    if (n_docs > 0x7FFFFFFFll) { set_error("sparse_build: more than 2^31-1 docs per shard"); return PRS_EINVAL; }
    const long long nnz = indptr[n_docs];
    if (nnz < 0 || (nnz > 0 && (!indices || !values))) { set_error("sparse_build: bad CSR"); return PRS_EINVAL; }
    int arch = prs_device_arch(device);
    if (arch < 0) return arch;
    if (arch != 100) { set_error("libprs is built for sm_100a (B200); device %d is sm_%d", device, arch); return PRS_ECUDA; }
    prs_sparse* sp = new (std::nothrow) prs_sparse();
    if (!sp) { set_error("out of host memory"); return PRS_ENOMEM; }
    sp->device = device; sp->vdtype = vdtype; sp->n_docs = n_docs; sp->n_terms = n_terms; sp->nnz = nnz;
    // host transpose (counting sort by term; doc order inside a term stays ascending).  The scatter is
    // split over the host cores by TERM range: every thread walks the whole CSR but only places the
    // postings of its own terms, so the writes of different threads never meet and no atomics are needed.
    const size_t vs = vdtype == PRS_F64 ? 8 : 4;
    std::vector<long long>& tptr = sp->h_tptr;
    std::vector<int> pdoc;
    std::vector<unsigned char> pval;
    try {
        tptr.assign((size_t)n_terms + 1, 0);
        pdoc.resize((size_t)std::max<long long>(nnz, 1));
        pval.resize((size_t)std::max<long long>(nnz, 1) * vs);
    } catch (...) { delete sp; set_error("out of host memory"); return PRS_ENOMEM; }
    for (long long dct = 0; dct < n_docs; ++dct)
        if (indptr[dct + 1] < indptr[dct]) { delete sp; set_error("sparse_build: indptr not monotone"); return PRS_EINVAL; }
    for (long long i = 0; i < nnz; ++i) {
        const int t = indices[i];
        if (t < 0 || t >= n_terms) { delete sp; set_error("sparse_build: term id %d out of range at nnz %lld", t, i); return PRS_EINVAL; }
        tptr[(size_t)t + 1]++;
    }
    for (int t = 0; t < n_terms; ++t) tptr[(size_t)t + 1] += tptr[t];
    {
        unsigned nthr = std::thread::hardware_concurrency();
        if (nthr < 1) nthr = 1;
        if (nthr > 32) nthr = 32;
        if (nnz < (1 << 22)) nthr = 1;
        // term ranges with about equal posting counts
        std::vector<int> cut(nthr + 1, n_terms);
        cut[0] = 0;
        for (unsigned w = 1; w < nthr; ++w) {
            const long long target = nnz / nthr * w;
            cut[w] = (int)(std::lower_bound(tptr.begin(), tptr.end(), target) - tptr.begin());
            if (cut[w] > n_terms) cut[w] = n_terms;
            if (cut[w] < cut[w - 1]) cut[w] = cut[w - 1];
        }
        auto work = [&](int t_lo, int t_hi) {
            if (t_lo >= t_hi) return;
            std::vector<long long> fill(tptr.begin() + t_lo, tptr.begin() + t_hi);
            for (long long dct = 0; dct < n_docs; ++dct) {
                for (long long i = indptr[dct]; i < indptr[dct + 1]; ++i) {
                    const int t = indices[i];
                    if (t < t_lo || t >= t_hi) continue;
                    const long long pos = fill[(size_t)(t - t_lo)]++;
                    pdoc[(size_t)pos] = (int)dct;
                    memcpy(&pval[(size_t)pos * vs], (const unsigned char*)values + (size_t)i * vs, vs);
                }
            }
        };
        if (nthr == 1) work(0, n_terms);
        else {
            std::vector<std::thread> pool;
            for (unsigned w = 0; w < nthr; ++w) pool.emplace_back(work, cut[w], cut[w + 1]);
            for (auto& th : pool) th.join();
        }
    }
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete sp; set_error("cudaGetDeviceProperties failed"); return PRS_ECUDA; }
    sp->sm_count = prop.multiProcessorCount;
    bool ok = cudaMalloc(&sp->tptr, ((size_t)n_terms + 1) * 8) == cudaSuccess &&
              cudaMalloc(&sp->pdoc, (size_t)std::max<long long>(nnz, 1) * 4) == cudaSuccess &&
              cudaMalloc(&sp->pval, (size_t)std::max<long long>(nnz, 1) * vs) == cudaSuccess;
    ok = ok && cudaMemcpy(sp->tptr, tptr.data(), ((size_t)n_terms + 1) * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && nnz > 0) {
        ok = cudaMemcpy(sp->pdoc, pdoc.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
             cudaMemcpy(sp->pval, pval.data(), (size_t)nnz * vs, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    ok = ok && cudaHostAlloc((void**)&sp->h_count, 64, cudaHostAllocMapped) == cudaSuccess &&
         cudaHostGetDevicePointer((void**)&sp->d_count, sp->h_count, 0) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        set_error("sparse_build: device allocation / upload failed (%lld postings)", nnz);
        prs_sparse_free(sp);
        return PRS_ENOMEM;
    }
    *sp->h_count = 0;
    *out = sp;
    return 0;
}

void prs_sparse_free(prs_sparse* sp) {
    if (!sp) return;
    DeviceGuard g(sp->device);
    cudaDeviceSynchronize();
    if (sp->tptr) cudaFree(sp->tptr);
    if (sp->pdoc) cudaFree(sp->pdoc);
    if (sp->pval) cudaFree(sp->pval);
    if (sp->pval32) cudaFree(sp->pval32);
    if (sp->maxw) cudaFree(sp->maxw);
    if (sp->skip_row) cudaFree(sp->skip_row);
    if (sp->skip) cudaFree(sp->skip);
    if (sp->h_count) cudaFreeHost(sp->h_count);
    if (sp->event) cudaEventDestroy(sp->event);
    sp->cand.release(); sp->cand_cnt.release(); sp->qptr.release(); sp->qterms.release(); sp->qw.release();
    sp->dI.release(); sp->dI2.release(); sp->dS.release();
    delete sp;
}

int64_t prs_sparse_ndocs(const prs_sparse* sp) { return sp ? sp->n_docs : -1; }
int64_t prs_sparse_nnz(const prs_sparse* sp) { return sp ? sp->nnz : -1; }
int64_t prs_sparse_last_postings(const prs_sparse* sp) {
    if (!sp) return -1;
    if (sp->count_on_device) {                 // counted by a kernel of the last device-side search
        DeviceGuard g(sp->device);
        cudaDeviceSynchronize();
        return (int64_t)*(volatile unsigned long long*)sp->h_count;
    }
    return sp->last_postings;
}
int prs_sparse_set_id_offset(prs_sparse* sp, int64_t off) {
    if (!sp) { set_error("null sparse index"); return PRS_EINVAL; }
    sp->id_offset = off;
    return 0;
}
int prs_sparse_set_mode(prs_sparse* sp, int mode) {
    if (!sp || mode < 0 || mode > 2) { set_error("sparse_set_mode: mode must be 0 (exact order), 1 (throughput) or 2 (throughput for batches of >= 16 queries)"); return PRS_EINVAL; }
    DeviceGuard g(sp->device);
    std::lock_guard<std::mutex> lock(sp->mu);
    sp->mode = mode;
    return mode == 1 ? sparse_prepare_fast(sp) : 0;
}
int prs_sparse_mode(const prs_sparse* sp) { return sp ? sp->mode : -1; }

static int sparse_check_args(prs_sparse* sp, int64_t nq, int k) {
    if (!sp) { set_error("null sparse index"); return PRS_EINVAL; }
    if (k < 1 || k > PRS_MAX_K) { set_error("sparse_search: k=%d out of range [1, %d]", k, PRS_MAX_K); return PRS_EINVAL; }
    if (nq < 0 || nq > 65535) { set_error("sparse_search: nq=%lld out of range [0, 65535] per call", (long long)nq); return PRS_EINVAL; }
    return 0;
}

int prs_sparse_search_device(prs_sparse* sp, const int64_t* q_indptr, const int32_t* q_terms, const double* q_weights,
                             int64_t nq, int k, double* S, int64_t* I, void* stream) {
    int rc = sparse_check_args(sp, nq, k);
    if (rc) return rc;
    if (nq == 0) return 0;
    if (!q_indptr || !S || !I) { set_error("sparse_search: null pointer"); return PRS_EINVAL; }
    DeviceGuard g(sp->device);
    std::lock_guard<std::mutex> lock(sp->mu);
    cudaStream_t st = (cudaStream_t)stream;
    if (sp->n_docs == 0) {
        sparse_fill_empty_kernel<<<(unsigned)((nq * k + 255) / 256), 256, 0, st>>>(S, (long long*)I, nq * k);
        PRS_LAUNCH_CHECK();
        return 0;
    }
    PRS_CUDA(cudaMemsetAsync(sp->d_count, 0, 8, st));
    count_postings_kernel<<<64, 256, 0, st>>>(sp->tptr, sp->n_terms, (const long long*)q_indptr, q_terms, (int)nq, sp->d_count);
    PRS_LAUNCH_CHECK();
    sp->count_on_device = true;
    return sparse_search_core(sp, (const long long*)q_indptr, q_terms, q_weights, nq, k, S, (long long*)I, st);
}

int prs_merge_topk_f64_device(const double* S_parts, const int64_t* I_parts, int nparts, int64_t nq, int k, int tie_high_id,
                              double* S, int64_t* I, int device, void* stream) {
    if (nparts < 1 || nparts > 64 || nq < 0 || k < 1 || k > PRS_MAX_K) { set_error("merge_f64: bad arguments"); return PRS_EINVAL; }
    if (nq == 0) return 0;
    if (!S_parts || !I_parts || !S || !I) { set_error("merge_f64: null pointer"); return PRS_EINVAL; }
    DeviceGuard g(device);
    const int n = nparts * k;
    const size_t smem = n <= 4096 ? (size_t)n * 16 : 16;
    PRS_CUDA(cudaFuncSetAttribute(merge_parts_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_parts_f64_kernel<<<(unsigned)nq, 256, smem, (cudaStream_t)stream>>>(S_parts, (const long long*)I_parts, nparts, nq, k, tie_high_id,
                                                                             S, (long long*)I);
    PRS_LAUNCH_CHECK();
    return 0;
}

int prs_sparse_search_host(prs_sparse* sp, const int64_t* q_indptr, const int32_t* q_terms, const double* q_weights,
                           int64_t nq, int k, double* S, int64_t* I) {
    int rc = sparse_check_args(sp, nq, k);
    if (rc) return rc;
    if (nq == 0) return 0;
    if (!q_indptr || !S || !I) { set_error("sparse_search: null pointer"); return PRS_EINVAL; }
    const long long ne = q_indptr[nq];
    if (ne < 0 || (ne > 0 && (!q_terms || !q_weights))) { set_error("sparse_search: bad query CSR"); return PRS_EINVAL; }
    DeviceGuard g(sp->device);
    std::lock_guard<std::mutex> lock(sp->mu);
    if (sp->n_docs == 0) {
        for (long long i = 0; i < nq * k; ++i) { S[i] = 0.0; I[i] = -1; }
        return 0;
    }
    long long touched = 0;
    for (long long e = 0; e < ne; ++e) {
        const int t = q_terms[e];
        if (t >= 0 && t < sp->n_terms) touched += sp->h_tptr[(size_t)t + 1] - sp->h_tptr[t];
    }
    sp->last_postings = touched;
    sp->count_on_device = false;
    if ((rc = sp->qptr.ensure((size_t)(nq + 1) * 8))) return rc;
    if ((rc = sp->qterms.ensure((size_t)std::max<long long>(ne, 1) * 4))) return rc;
    if ((rc = sp->qw.ensure((size_t)std::max<long long>(ne, 1) * 8))) return rc;
    if ((rc = sp->dI2.ensure((size_t)nq * k * 8))) return rc;
    if ((rc = sp->dS.ensure((size_t)nq * k * 8))) return rc;
    PRS_CUDA(cudaMemcpyAsync(sp->qptr.p, q_indptr, (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, 0));
    if (ne > 0) {
        PRS_CUDA(cudaMemcpyAsync(sp->qterms.p, q_terms, (size_t)ne * 4, cudaMemcpyHostToDevice, 0));
        PRS_CUDA(cudaMemcpyAsync(sp->qw.p, q_weights, (size_t)ne * 8, cudaMemcpyHostToDevice, 0));
    }
    if ((rc = sparse_search_core(sp, (const long long*)sp->qptr.p, (const int*)sp->qterms.p, (const double*)sp->qw.p, nq, k,
                                 (double*)sp->dS.p, (long long*)sp->dI2.p, 0))) return rc;
    PRS_CUDA(cudaMemcpyAsync(S, sp->dS.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, 0));
    PRS_CUDA(cudaMemcpyAsync(I, sp->dI2.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, 0));
    PRS_CUDA(cudaStreamSynchronize(0));
    return 0;
}

}  // extern "C"
