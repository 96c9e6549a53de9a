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
#include <algorithm>
#include <type_traits>
#include <cerrno>
#include <mutex>
#include <new>
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
// results of a query by (score desc, id desc) on the float64 values.
template <typename VT>
__global__ void sparse_rescore_kernel(const long long* __restrict__ tptr, const int* __restrict__ pdoc, const VT* __restrict__ pval,
                                      int n_terms, const long long* __restrict__ q_indptr, const int* __restrict__ q_terms,
                                      const double* __restrict__ q_weights, int k, const long long* __restrict__ I_in,
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
        S[(size_t)q * k + rank] = (id < 0) ? 0.0 : s;
        I[(size_t)q * k + rank] = (id < 0) ? -1ll : id + id_offset;
    }
}

}  // namespace prs

using namespace prs;

struct prs_sparse {
    int device = 0, sm_count = 0, vdtype = PRS_F32;
    long long n_docs = 0, nnz = 0, id_offset = 0, last_postings = 0;
    int n_terms = 0;
    long long* tptr = nullptr;
    int* pdoc = nullptr;
    void* pval = nullptr;
    std::vector<long long> h_tptr;
    std::mutex mu;
    DevBuf cand, cand_cnt, qptr, qterms, qw, dI, dI2, dS;
};

extern "C" {

int prs_sparse_build(const int64_t* indptr, const int32_t* indices, const void* values, int vdtype, int64_t n_docs,
                     int32_t n_terms, int device, prs_sparse** out) {
    if (!out) { set_error("sparse_build: out is null"); return PRS_EINVAL; }
    *out = nullptr;
    if (!indptr || n_docs < 0 || n_terms < 0 || (vdtype != PRS_F32 && vdtype != PRS_F64)) { set_error("sparse_build: bad arguments"); return PRS_EINVAL; }
    if (n_docs > 0xFFFFFFFFll) { set_error("sparse_build: more than 2^32-1 docs per shard"); return PRS_EINVAL; }
    const long long nnz = indptr[n_docs];
    if (nnz < 0 || (nnz > 0 && (!indices || !values))) { set_error("sparse_build: bad CSR"); return PRS_EINVAL; }
    int arch = prs_device_arch(device);
    if (arch < 0) return arch;
    if (arch != 100) { set_error("libprs is built for sm_100a (B200); device %d is sm_%d", device, arch); return PRS_ECUDA; }
    prs_sparse* sp = new (std::nothrow) prs_sparse();
    if (!sp) { set_error("out of host memory"); return PRS_ENOMEM; }
    sp->device = device; sp->vdtype = vdtype; sp->n_docs = n_docs; sp->n_terms = n_terms; sp->nnz = nnz;
    // host transpose (counting sort by term; doc order inside a term stays ascending)
    const size_t vs = vdtype == PRS_F64 ? 8 : 4;
    std::vector<long long>& tptr = sp->h_tptr;
    std::vector<int> pdoc;
    std::vector<unsigned char> pval;
    try {
        tptr.assign((size_t)n_terms + 1, 0);
        pdoc.resize((size_t)std::max<long long>(nnz, 1));
        pval.resize((size_t)std::max<long long>(nnz, 1) * vs);
    } catch (...) { delete sp; set_error("out of host memory"); return PRS_ENOMEM; }
    for (long long i = 0; i < nnz; ++i) {
        const int t = indices[i];
        if (t < 0 || t >= n_terms) { delete sp; set_error("sparse_build: term id %d out of range at nnz %lld", t, i); return PRS_EINVAL; }
        tptr[(size_t)t + 1]++;
    }
    for (int t = 0; t < n_terms; ++t) tptr[(size_t)t + 1] += tptr[t];
    {
        std::vector<long long> fill(tptr.begin(), tptr.end() - 1);
        for (long long dct = 0; dct < n_docs; ++dct) {
            if (indptr[dct + 1] < indptr[dct]) { delete sp; set_error("sparse_build: indptr not monotone"); return PRS_EINVAL; }
            for (long long i = indptr[dct]; i < indptr[dct + 1]; ++i) {
                const long long pos = fill[indices[i]]++;
                pdoc[(size_t)pos] = (int)dct;
                memcpy(&pval[(size_t)pos * vs], (const unsigned char*)values + (size_t)i * vs, vs);
            }
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
    if (!ok) {
        cudaGetLastError();
        set_error("sparse_build: device allocation / upload failed (%lld postings)", nnz);
        prs_sparse_free(sp);
        return PRS_ENOMEM;
    }
    *out = sp;
    return 0;
}

void prs_sparse_free(prs_sparse* sp) {
    if (!sp) return;
    DeviceGuard g(sp->device);
    if (sp->tptr) cudaFree(sp->tptr);
    if (sp->pdoc) cudaFree(sp->pdoc);
    if (sp->pval) cudaFree(sp->pval);
    sp->cand.release(); sp->cand_cnt.release(); sp->qptr.release(); sp->qterms.release(); sp->qw.release();
    sp->dI.release(); sp->dI2.release(); sp->dS.release();
    delete sp;
}

int64_t prs_sparse_ndocs(const prs_sparse* sp) { return sp ? sp->n_docs : -1; }
int64_t prs_sparse_nnz(const prs_sparse* sp) { return sp ? sp->nnz : -1; }
int64_t prs_sparse_last_postings(const prs_sparse* sp) { return sp ? sp->last_postings : -1; }
int prs_sparse_set_id_offset(prs_sparse* sp, int64_t off) {
    if (!sp) { set_error("null sparse index"); return PRS_EINVAL; }
    sp->id_offset = off;
    return 0;
}

int prs_sparse_search_host(prs_sparse* sp, const int64_t* q_indptr, const int32_t* q_terms, const double* q_weights,
                           int64_t nq, int k, double* S, int64_t* I) {
    if (!sp) { set_error("null sparse index"); return PRS_EINVAL; }
    if (k < 1 || k > PRS_MAX_K) { set_error("sparse_search: k=%d out of range [1, %d]", k, PRS_MAX_K); return PRS_EINVAL; }
    if (nq < 0 || nq > 65535) { set_error("sparse_search: nq=%lld out of range [0, 65535] per call", (long long)nq); return PRS_EINVAL; }
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
    int rc;
    const long long n_tiles = (sp->n_docs + SP_TILE - 1) / SP_TILE;
    // enough CTAs to fill the machine, but never more parts than tiles
    long long want = ((long long)sp->sm_count * 3 * 4 + nq - 1) / nq;       // ~4 waves of 3 CTAs per SM
    if (want < 1) want = 1;
    const int parts = (int)std::min<long long>(n_tiles, want);
    if ((rc = sp->qptr.ensure((size_t)(nq + 1) * 8))) return rc;
    if ((rc = sp->qterms.ensure((size_t)std::max<long long>(ne, 1) * 4))) return rc;
    if ((rc = sp->qw.ensure((size_t)std::max<long long>(ne, 1) * 8))) return rc;
    if ((rc = sp->cand.ensure((size_t)parts * nq * k * 8))) return rc;
    if ((rc = sp->cand_cnt.ensure((size_t)parts * nq * 4))) return rc;
    if ((rc = sp->dI.ensure((size_t)nq * k * 8))) return rc;
    if ((rc = sp->dI2.ensure((size_t)nq * k * 8))) return rc;
    if ((rc = sp->dS.ensure((size_t)nq * k * 8))) return rc;
    PRS_CUDA(cudaMemcpyAsync(sp->qptr.p, q_indptr, (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, 0));
    if (ne > 0) {
        PRS_CUDA(cudaMemcpyAsync(sp->qterms.p, q_terms, (size_t)ne * 4, cudaMemcpyHostToDevice, 0));
        PRS_CUDA(cudaMemcpyAsync(sp->qw.p, q_weights, (size_t)ne * 8, cudaMemcpyHostToDevice, 0));
    }
    const int sortn = next_pow2(k + SP_THREADS);
    const size_t smem = (size_t)SP_TILE * 8 + (size_t)sortn * 8 + SP_QCHUNK * 32 + 16;
    dim3 grid((unsigned)parts, (unsigned)nq);
    if (sp->vdtype == PRS_F64) {
        PRS_CUDA(cudaFuncSetAttribute(sparse_score_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sparse_score_kernel<double><<<grid, SP_THREADS, smem, 0>>>(sp->tptr, sp->pdoc, (const double*)sp->pval, sp->n_docs, sp->n_terms,
                                                                   (const long long*)sp->qptr.p, (const int*)sp->qterms.p,
                                                                   (const double*)sp->qw.p, (int)nq, k, sortn, (u64*)sp->cand.p,
                                                                   (int*)sp->cand_cnt.p);
    } else {
        PRS_CUDA(cudaFuncSetAttribute(sparse_score_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sparse_score_kernel<float><<<grid, SP_THREADS, smem, 0>>>(sp->tptr, sp->pdoc, (const float*)sp->pval, sp->n_docs, sp->n_terms,
                                                                  (const long long*)sp->qptr.p, (const int*)sp->qterms.p,
                                                                  (const double*)sp->qw.p, (int)nq, k, sortn, (u64*)sp->cand.p,
                                                                  (int*)sp->cand_cnt.p);
    }
    PRS_LAUNCH_CHECK();
    {
        const size_t msmem = (size_t)sortn * 8 + 16;
        PRS_CUDA(cudaFuncSetAttribute(sparse_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
        sparse_merge_kernel<<<(unsigned)nq, SP_THREADS, msmem, 0>>>((const u64*)sp->cand.p, (const int*)sp->cand_cnt.p, parts, (int)nq, k,
                                                                    sortn, (long long*)sp->dI.p);
        PRS_LAUNCH_CHECK();
    }
    {
        const size_t rsmem = (size_t)k * 16;
        const int nt = k < 32 ? 32 : (k > 256 ? 256 : (k + 31) / 32 * 32);
        if (sp->vdtype == PRS_F64)
            sparse_rescore_kernel<double><<<(unsigned)nq, nt, rsmem, 0>>>(sp->tptr, sp->pdoc, (const double*)sp->pval, sp->n_terms,
                                                                          (const long long*)sp->qptr.p, (const int*)sp->qterms.p,
                                                                          (const double*)sp->qw.p, k, (const long long*)sp->dI.p,
                                                                          sp->id_offset, (double*)sp->dS.p, (long long*)sp->dI2.p);
        else
            sparse_rescore_kernel<float><<<(unsigned)nq, nt, rsmem, 0>>>(sp->tptr, sp->pdoc, (const float*)sp->pval, sp->n_terms,
                                                                         (const long long*)sp->qptr.p, (const int*)sp->qterms.p,
                                                                         (const double*)sp->qw.p, k, (const long long*)sp->dI.p,
                                                                         sp->id_offset, (double*)sp->dS.p, (long long*)sp->dI2.p);
        PRS_LAUNCH_CHECK();
    }
    PRS_CUDA(cudaMemcpyAsync(S, sp->dS.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, 0));
    PRS_CUDA(cudaMemcpyAsync(I, sp->dI2.p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, 0));
    PRS_CUDA(cudaStreamSynchronize(0));
    return 0;
}

}  // extern "C"
