/*
 * prs.h -- C ABI of libprs.so, the B200-native retrieval engine that drops in under the
 * retriever of alirezafarzipour/persian-rag-system (reference paths below are relative to
 * that repository).
 *
 * The reference is pure Python; its hot path bottoms out in three third-party calls.  Each
 * entry point here is what a binding (ctypes / cffi / pybind) for that path would bind; the
 * call it replaces is cited next to it.  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative PRS_E* code; prs_last_error() gives the
 *     thread-local message of the last failure on the calling thread;
 *   - "host" pointers are ordinary (pageable or pinned) host memory, "device" pointers are CUDA
 *     device pointers on the index's device; `stream` is a cudaStream_t passed as void* (NULL =
 *     the legacy default stream);
 *   - the library owns the device copies it makes; the caller owns every buffer it passes in;
 *   - search is re-entrant for concurrent readers (per-call workspace under an internal lock);
 *     add/write must not run concurrently with search (same contract as faiss);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     PRS_ECUDA.
 */
#ifndef PRS_H
#define PRS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes */
#define PRS_OK        0
#define PRS_EINVAL   -1   /* bad argument (dimension mismatch, k out of range, null pointer) */
#define PRS_ECUDA    -2   /* CUDA runtime / driver failure, or no device */
#define PRS_EIO      -3   /* file could not be read / written / parsed */
#define PRS_ENOMEM   -4   /* host or device allocation failed */
#define PRS_EUNSUP   -5   /* valid request this build does not implement */

/* metric: numeric values are faiss MetricType (index file field `metric_type`) */
#define PRS_METRIC_IP 0   /* faiss.METRIC_INNER_PRODUCT, fourcc IxFI */
#define PRS_METRIC_L2 1   /* faiss.METRIC_L2, fourcc IxF2 -- what the reference uses */

/* element types (corpus storage and tensor hand-off) */
#define PRS_F32  0
#define PRS_F16  1
#define PRS_BF16 2
#define PRS_F64  3        /* sparse weights only */

#define PRS_MAX_K 1024

typedef struct prs_index  prs_index;    /* flat dense index  (faiss.IndexFlatL2 / IndexFlatIP) */
typedef struct prs_sparse prs_sparse;   /* inverted sparse index (rank_bm25.BM25Okapi / TF-IDF matrix) */

const char* prs_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t prs_launch_count(void);
/* compute capability * 10 of `device` (100 for B200), or a negative error */
int prs_device_arch(int device);

/* ---------------------------------------------------------------------------------------
 * Flat dense index.
 * replaces: faiss.IndexFlatL2(dimension)            src/create_embeddings.py:130
 *           faiss.IndexFlatL2(dimension)            scripts/phase3_pdf_chunking.py:47
 * `storage` selects how rows are kept in HBM: PRS_F32 (exact-parity mode, what the reference
 * stores), PRS_F16 / PRS_BF16 (throughput mode, fp32 accumulate).
 * ------------------------------------------------------------------------------------- */
int prs_index_create(int d, int metric, int storage, int device, prs_index** out);
void prs_index_free(prs_index* idx);
/* pre-size the device buffers for `n_total` rows (avoids regrowth copies for large shards) */
int prs_index_reserve(prs_index* idx, int64_t n_total);

/* replaces: index.add(embeddings)                   src/create_embeddings.py:133
 *           index.add(batch)                        scripts/phase3_pdf_chunking.py:64 */
int prs_index_add_host(prs_index* idx, const float* x, int64_t n);
/* same, rows already on the device (dtype PRS_F32/F16/BF16, row-major [n, d], contiguous) */
int prs_index_add_device(prs_index* idx, const void* x, int dtype, int64_t n, void* stream);

int64_t prs_index_ntotal(const prs_index* idx);      /* index.ntotal   src/retrieval.py:56 */
int     prs_index_d(const prs_index* idx);           /* index.d        src/create_embeddings.py:286 */
int     prs_index_metric(const prs_index* idx);
int     prs_index_storage(const prs_index* idx);

/* replaces: faiss_index.search(query_embedding, top_k)   src/retrieval.py:102
 *           index.search(test_vector, 1)                 src/create_embeddings.py:291
 * q: [nq, d] float32 row-major.  D: [nq, k] float32 (squared L2 ascending, or inner product
 * descending).  I: [nq, k] int64 row ids, -1 (and +/-FLT_MAX in D) where fewer than k rows exist.
 * Ties are ordered (value, id ascending). */
int prs_index_search_host(prs_index* idx, const float* q, int64_t nq, int k, float* D, int64_t* I);
/* same with q, D, I on the device; qdtype PRS_F32/F16/BF16.  Asynchronous on `stream`. */
int prs_index_search_device(prs_index* idx, const void* q, int qdtype, int64_t nq, int k,
                            float* D, int64_t* I, void* stream);
/* id_offset is added to every returned row id (row-sharded corpora: global id = local + offset) */
int prs_index_set_id_offset(prs_index* idx, int64_t id_offset);
/* force a kernel family: 0 = automatic, 1 = CUDA-core scan, 2 = tcgen05 scan (tests/bench) */
int prs_index_set_path(prs_index* idx, int path);
/* which family the last search on this index used (1 or 2), and its main kernel's name */
int prs_index_last_path(const prs_index* idx);
/* One-launch search (default on): on the tcgen05 path with nq <= 128 and k <= 16 the query preparation, the scan
 * and the merge run as ONE cooperative kernel -- the epilogue threads convert their own query rows, a grid barrier
 * replaces the kernel boundaries, and the CTAs merge the queries among themselves.  (Host-mapped or peer-resident
 * queries keep the preparation kernel, which reads them exactly once.)  0 restores the three-kernel sequence (A/B
 * measurements).  Row-sharded searches are two launches by default (the scan prepares its own queries, the merge + exchange
 * kernel follows); 2 makes them push their lists to the peers in the scan kernel's tail and finish with one small
 * wait-and-merge kernel (correct, measured slightly slower at N = 2); 3 asks for the two-launch form on an unsharded index.
 * prs_index_last_fused: 1 if the last search was one launch. */
int prs_index_set_fused(prs_index* idx, int enable);
int prs_index_last_fused(const prs_index* idx);

/* bench instrumentation (bench.py's roofline): while enabled, every launch of the scan kernel
 * (the dominant kernel of a search) is bracketed by CUDA events on the launching stream.
 * prs_index_scan_time synchronises them, returns the summed device time and the number of
 * launches since the previous call, and resets both. */
int prs_index_set_timing(prs_index* idx, int enable);
int prs_index_scan_time(prs_index* idx, double* total_ms, int64_t* launches);
/* same instrumentation for the two small kernels around the scan: summed device time of the
 * query-preparation kernel and of the merge (or merge+exchange) kernel since the last call */
int prs_index_phase_times(prs_index* idx, double* prep_ms, double* merge_ms);

/* copy rows [i0, i0+n) back as float32 (index.reconstruct_n) */
int prs_index_reconstruct_host(prs_index* idx, int64_t i0, int64_t n, float* out);

/* replaces: faiss.write_index(index, index_file)    src/create_embeddings.py:136
 *           faiss.read_index(faiss_index_file)      src/retrieval.py:55, src/create_embeddings.py:284
 * fp32 storage reads/writes faiss's byte-exact IndexFlat file (fourcc IxF2 / IxFI).  fp16/bf16
 * storage writes a tagged container (fourcc PRSh / PRSb: the same 45-byte header followed by the
 * 16-bit rows) that only this library reads.  prs_index_read converts to `storage` on load. */
int prs_index_write(prs_index* idx, const char* path);
int prs_index_read(const char* path, int storage, int device, prs_index** out);

/* Sharded container: one file per shard = [128-byte header | corpus exactly as it sits in HBM
 * (16-bit: T64 blocks; fp32: row-major) | float32 squared norms], page-aligned sections (mmap-able).  Loading is a
 * straight double-buffered copy through page-locked staging (one cudaMemcpyAsync per 64 MB chunk, file reads
 * overlapping the copies): no conversion kernel, no norm recomputation.  The reference's anchor is the single
 * faiss file at src/create_embeddings.py:136 / src/retrieval.py:55; this is what a 307 GB corpus needs instead.
 * prs_index_read_shard: *out == NULL -> new index; otherwise the shard is APPENDED to *out (consecutive shards of
 * one rank; 16-bit images concatenate at 64-row boundaries only).  The manifest (which shard goes to which rank,
 * global ids) is kept by the host layer (container.py). */
int prs_index_write_shard(prs_index* idx, const char* path);
int prs_index_read_shard(const char* path, int device, prs_index** out);

/* ---------------------------------------------------------------------------------------
 * Row-sharded search: merge step.  Each of `nparts` shards contributes its local
 * top-k for the same nq queries, concatenated as D_parts/I_parts [nparts, nq, k] on the device
 * (what ncclAllGather of the per-rank results produces, in rank order; shards hold ascending
 * contiguous row blocks).  Writes the global top-k.  largest: 1 for IP / sparse scores, 0 for L2.
 * tie_high_id: 0 -> ties by id ascending (dense), 1 -> id descending (sparse path).
 * ------------------------------------------------------------------------------------- */
int prs_merge_topk_device(const float* D_parts, const int64_t* I_parts, int nparts, int64_t nq, int k,
                          int largest, int tie_high_id, float* D, int64_t* I, int device, void* stream);

/* same for the sparse path ("shard by doc range the same way"): float64 scores, always largest
 * first; tie_high_id = 1 reproduces the unsharded order (score desc, id desc). */
int prs_merge_topk_f64_device(const double* S_parts, const int64_t* I_parts, int nparts, int64_t nq, int k,
                              int tie_high_id, double* S, int64_t* I, int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Row-sharded search with the exchange fused into the merge kernel (one process per GPU).
 * Each rank creates an exchange buffer, publishes its CUDA IPC handle (prs_xchg_handle_bytes()
 * bytes), gathers the handles of all ranks of the box (any transport: torch.distributed, MPI, a
 * file) and opens them; the buffers are then mapped peer memory over NVLink.  A sharded search is
 * the local scan followed by ONE kernel that merges the local per-CTA lists, stores the local
 * top-k into every peer's buffer, waits for the peers' lists and merges them: no collective
 * launch, 12*nq*k bytes per rank pair.  It is a collective: every rank must call it the same
 * number of times with the same nq and k (a rank whose shard is empty contributes empty lists).  Results are identical
 * on every rank and identical to the unsharded index (ties on global ids).
 * nq_cap / k_cap bound the nq*k of one search (buffer = 2*n_ranks*nq_cap*k_cap*12 bytes).
 * ------------------------------------------------------------------------------------- */
typedef struct prs_xchg prs_xchg;
int prs_xchg_create(int device, int n_ranks, int rank, int64_t nq_cap, int k_cap, prs_xchg** out);
int prs_xchg_handle_bytes(void);
int prs_xchg_get_handle(prs_xchg* x, void* handle_out);
/* handles: n_ranks consecutive handles in rank order (this rank's own entry is ignored) */
int prs_xchg_open_peers(prs_xchg* x, const void* handles);
/* 0, or PRS_ECUDA if a search timed out waiting for a peer (synchronises the device, then resets the
 * report).  A query whose peer lists did not arrive returns the empty-result sentinel (ids -1), never a
 * merge of stale lists; the report also surfaces, without a synchronisation, as the return code of the
 * NEXT prs_index_search_sharded_device call on the same exchange context. */
int prs_xchg_status(prs_xchg* x);
/* how long a search waits for a peer's lists before giving up (default 2000 ms) */
int prs_xchg_set_timeout_ms(prs_xchg* x, int64_t ms);
void prs_xchg_free(prs_xchg* x);
int prs_index_search_sharded_device(prs_index* idx, prs_xchg* x, const void* q, int qdtype, int64_t nq, int k,
                                    float* D, int64_t* I, void* stream);

/* link exchange buffers created by ONE process on different devices (peer access enabled between them): the
 * in-process alternative to prs_xchg_get_handle / prs_xchg_open_peers.  xs[i] must have rank i of n. */
int prs_xchg_link_local(prs_xchg** xs, int n);

/* ---------------------------------------------------------------------------------------
 * One process, several GPUs (`devices=[...]`).  The reference's retriever is a single Python
 * process (src/retrieval.py:13; scripts/gradio_luncher.py:354-362 shares one instance between threads); a
 * prs_group lets that process use every GPU of the box: one shard index, one exchange buffer and one host
 * worker thread per device, peer access between all pairs, the same fused merge + NVLink exchange kernel as
 * the one-process-per-GPU layout.  Rows are dealt in contiguous ascending blocks of ceil(n_total / ndev) rows
 * (call prs_group_reserve first; the first add fixes the block size otherwise), results are identical to the
 * single index.  nq_cap / k_cap bound one exchange (larger batches are chunked).  q / D / I of the device entry
 * point live on devs[0] (or in page-locked host memory); `stream` is a stream of devs[0].
 * ------------------------------------------------------------------------------------- */
typedef struct prs_group prs_group;
int prs_group_create(int d, int metric, int storage, const int* devs, int ndev, int64_t nq_cap, int k_cap, prs_group** out);
void prs_group_free(prs_group* g);
int prs_group_ndev(const prs_group* g);
int64_t prs_group_ntotal(const prs_group* g);
int prs_group_d(const prs_group* g);
int prs_group_metric(const prs_group* g);
int prs_group_storage(const prs_group* g);
int64_t prs_group_shard_rows(const prs_group* g, int i);
int prs_group_reserve(prs_group* g, int64_t n_total);
int prs_group_add_host(prs_group* g, const float* x, int64_t n);
int prs_group_add_device(prs_group* g, const void* x, int dtype, int64_t n, void* stream);
int prs_group_search_host(prs_group* g, const float* q, int64_t nq, int k, float* D, int64_t* I);
int prs_group_search_device(prs_group* g, const void* q, int qdtype, int64_t nq, int k, float* D, int64_t* I, void* stream);
int prs_group_reconstruct_host(prs_group* g, int64_t i0, int64_t n, float* out);

/* ---------------------------------------------------------------------------------------
 * Sparse scoring (BM25 / TF-IDF) over an inverted index.
 * replaces: BM25Okapi(tokenized_chunks)             src/retrieval.py:67   (build)
 *           bm25_index.get_scores(query_tokens)     src/retrieval.py:127  + np.argsort(...)[::-1][:k] :130
 *           cosine_similarity(query_vector, matrix) src/retrieval.py:156  + np.argsort(...)[::-1][:k] :159
 * The caller supplies the doc-by-term matrix in CSR (indptr [n_docs+1] int64, indices [nnz]
 * int32 term ids, values [nnz] of `vdtype` PRS_F32 or PRS_F64): BM25 per-posting weights
 * idf(t)*tf*(k1+1)/(tf+k1*(1-b+b*dl/avgdl)), or l2-normalised tf-idf.  The library transposes it
 * into term-major postings on the device.
 * ------------------------------------------------------------------------------------- */
int prs_sparse_build(const int64_t* indptr, const int32_t* indices, const void* values, int vdtype,
                     int64_t n_docs, int32_t n_terms, int device, prs_sparse** out);
void prs_sparse_free(prs_sparse* sp);
int64_t prs_sparse_ndocs(const prs_sparse* sp);
int64_t prs_sparse_nnz(const prs_sparse* sp);
/* queries in CSR: q_indptr [nq+1] int64, q_terms [.] int32 term ids (repeats allowed, order kept,
 * ids outside [0, n_terms) are ignored like unknown words), q_weights [.] float64 (1.0 for BM25).
 * score(doc) = sum over the query's entries IN ORDER of q_weight * posting weight, accumulated in
 * float64.  Returns the k best docs per query ordered (score desc, id DESC) -- the order of
 * np.argsort(scores, kind="stable")[::-1] -- including zero-score docs, as the reference does.
 * S: [nq, k] float64 scores, I: [nq, k] int64 doc ids (-1 padded).  Host pointers. */
int prs_sparse_search_host(prs_sparse* sp, const int64_t* q_indptr, const int32_t* q_terms,
                           const double* q_weights, int64_t nq, int k, double* S, int64_t* I);
/* same with the query CSR and the outputs on the device; asynchronous on `stream` (searches on one index are
 * ordered one after the other: they share its workspace).  This is what the hybrid fusion consumes. */
int prs_sparse_search_device(prs_sparse* sp, const int64_t* q_indptr, const int32_t* q_terms,
                             const double* q_weights, int64_t nq, int k, double* S, int64_t* I, void* stream);
/* scoring kernel: 0 = exact order (default: float64 accumulation in query-entry order, the summation order of
 * rank_bm25 / scipy), 1 = throughput (queries batched per CTA share their postings, fixed-point accumulation
 * with shared-memory integer atomics, k + 16 candidates).  Both re-score their candidates exactly in float64,
 * so the returned scores are bit-identical; the id lists can differ only among docs whose scores tie to ~1e-7.
 * 2 = automatic: exact order for single queries and small batches (the reference's call shape), throughput for
 * batches of 16 queries or more (its structures are built on the first such batch). */
int prs_sparse_set_mode(prs_sparse* sp, int mode);
int prs_sparse_mode(const prs_sparse* sp);
/* sum of postings touched by the last search (8 or 12 bytes each): the algorithmic bytes */
int64_t prs_sparse_last_postings(const prs_sparse* sp);
int prs_sparse_set_id_offset(prs_sparse* sp, int64_t id_offset);

/* ---------------------------------------------------------------------------------------
 * Hybrid fusion (dense + BM25), batched, on the device.
 * replaces: the fusion loop of retrieve_hybrid        src/retrieval.py:181-216
 * Inputs are the two top-2k lists of every query as the search entry points above leave them on the
 * device: D_dense/I_dense [nq, kd] (squared L2 ascending -> similarity 1/(1+d), src/retrieval.py:108) and
 * S_sparse/I_sparse [nq, ks] (BM25 scores).  Row ids outside [0, n_chunks) are dropped (:106).  Each list is
 * divided by its own maximum (0 if that is not positive), weighted, summed per row id; the union is ordered
 * by fused score descending, ties in insertion order (dense hits first, then BM25-only hits -- Python's
 * stable sort over the reference's dict), and cut to top_k.  S_out [nq, top_k] float64, I_out [nq, top_k]
 * int64 (-1 padded).  Asynchronous on `stream`.
 * ------------------------------------------------------------------------------------- */
int prs_hybrid_fuse_device(const float* D_dense, const int64_t* I_dense, int kd, const double* S_sparse,
                           const int64_t* I_sparse, int ks, int64_t nq, int64_t n_chunks, double dense_weight,
                           double sparse_weight, int top_k, double* S_out, int64_t* I_out, int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * IVF-Flat (the reference's branch for >= 1000 embeddings, scripts/phase3_pdf_chunking.py:45-57:
 * faiss.IndexIVFFlat(faiss.IndexFlatL2(d), d, nlist); index.train(first 10 000 rows); default nprobe = 1).
 * Nearest-centroid assignment, the coarse probe and the scan of an inverted list are exact flat searches and use
 * the prs_index_* entry points (host logic: ivf.py).  This is the remaining piece: the k-means centroid update,
 * centroids[c] = mean of the rows assigned to c, summed in row order in float32 like faiss's compute_centroids
 * (bit-identical to a sequential restatement).  x [n, d] float32, assign [n] int64 in [0, k), centroids [k, d]
 * in/out (an empty cluster keeps its centroid), counts [k] out -- all on the device.  Asynchronous on `stream`.
 * ------------------------------------------------------------------------------------- */
int prs_centroid_update_device(const float* x, int64_t n, int d, const int64_t* assign, int k, float* centroids,
                               int64_t* counts, int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Encoder output epilogue: attention-masked mean pooling (+ optional L2 normalisation).
 * replaces the tail of SentenceTransformer.encode  src/retrieval.py:98, src/create_embeddings.py:97-101
 *   out[b,:] = sum_t hidden[b,t,:]*mask[b,t] / max(sum_t mask[b,t], 1e-9);  if normalize:
 *   out[b,:] /= max(||out[b,:]||_2, 1e-12)
 * hidden: [B,T,H] device, dtype PRS_F32/F16/BF16; mask: [B,T] device int64 (0/1); out: [B,H]
 * device float32.  Asynchronous on `stream`.
 * ------------------------------------------------------------------------------------- */
int prs_pool_norm(const void* hidden, int dtype, const int64_t* mask, int B, int T, int H,
                  int normalize, float* out, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PRS_H */
