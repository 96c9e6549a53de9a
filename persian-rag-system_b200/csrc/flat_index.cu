// flat_index.cu -- host side of the flat dense index behind the C ABI in include/prs.h.
//
// Mirrors the faiss objects the reference touches (IndexFlatL2 at src/create_embeddings.py:130,
// add :133, write_index :136, read_index src/retrieval.py:55, search src/retrieval.py:102) and
// dispatches to the sm_100a kernels in flat_simt.cuh / flat_umma.cuh.  No CPU compute path.
#include <mutex>
#include <new>
#include <vector>

#include <algorithm>
#include <cerrno>

#include "flat_simt_launch.h"
#include "flat_umma.cuh"
#include "host_common.h"
#include "topk_merge.cuh"
#include "xchg.cuh"

namespace prs {

thread_local std::string t_error;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}

static inline int elem_size(int dt) { return dt == PRS_F32 ? 4 : (dt == PRS_F64 ? 8 : 2); }

// ------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// one warp per row: convert [n, d] -> storage rows [n, pitch] (zero padded) and write the squared
// norm of the STORED (rounded) row, which is what the expanded L2 form needs.
template <typename TI, typename TO>
__global__ void ingest_rows_kernel(const TI* __restrict__ in, long long n, int d, int pitch,
                                   TO* __restrict__ out, float* __restrict__ norm) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float acc = 0.f;
    for (int c = lane; c < pitch; c += 32) {
        TO o = from_f32<TO>(0.f);
        if (c < d) o = from_f32<TO>(to_f32<TI>(in[row * d + c]));
        out[row * pitch + c] = o;
        const float f = to_f32<TO>(o);
        acc = fmaf(f, f, acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) norm[row] = acc;
}

// 16-bit storage: same, but rows go into the T64 layout (common.cuh).  One warp per row, one
// 16-byte chunk (8 elements) per lane and step.
template <typename TI, typename TO>
__global__ void ingest_rows_t64_kernel(const TI* __restrict__ in, long long n, int d, int pitch, long long row0,
                                       unsigned char* __restrict__ out, float* __restrict__ norm) {
    const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const long long row = row0 + i;
    float acc = 0.f;
    for (int ch = lane; ch < (pitch >> 3); ch += 32) {
        __align__(16) TO o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = ch * 8 + e;
            o[e] = from_f32<TO>(c < d ? to_f32<TI>(in[i * d + c]) : 0.f);
            const float f = to_f32<TO>(o[e]);
            acc = fmaf(f, f, acc);
        }
        *reinterpret_cast<uint4*>(out + t64_offset(row, ch, pitch)) = *reinterpret_cast<const uint4*>(o);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) norm[row] = acc;
}

template <typename TI>
__global__ void to_f32_kernel(const TI* __restrict__ in, long long n, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = to_f32<TI>(in[i]);
}

template <typename TS>
__global__ void reconstruct_kernel(const TS* __restrict__ x, long long i0, long long n, int d, int pitch,
                                   float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * d) return;
    const long long r = i / d;
    const int c = (int)(i - r * d);
    if (sizeof(TS) == 4) out[i] = to_f32<TS>(x[(i0 + r) * pitch + c]);
    else out[i] = to_f32<TS>(*reinterpret_cast<const TS*>(reinterpret_cast<const unsigned char*>(x) + t64_offset(i0 + r, c >> 3, pitch) + (c & 7) * 2));
}

__global__ void fill_empty_kernel(float* D, long long* I, long long n, float dv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { D[i] = dv; I[i] = -1; }
}

// ------------------------------------------------------------------------------------------
// row-sharded merge: [nparts, nq, k] (score, id) lists -> [nq, k]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MERGE_THREADS) merge_parts_kernel(
    const float* __restrict__ Dp, const long long* __restrict__ Ip, int nparts, long long nq, int k, int sortn,
    int largest, int tie_high, float* __restrict__ D, long long* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char msm[];
    u64* buf = reinterpret_cast<u64*>(msm);
    u64* heads = buf + sortn;
    int* s_n = reinterpret_cast<int*>(heads + MERGE_THREADS);
    const long long q = blockIdx.x;
    const int tid = threadIdx.x;
    // position p = part * k + j.  Within a part equal scores are already ordered by the tie
    // rule and parts hold ascending row blocks, so ordering ties by position reproduces the
    // global-id tie rule without needing 64-bit ids in the key.
    auto fetch = [&](long long i) -> u64 {
        const int part = (int)(i / k), j = (int)(i - (long long)part * k);
        const size_t o = ((size_t)part * nq + q) * k + j;
        if (Ip[o] < 0) return 0ull;
        const float s = sanitize(largest ? Dp[o] : -Dp[o]);
        const uint32_t lo = tie_high ? (uint32_t)(part * k + (k - 1 - j)) : ~(uint32_t)(part * k + j);
        return ((u64)f2ord(s) << 32) | lo;
    };
    const int n = block_topk_lists<MERGE_THREADS>(fetch, nparts, k, k, buf, sortn, heads, s_n, tid);
    for (int j = tid; j < k; j += MERGE_THREADS) {
        if (j < n) {
            const uint32_t lo = (uint32_t)buf[j];
            int part, jj;
            if (tie_high) { part = (int)(lo / k); jj = k - 1 - (int)(lo - (uint32_t)part * k); }
            else { const uint32_t p = ~lo; part = (int)(p / k); jj = (int)(p - (uint32_t)part * k); }
            const size_t o = ((size_t)part * nq + q) * k + jj;
            D[q * k + j] = Dp[o];
            I[q * k + j] = Ip[o];
        } else {
            D[q * k + j] = largest ? -3.402823466e+38f : 3.402823466e+38f;
            I[q * k + j] = -1;
        }
    }
}

}  // namespace prs

using namespace prs;

// ------------------------------------------------------------------------------------------
// the index object
// ------------------------------------------------------------------------------------------
struct prs_index {
    int d = 0, pitch = 0, metric = PRS_METRIC_L2, storage = PRS_F32, device = 0, sm_count = 0;
    long long n = 0, cap_rows = 0;
    void* x = nullptr;
    float* xnorm = nullptr;
    long long id_offset = 0;
    int path_force = 0, last_path = 0;
    int l2_rerank = 1;                   // 16-bit storage, L2: direct-form distances for the selected rows (see topk_merge.cuh)
    int fuse = 1, last_fused = 0;        // one-launch search (prep + scan + merge in one cooperative kernel) when the shape allows
    std::mutex mu, host_mu;
    // Search workspaces.  Two slots used round-robin so that two searches can be IN FLIGHT on two
    // streams (the merge / exchange kernel of one overlaps the scan of the next); a slot is reused
    // only after the search that last used it has finished (event wait when the stream differs).
    struct Workspace {
        DevBuf lists, cand, cand_cnt, qf32, qnorm;
        UmmaState umma;
        cudaEvent_t event = nullptr;
        cudaStream_t stream = nullptr;
        bool used = false;
    };
    static constexpr int NSLOT = 2;
    Workspace slot[NSLOT];
    Workspace* cur = &slot[0];          // workspace of the search being issued (set under mu)
    unsigned next_slot = 0;
    DevBuf hD, hI, hQ, stage;
    PinnedBuf pQ, pD, pI;                // page-locked staging of small pageable host calls
    ScanTimer timer, timer_prep, timer_merge;
};
static thread_local struct prs_xchg* t_xchg = nullptr;   // set by the calling thread for the duration of a sharded search

static int index_grow(prs_index* idx, long long n_total) {
    if (idx->storage != PRS_F32) n_total = (n_total + BLK_ROWS - 1) / BLK_ROWS * BLK_ROWS;   // whole T64 blocks
    if (n_total <= idx->cap_rows) return 0;
    const size_t es = elem_size(idx->storage);
    void* nx = nullptr;
    float* nn = nullptr;
    const size_t xb = (size_t)n_total * idx->pitch * es;
    cudaError_t e = cudaMalloc(&nx, xb ? xb : 256);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu bytes) for corpus failed: %s", xb, cudaGetErrorString(e)); return PRS_ENOMEM; }
    e = cudaMalloc(&nn, (size_t)n_total * 4 + 256);
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(nx); set_error("cudaMalloc for norms failed: %s", cudaGetErrorString(e)); return PRS_ENOMEM; }
    if (idx->storage != PRS_F32) PRS_CUDA(cudaMemset(nx, 0, xb ? xb : 256));   // rows of a partial block read as zeros
    if (idx->n > 0) {
        const long long used = idx->storage != PRS_F32 ? (idx->n + BLK_ROWS - 1) / BLK_ROWS * BLK_ROWS : idx->n;
        PRS_CUDA(cudaMemcpy(nx, idx->x, (size_t)used * idx->pitch * es, cudaMemcpyDeviceToDevice));
        PRS_CUDA(cudaMemcpy(nn, idx->xnorm, (size_t)idx->n * 4, cudaMemcpyDeviceToDevice));
    }
    if (idx->x) cudaFree(idx->x);
    if (idx->xnorm) cudaFree(idx->xnorm);
    idx->x = nx; idx->xnorm = nn; idx->cap_rows = n_total;
    return 0;
}

template <typename TI>
static int ingest_dispatch(prs_index* idx, const void* x, long long n, cudaStream_t st) {
    const int wpb = 8;
    const unsigned grid = (unsigned)((n + wpb - 1) / wpb);
    const size_t es = elem_size(idx->storage);
    unsigned char* dst = (unsigned char*)idx->x + (size_t)idx->n * idx->pitch * es;
    float* nrm = idx->xnorm + idx->n;
    switch (idx->storage) {
        case PRS_F32: ingest_rows_kernel<TI, float><<<grid, wpb * 32, 0, st>>>((const TI*)x, n, idx->d, idx->pitch, (float*)dst, nrm); break;
        case PRS_F16: ingest_rows_t64_kernel<TI, __half><<<grid, wpb * 32, 0, st>>>((const TI*)x, n, idx->d, idx->pitch, idx->n, (unsigned char*)idx->x, idx->xnorm); break;
        default: ingest_rows_t64_kernel<TI, __nv_bfloat16><<<grid, wpb * 32, 0, st>>>((const TI*)x, n, idx->d, idx->pitch, idx->n, (unsigned char*)idx->x, idx->xnorm); break;
    }
    PRS_LAUNCH_CHECK();
    return 0;
}

static int add_device_impl(prs_index* idx, const void* x, int dtype, long long n, cudaStream_t st) {
    if (n == 0) return 0;
    if (idx->n + n > idx->cap_rows) {
        // growth copies run on the legacy stream; make sure earlier async work is done
        PRS_CUDA(cudaStreamSynchronize(st));
        long long want = idx->n + n;
        if (idx->cap_rows > 0) { long long g = idx->cap_rows + idx->cap_rows / 2; if (g > want) want = g; }
        int rc = index_grow(idx, want);
        if (rc) return rc;
    }
    int rc;
    switch (dtype) {
        case PRS_F32: rc = ingest_dispatch<float>(idx, x, n, st); break;
        case PRS_F16: rc = ingest_dispatch<__half>(idx, x, n, st); break;
        case PRS_BF16: rc = ingest_dispatch<__nv_bfloat16>(idx, x, n, st); break;
        default: set_error("add: unsupported dtype %d", dtype); return PRS_EINVAL;
    }
    if (rc) return rc;
    idx->n += n;
    return 0;
}

// ------------------------------------------------------------------------------------------
// SIMT path dispatch
// ------------------------------------------------------------------------------------------
static int launch_simt(int storage, bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (storage) {
        case PRS_F32: return launch_simt_f32(l2, QB, R, p, grid, smem, st);
        case PRS_F16: return launch_simt_f16(l2, QB, R, p, grid, smem, st);
        default: return launch_simt_bf16(l2, QB, R, p, grid, smem, st);
    }
}

static inline int merge_sortn(long long total, int k) {
    const long long one_shot = total <= MERGE_ONESHOT ? total : 0;
    return next_pow2((int)std::max<long long>(k + MERGE_THREADS, one_shot));
}

static int launch_merge(prs_index* idx, int parts, long long nq, int k, int out_mode, const float* qnorm,
                        float* D, int64_t* I, cudaStream_t st) {
    const int sortn = merge_sortn((long long)parts * k, k);
    const size_t smem = merge_smem_bytes(sortn, MERGE_THREADS, k);
    struct T { ScanTimer& t; cudaStream_t s; T(ScanTimer& t_, cudaStream_t s_) : t(t_), s(s_) { t.begin(s); } ~T() { t.end(s); } } tm(idx->timer_merge, st);
    // 16-bit storage + L2 (tcgen05 scan, expanded form): the merge recomputes the selected rows' distances in
    // the direct form from the stored rows and the 16-bit queries the prep kernel left behind
    Rerank rr{nullptr, nullptr, nullptr, 0, idx->d, idx->pitch, idx->storage == PRS_BF16 ? 1 : 0};
    UmmaState& us = idx->cur->umma;
    if (out_mode == 2 && idx->l2_rerank) {
        rr.x = (const unsigned char*)idx->x;
        if (us.prep_in_scan) { rr.q = us.q_orig; rr.qdtype = us.q_dtype; }      // no 16-bit image: round the original queries on the fly
        else rr.qlow = (const uint16_t*)us.qlow.p;
    }
    uint32_t* zero = us.prep_in_scan ? (uint32_t*)us.boot.p : nullptr;
    const long long zero_words = us.prep_in_scan ? us.zero_words : 0;
    us.prep_in_scan = false;                                                      // consumed by this merge
    if (t_xchg) {
        prs_xchg* x = t_xchg;
        ++x->gen;
        if (x->used) PRS_CUDA(cudaStreamWaitEvent(st, x->event, 0));
        PRS_CUDA(cudaFuncSetAttribute(merge_xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        merge_xchg_kernel<<<(unsigned)nq, MERGE_THREADS, smem, st>>>((const u64*)idx->cur->cand.p, parts,
                                                                    (int)nq, k, sortn, out_mode, qnorm, idx->id_offset,
                                                                    idx->metric == PRS_METRIC_IP ? 1 : 0, rr, x->view, x->gen,
                                                                    x->timeout_ns, D, (long long*)I, x->d_status, zero, zero_words);
        PRS_LAUNCH_CHECK();
        PRS_CUDA(cudaEventRecord(x->event, st));
        x->used = true;
        return 0;
    }
    PRS_CUDA(cudaFuncSetAttribute(merge_cand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_cand_kernel<<<(unsigned)nq, MERGE_THREADS, smem, st>>>((const u64*)idx->cur->cand.p, parts,
                                                                (int)nq, k, sortn, out_mode, qnorm, idx->id_offset, rr, D,
                                                                (long long*)I, zero, zero_words);
    PRS_LAUNCH_CHECK();
    return 0;
}

static int search_simt(prs_index* idx, const float* qf, int q_stride, long long nq, int k, float* D, int64_t* I, cudaStream_t st) {
    const int es = elem_size(idx->storage);
    const int row_bytes = idx->pitch * es;
    if ((long long)(idx->storage != PRS_F32 ? BLK_ROWS : SIMT_NW) * row_bytes > 100 * 1024) {
        set_error("flat scan: d=%d too large for the shared-memory tile (row of %d bytes)", idx->d, row_bytes);
        return PRS_EUNSUP;
    }
    const int cap = std::max(128, next_pow2(2 * k));
    const int sortn = next_pow2(k + SIMT_THREADS);
    const int qb_max = nq >= 8 ? 8 : (nq >= 4 ? 4 : (nq >= 2 ? 2 : 1));
    int grid = 0;
    int rc;
    long long done = 0;
    // workspace is sized for the widest group; every group uses the same grid
    const bool tiled = idx->storage != PRS_F32;
    int R0 = 4;
    while (R0 > 1 && ((!tiled && SIMT_NW * R0 * row_bytes > 48 * 1024) || R0 * qb_max > 32)) R0 >>= 1;
    // tile geometry: fp32 rows are row-major (any multiple of NW*R rows); 16-bit corpora are in the
    // T64 layout, so a tile is a whole number of 64-row blocks
    auto tile_rows_for = [&](int R) -> int {
        if (tiled) return BLK_ROWS * std::max(1, 32768 / (BLK_ROWS * row_bytes));
        return SIMT_NW * R * std::max(1, 32768 / (SIMT_NW * R * row_bytes));
    };
    {
        const int tile_rows = tile_rows_for(R0);
        const long long n_tiles = (idx->n + tile_rows - 1) / tile_rows;
        grid = (int)std::min<long long>(idx->sm_count, n_tiles);
        // a corpus of a few tiles (<= 256 KB: the reference's own indices hold 125-500 rows) goes through ONE CTA, which then
        // writes D / I itself: one launch instead of scan + merge (latency is all that matters at this size)
        if (!t_xchg && (long long)idx->n * row_bytes <= 256 * 1024) grid = 1;
    }
    if ((rc = idx->cur->lists.ensure((size_t)grid * SIMT_NW * qb_max * cap * 8))) return rc;
    if ((rc = idx->cur->cand.ensure((size_t)grid * nq * k * 8))) return rc;
    if ((rc = idx->cur->cand_cnt.ensure((size_t)grid * nq * 4))) return rc;
    while (done < nq) {
        const long long left = nq - done;
        const int QB = left >= 8 ? 8 : (left >= 4 ? 4 : (left >= 2 ? 2 : 1));
        SimtParams p;
        p.x = idx->x; p.q = qf + (size_t)done * q_stride; p.n_rows = idx->n; p.d = idx->d; p.pitch = idx->pitch;
        p.q_stride = q_stride; p.nq = QB; p.k = k; p.cap = cap;
        // R and tile geometry are fixed per search (R0) so that `grid` and the workspace agree
        const int R = R0;
        p.tile_rows = tile_rows_for(R);
        const size_t qbytes = (((size_t)QB * idx->pitch * 4) + 127) & ~(size_t)127;
        const size_t tile_bytes = (size_t)p.tile_rows * row_bytes;
        int stages = (int)((226 * 1024 - 512 - qbytes) / tile_bytes);
        if (stages > SIMT_MAX_STAGES) stages = SIMT_MAX_STAGES;
        if (stages < 2 || (size_t)stages * tile_bytes < (size_t)sortn * 8) {
            set_error("flat scan: cannot fit the pipeline in shared memory (d=%d, k=%d)", idx->d, k);
            return PRS_EUNSUP;
        }
        p.stages = stages;
        p.lists = (u64*)idx->cur->lists.p; p.cand = (u64*)idx->cur->cand.p; p.cand_cnt = (int*)idx->cur->cand_cnt.p;
        p.nq_total = (int)nq; p.q0 = (int)done; p.sortn = sortn;
        p.D = D; p.I = (long long*)I; p.id_offset = idx->id_offset; p.out_mode = idx->metric == PRS_METRIC_L2 ? 1 : 0;
        p.direct = (grid == 1 && !t_xchg) ? 1 : 0;     // one CTA holds the whole (tiny) index: it writes D / I itself
        const size_t smem = 512 + qbytes + (size_t)stages * tile_bytes;
        idx->timer.begin(st);
        if ((rc = launch_simt(idx->storage, idx->metric == PRS_METRIC_L2, QB, R, p, grid, smem, st))) return rc;
        idx->timer.end(st);
        done += QB;
    }
    if (grid == 1 && !t_xchg) return 0;                // the scan kernel wrote D / I (SimtParams::direct)
    return launch_merge(idx, grid, nq, k, idx->metric == PRS_METRIC_L2 ? 1 : 0, nullptr, D, I, st);
}

static int search_device_impl(prs_index* idx, const void* q, int qdtype, long long nq, int k, float* D, int64_t* I, cudaStream_t st) {
    if (k < 1 || k > PRS_MAX_K) { set_error("search: k=%d out of range [1, %d]", k, PRS_MAX_K); return PRS_EINVAL; }
    if (nq < 0) { set_error("search: nq=%lld < 0", nq); return PRS_EINVAL; }
    if (nq == 0) return 0;
    if (!q || !D || !I) { set_error("search: null pointer"); return PRS_EINVAL; }
    if (nq > (1 << 22)) { set_error("search: nq=%lld too large for one call", nq); return PRS_EINVAL; }
    std::lock_guard<std::mutex> lock(idx->mu);
    // the scan workspace is shared by all searches on this index: order a search issued on a new
    // stream after the previous one (same-stream searches are ordered already)
    prs_index::Workspace* ws = &idx->slot[idx->next_slot++ % prs_index::NSLOT];
    idx->cur = ws;
    ws->umma.prep_in_scan = false;
    if (!ws->event) PRS_CUDA(cudaEventCreateWithFlags(&ws->event, cudaEventDisableTiming));
    if (ws->used && ws->stream != st) PRS_CUDA(cudaStreamWaitEvent(st, ws->event, 0));
    struct Rec { prs_index::Workspace* w; cudaStream_t s; ~Rec() { cudaEventRecord(w->event, s); w->stream = s; w->used = true; } } rec{ws, st};
    if (idx->n == 0 && t_xchg) {
        // an empty shard of a row-sharded corpus still takes part in the exchange: it contributes empty lists
        return launch_merge(idx, 0, nq, k, idx->metric == PRS_METRIC_L2 ? 1 : 0, nullptr, D, I, st);
    }
    if (idx->n == 0) {
        const long long tot = nq * k;
        fill_empty_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(D, (long long*)I, tot,
                                                                        idx->metric == PRS_METRIC_L2 ? 3.402823466e+38f : -3.402823466e+38f);
        PRS_LAUNCH_CHECK();
        return 0;
    }
    int rc;
    int path = idx->path_force;
    const bool wide = umma_wide_eligible(idx->storage, idx->pitch, nq, k, idx->n);
    if (path == 0) path = (umma_eligible(idx->storage, idx->d, idx->pitch, nq, k) || wide) ? 2 : 1;
    if (path == 2 && wide) {
        // 16 < k <= 1024: sample -> threshold -> collect -> select (flat_umma.cuh), then the usual merge of ONE
        // part; large batches go through in chunks so that the collection buffer stays bounded
        idx->last_path = 2;
        const long long chunk = umma_wide_chunk(k);
        const size_t qes = elem_size(qdtype);
        for (long long q0 = 0; q0 < nq; q0 += chunk) {
            const long long nqc = std::min<long long>(chunk, nq - q0);
            if ((rc = idx->cur->qnorm.ensure((size_t)nqc * 4))) return rc;
            if ((rc = search_umma_wide(idx->cur->umma, idx->x, idx->xnorm, idx->n, idx->d, idx->pitch, idx->storage, idx->metric, idx->sm_count,
                                       (const unsigned char*)q + (size_t)q0 * idx->d * qes, qdtype, nqc, k, (float*)idx->cur->qnorm.p,
                                       idx->cur->cand, idx->cur->cand_cnt, st, &idx->timer, &idx->timer_prep))) return rc;
            if ((rc = launch_merge(idx, 1, nqc, k, idx->metric == PRS_METRIC_L2 ? 2 : 0, (const float*)idx->cur->qnorm.p,
                                   D + (size_t)q0 * k, I + (size_t)q0 * k, st))) return rc;
        }
        return 0;
    }
    if (path == 2) {
        if (!umma_eligible(idx->storage, idx->d, idx->pitch, nq, k)) {
            set_error("tcgen05 path needs fp16/bf16 storage, d <= 768 and k <= %d, or k <= %d on >= 32k rows (storage=%d, d=%d, k=%d)", UMMA_MAX_K, UMMA_WIDE_MAX_K, idx->storage, idx->d, k);
            return PRS_EUNSUP;
        }
        idx->last_path = 2;
        if ((rc = idx->cur->qnorm.ensure((size_t)nq * 4))) return rc;
        int parts = 0;
        const int out_mode = idx->metric == PRS_METRIC_L2 ? 2 : 0;
        UmmaTail tail;
        // row-sharded searches are TWO launches -- the scan (which prepares its own queries) and the merge + exchange kernel --
        // unless fuse == 2: measured at N = 2 (1M x 768, B = 64) the push-in-the-tail variant is 0.1855 ms per step against
        // 0.1810 ms (the scan kernel's tail then pays the NVLink store acknowledgements before it can release its flags).
        // fuse == 3 asks for the same two launches on an unsharded index (tests)
        tail.enable = t_xchg ? idx->fuse == 2 : (idx->fuse == 1 || idx->fuse == 2);
        tail.prep_in_scan = idx->fuse != 0;             // row-sharded default: scan (with its own query preparation) + merge/exchange
        tail.out_mode = out_mode; tail.largest = idx->metric == PRS_METRIC_IP ? 1 : 0; tail.id_offset = idx->id_offset;
        tail.D = D; tail.I = (long long*)I; tail.device = idx->device;
        tail.rerank_x = (out_mode == 2 && idx->l2_rerank) ? (const unsigned char*)idx->x : nullptr;
        tail.xchg = t_xchg;
        tail.timer_merge = &idx->timer_merge;
        bool fused = false;
        if ((rc = search_umma(idx->cur->umma, idx->x, idx->xnorm, idx->n, idx->d, idx->pitch, idx->storage, idx->metric, idx->sm_count,
                              q, qdtype, nq, k, (float*)idx->cur->qnorm.p, idx->cur->cand, idx->cur->cand_cnt, &parts, st, &idx->timer, &idx->timer_prep,
                              &tail, &fused))) return rc;
        idx->last_fused = fused ? 1 : 0;
        if (fused) return 0;                       // the scan kernel merged (and exchanged) the lists itself
        return launch_merge(idx, parts, nq, k, out_mode, (const float*)idx->cur->qnorm.p, D, I, st);
    }
    const float* qf = (const float*)q;
    if (qdtype != PRS_F32) {
        const long long tot = nq * idx->d;
        if ((rc = idx->cur->qf32.ensure((size_t)tot * 4))) return rc;
        if (qdtype == PRS_F16) to_f32_kernel<__half><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const __half*)q, tot, (float*)idx->cur->qf32.p);
        else if (qdtype == PRS_BF16) to_f32_kernel<__nv_bfloat16><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)q, tot, (float*)idx->cur->qf32.p);
        else { set_error("search: unsupported query dtype %d", qdtype); return PRS_EINVAL; }
        PRS_LAUNCH_CHECK();
        qf = (const float*)idx->cur->qf32.p;
    }
    idx->last_path = 1;
    return search_simt(idx, qf, idx->d, nq, k, D, I, st);
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char* prs_last_error(void) { return t_error.c_str(); }
int64_t prs_launch_count(void) { return (int64_t)g_launches.load(); }

int prs_device_arch(int device) {
    static std::atomic<int> cache[64];                       // cudaGetDeviceProperties costs milliseconds: ask once per device
    if (device >= 0 && device < 64 && cache[device].load() > 0) return cache[device].load();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device available"); return PRS_ECUDA; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (have %d)", device, ndev); return PRS_EINVAL; }
    cudaDeviceProp prop;
    PRS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (device < 64) cache[device].store(prop.major * 10 + prop.minor);
    return prop.major * 10 + prop.minor;
}

int prs_index_create(int d, int metric, int storage, int device, prs_index** out) {
    if (!out) { set_error("create: out is null"); return PRS_EINVAL; }
    *out = nullptr;
    if (d < 1 || d > 65536) { set_error("create: d=%d out of range", d); return PRS_EINVAL; }
    if (metric != PRS_METRIC_IP && metric != PRS_METRIC_L2) { set_error("create: unknown metric %d", metric); return PRS_EINVAL; }
    if (storage != PRS_F32 && storage != PRS_F16 && storage != PRS_BF16) { set_error("create: unknown storage %d", storage); return PRS_EINVAL; }
    int arch = prs_device_arch(device);
    if (arch < 0) return arch;
    if (arch != 100) { set_error("libprs is built for sm_100a (B200); device %d is sm_%d", device, arch); return PRS_ECUDA; }
    prs_index* idx = new (std::nothrow) prs_index();
    if (!idx) { set_error("out of host memory"); return PRS_ENOMEM; }
    idx->d = d; idx->pitch = (d + 63) / 64 * 64; idx->metric = metric; idx->storage = storage; idx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete idx; set_error("cudaGetDeviceProperties failed"); return PRS_ECUDA; }
    idx->sm_count = prop.multiProcessorCount;
    *out = idx;
    return 0;
}

void prs_index_free(prs_index* idx) {
    if (!idx) return;
    DeviceGuard g(idx->device);
    if (idx->x) cudaFree(idx->x);
    if (idx->xnorm) cudaFree(idx->xnorm);
    for (auto& w : idx->slot) {
        w.lists.release(); w.cand.release(); w.cand_cnt.release(); w.qf32.release(); w.qnorm.release();
        w.umma.release();
        if (w.event) cudaEventDestroy(w.event);
    }
    idx->hD.release(); idx->hI.release(); idx->hQ.release(); idx->stage.release();
    idx->pQ.release(); idx->pD.release(); idx->pI.release();
    delete idx;
}

int prs_index_reserve(prs_index* idx, int64_t n_total) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    std::lock_guard<std::mutex> lock(idx->mu);
    PRS_CUDA(cudaDeviceSynchronize());
    return index_grow(idx, n_total);
}

int prs_index_add_device(prs_index* idx, const void* x, int dtype, int64_t n, void* stream) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    if (n < 0 || (n > 0 && !x)) { set_error("add: bad arguments"); return PRS_EINVAL; }
    if (idx->n + n > 0xFFFFFFFFll) { set_error("add: more than 2^32-1 rows per index (shard it)"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    std::lock_guard<std::mutex> lock(idx->mu);
    return add_device_impl(idx, x, dtype, n, (cudaStream_t)stream);
}

int prs_index_add_host(prs_index* idx, const float* x, int64_t n) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    if (n < 0 || (n > 0 && !x)) { set_error("add: bad arguments"); return PRS_EINVAL; }
    if (idx->n + n > 0xFFFFFFFFll) { set_error("add: more than 2^32-1 rows per index (shard it)"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    std::lock_guard<std::mutex> lock(idx->mu);
    const long long chunk = std::max<long long>(1, (64ll << 20) / ((long long)idx->d * 4));
    int rc;
    if (idx->n + n > idx->cap_rows) {
        PRS_CUDA(cudaDeviceSynchronize());
        if ((rc = index_grow(idx, idx->n + n))) return rc;
    }
    for (long long o = 0; o < n; o += chunk) {
        const long long c = std::min<long long>(chunk, n - o);
        if ((rc = idx->stage.ensure((size_t)c * idx->d * 4))) return rc;
        PRS_CUDA(cudaMemcpy(idx->stage.p, x + (size_t)o * idx->d, (size_t)c * idx->d * 4, cudaMemcpyHostToDevice));
        if ((rc = add_device_impl(idx, idx->stage.p, PRS_F32, c, 0))) return rc;
        PRS_CUDA(cudaStreamSynchronize(0));
    }
    return 0;
}

int64_t prs_index_ntotal(const prs_index* idx) { return idx ? idx->n : -1; }
int prs_index_d(const prs_index* idx) { return idx ? idx->d : -1; }
int prs_index_metric(const prs_index* idx) { return idx ? idx->metric : -1; }
int prs_index_storage(const prs_index* idx) { return idx ? idx->storage : -1; }
int prs_index_last_path(const prs_index* idx) { return idx ? idx->last_path : -1; }

int prs_index_set_id_offset(prs_index* idx, int64_t off) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    idx->id_offset = off;
    return 0;
}
int prs_index_set_path(prs_index* idx, int path) {
    if (!idx || path < 0 || path > 2) { set_error("set_path: bad arguments"); return PRS_EINVAL; }
    idx->path_force = path;
    return 0;
}

int prs_index_set_fused(prs_index* idx, int enable) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    idx->fuse = enable < 0 ? 0 : (enable > 3 ? 3 : enable);
    return 0;
}
int prs_index_last_fused(const prs_index* idx) { return idx ? idx->last_fused : -1; }

int prs_index_set_timing(prs_index* idx, int enable) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    std::lock_guard<std::mutex> lock(idx->mu);
    idx->timer.enabled = idx->timer_prep.enabled = idx->timer_merge.enabled = enable != 0;
    return 0;
}
int prs_index_phase_times(prs_index* idx, double* prep_ms, double* merge_ms) {
    if (!idx || !prep_ms || !merge_ms) { set_error("phase_times: bad arguments"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    std::lock_guard<std::mutex> lock(idx->mu);
    long long n = 0;
    int rc = idx->timer_prep.collect(prep_ms, &n);
    if (rc) return rc;
    return idx->timer_merge.collect(merge_ms, &n);
}
int prs_index_scan_time(prs_index* idx, double* total_ms, int64_t* launches) {
    if (!idx || !total_ms || !launches) { set_error("scan_time: bad arguments"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    std::lock_guard<std::mutex> lock(idx->mu);
    long long n = 0;
    int rc = idx->timer.collect(total_ms, &n);
    *launches = n;
    return rc;
}

int prs_index_search_device(prs_index* idx, const void* q, int qdtype, int64_t nq, int k, float* D, int64_t* I, void* stream) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    return search_device_impl(idx, q, qdtype, nq, k, D, I, (cudaStream_t)stream);
}

// page-locked host memory that the device can address (cudaHostAlloc / cudaHostRegister / torch pin_memory)
static bool mapped_host_pointer(const void* p, void** dptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dptr = a.devicePointer;
    return true;
}

int prs_index_search_host(prs_index* idx, const float* q, int64_t nq, int k, float* D, int64_t* I) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    if (k < 1 || k > PRS_MAX_K) { set_error("search: k=%d out of range [1, %d]", k, PRS_MAX_K); return PRS_EINVAL; }
    if (nq < 0) { set_error("search: nq < 0"); return PRS_EINVAL; }
    if (nq == 0) return 0;
    if (!q || !D || !I) { set_error("search: null pointer"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    int rc;
    // concurrent host searches on one index serialise here (re-entrant, not parallel): they share
    // the index's staging buffers and the legacy stream
    std::lock_guard<std::mutex> hl(idx->host_mu);
    void *dq, *dD, *dI;
    // Pinned caller buffers + the tcgen05 path: no staging copies.  The query-preparation kernel reads
    // every query element exactly once straight from the pinned host buffer (the host->device
    // transfer of the step) and the merge kernel stores the k results per query straight into the
    // caller's pinned D / I (the device->host transfer); one stream synchronisation ends the call.
    // (The CUDA-core scan re-reads the queries in every CTA, so it keeps the staged copy.)
    const bool zero_copy = idx->path_force != 1 && idx->n > 0 &&
        (umma_eligible(idx->storage, idx->d, idx->pitch, nq, k) || umma_wide_eligible(idx->storage, idx->pitch, nq, k, idx->n));
    if (zero_copy && mapped_host_pointer(q, &dq) && mapped_host_pointer(D, &dD) && mapped_host_pointer(I, &dI)) {
        if ((rc = search_device_impl(idx, dq, PRS_F32, nq, k, (float*)dD, (int64_t*)dI, 0))) return rc;
        PRS_CUDA(cudaStreamSynchronize(0));
        return 0;
    }
    // Pageable caller buffers (what `index.search(numpy)` passes), small batches: the same zero-copy search through the
    // index's own page-locked buffers -- one host memcpy in, one out -- instead of three staged driver copies.
    if (zero_copy && (size_t)nq * idx->d * 4 <= (size_t)(4u << 20) && (size_t)nq * k * 8 <= (size_t)(4u << 20)) {
        const size_t qb = (size_t)nq * idx->d * 4, db = (size_t)nq * k * 4, ib = (size_t)nq * k * 8;
        if ((rc = idx->pQ.ensure(qb)) || (rc = idx->pD.ensure(db)) || (rc = idx->pI.ensure(ib))) return rc;
        memcpy(idx->pQ.p, q, qb);
        if ((rc = search_device_impl(idx, idx->pQ.dp, PRS_F32, nq, k, (float*)idx->pD.dp, (int64_t*)idx->pI.dp, 0))) return rc;
        PRS_CUDA(cudaStreamSynchronize(0));
        memcpy(D, idx->pD.p, db);
        memcpy(I, idx->pI.p, ib);
        return 0;
    }
    // Other small pageable calls (the CUDA-core scan re-reads its queries in every CTA, so they are copied to the device):
    // the copy comes out of the index's page-locked buffer (a true asynchronous DMA, no driver staging) and the merge
    // kernel stores the results straight into page-locked memory -- no device->host copies at all.  This is the
    // reference's own call shape (nq = 1, k = 5 on a 125-row fp32 index, src/retrieval.py:102).
    if ((size_t)nq * idx->d * 4 <= (size_t)(4u << 20) && (size_t)nq * k * 8 <= (size_t)(4u << 20)) {
        const size_t qb = (size_t)nq * idx->d * 4, db = (size_t)nq * k * 4, ib = (size_t)nq * k * 8;
        if ((rc = idx->pQ.ensure(qb)) || (rc = idx->pD.ensure(db)) || (rc = idx->pI.ensure(ib))) return rc;
        {
            std::lock_guard<std::mutex> lock(idx->mu);
            if ((rc = idx->hQ.ensure(qb))) return rc;
        }
        memcpy(idx->pQ.p, q, qb);
        // a tiny index is scanned by ONE CTA (search_simt): it reads its few queries straight from the page-locked buffer
        const bool one_cta = !zero_copy && idx->path_force != 2 && idx->n > 0 && nq <= 8 &&
                             (long long)idx->n * idx->pitch * elem_size(idx->storage) <= 256 * 1024;
        const void* qsrc = idx->pQ.dp;
        if (!one_cta) {
            PRS_CUDA(cudaMemcpyAsync(idx->hQ.p, idx->pQ.p, qb, cudaMemcpyHostToDevice, 0));
            qsrc = idx->hQ.p;
        }
        if ((rc = search_device_impl(idx, qsrc, PRS_F32, nq, k, (float*)idx->pD.dp, (int64_t*)idx->pI.dp, 0))) return rc;
        PRS_CUDA(cudaStreamSynchronize(0));
        memcpy(D, idx->pD.p, db);
        memcpy(I, idx->pI.p, ib);
        return 0;
    }
    {
        std::lock_guard<std::mutex> lock(idx->mu);
        if ((rc = idx->hQ.ensure((size_t)nq * idx->d * 4))) return rc;
        if ((rc = idx->hD.ensure((size_t)nq * k * 4))) return rc;
        if ((rc = idx->hI.ensure((size_t)nq * k * 8))) return rc;
        dq = idx->hQ.p; dD = idx->hD.p; dI = idx->hI.p;
    }
    PRS_CUDA(cudaMemcpyAsync(dq, q, (size_t)nq * idx->d * 4, cudaMemcpyHostToDevice, 0));
    if ((rc = search_device_impl(idx, dq, PRS_F32, nq, k, (float*)dD, (int64_t*)dI, 0))) return rc;
    PRS_CUDA(cudaMemcpyAsync(D, dD, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, 0));
    PRS_CUDA(cudaMemcpyAsync(I, dI, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, 0));
    PRS_CUDA(cudaStreamSynchronize(0));
    return 0;
}

int prs_index_reconstruct_host(prs_index* idx, int64_t i0, int64_t n, float* out) {
    if (!idx) { set_error("null index"); return PRS_EINVAL; }
    if (i0 < 0 || n < 0 || i0 + n > idx->n || (n > 0 && !out)) { set_error("reconstruct: range out of bounds"); return PRS_EINVAL; }
    if (n == 0) return 0;
    DeviceGuard g(idx->device);
    std::lock_guard<std::mutex> lock(idx->mu);
    const long long chunk = std::max<long long>(1, (64ll << 20) / ((long long)idx->d * 4));
    int rc;
    for (long long o = 0; o < n; o += chunk) {
        const long long c = std::min<long long>(chunk, n - o);
        if ((rc = idx->stage.ensure((size_t)c * idx->d * 4))) return rc;
        const long long tot = c * idx->d;
        const unsigned grid = (unsigned)((tot + 255) / 256);
        switch (idx->storage) {
            case PRS_F32: reconstruct_kernel<float><<<grid, 256>>>((const float*)idx->x, i0 + o, c, idx->d, idx->pitch, (float*)idx->stage.p); break;
            case PRS_F16: reconstruct_kernel<__half><<<grid, 256>>>((const __half*)idx->x, i0 + o, c, idx->d, idx->pitch, (float*)idx->stage.p); break;
            default: reconstruct_kernel<__nv_bfloat16><<<grid, 256>>>((const __nv_bfloat16*)idx->x, i0 + o, c, idx->d, idx->pitch, (float*)idx->stage.p); break;
        }
        PRS_LAUNCH_CHECK();
        PRS_CUDA(cudaMemcpy(out + (size_t)o * idx->d, idx->stage.p, (size_t)tot * 4, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int prs_merge_topk_device(const float* Dp, const int64_t* Ip, int nparts, int64_t nq, int k, int largest,
                          int tie_high_id, float* D, int64_t* I, int device, void* stream) {
    if (nparts < 1 || nq < 0 || k < 1 || k > PRS_MAX_K) { set_error("merge: bad arguments"); return PRS_EINVAL; }
    if ((long long)nparts * k > (1ll << 30)) { set_error("merge: nparts*k too large"); return PRS_EINVAL; }
    if (nq == 0) return 0;
    if (!Dp || !Ip || !D || !I) { set_error("merge: null pointer"); return PRS_EINVAL; }
    DeviceGuard g(device);
    const int sortn = merge_sortn((long long)nparts * k, k);
    const size_t smem = (size_t)sortn * 8 + MERGE_THREADS * 8 + 16;
    PRS_CUDA(cudaFuncSetAttribute(merge_parts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_parts_kernel<<<(unsigned)nq, MERGE_THREADS, smem, (cudaStream_t)stream>>>(Dp, (const long long*)Ip, nparts, nq, k, sortn,
                                                                                   largest, tie_high_id, D, (long long*)I);
    PRS_LAUNCH_CHECK();
    return 0;
}

// ---- row-sharded search over peer memory (xchg.cuh) ----
int prs_xchg_create(int device, int n_ranks, int rank, int64_t nq_cap, int k_cap, prs_xchg** out) {
    if (!out) { set_error("xchg_create: out is null"); return PRS_EINVAL; }
    *out = nullptr;
    if (n_ranks < 1 || n_ranks > XCHG_MAX_RANKS || rank < 0 || rank >= n_ranks || nq_cap < 1 || nq_cap > (1 << 22) ||
        k_cap < 1 || k_cap > PRS_MAX_K) { set_error("xchg_create: bad arguments"); return PRS_EINVAL; }
    int arch = prs_device_arch(device);
    if (arch < 0) return arch;
    DeviceGuard g(device);
    prs_xchg* x = new (std::nothrow) prs_xchg();
    if (!x) { set_error("out of host memory"); return PRS_ENOMEM; }
    x->device = device; x->G = n_ranks; x->rank = rank; x->nq_cap = (int)nq_cap; x->cap = nq_cap * k_cap;
    x->bytes = xchg_ids_bytes(x) + 2 * xchg_vals_bytes(x) + xchg_flags_bytes(x) + 256;
    cudaError_t e = cudaMalloc(&x->base, x->bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("xchg_create: cudaMalloc(%zu) failed: %s", x->bytes, cudaGetErrorString(e));
        delete x;
        return PRS_ENOMEM;
    }
    PRS_CUDA(cudaMemset(x->base, 0, x->bytes));
    PRS_CUDA(cudaDeviceSynchronize());
    if (cudaHostAlloc((void**)&x->h_status, 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&x->d_status, x->h_status, 0) != cudaSuccess ||
        cudaEventCreateWithFlags(&x->event, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        set_error("xchg_create: could not allocate the status word / event");
        if (x->h_status) cudaFreeHost(x->h_status);
        cudaFree(x->base);
        delete x;
        return PRS_ENOMEM;
    }
    *x->h_status = 0;
    x->peer_base[rank] = x->base;
    xchg_fill_view(x);
    *out = x;
    return 0;
}

int prs_xchg_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int prs_xchg_get_handle(prs_xchg* x, void* handle_out) {
    if (!x || !handle_out) { set_error("xchg_get_handle: bad arguments"); return PRS_EINVAL; }
    DeviceGuard g(x->device);
    cudaIpcMemHandle_t h;
    PRS_CUDA(cudaIpcGetMemHandle(&h, x->base));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

int prs_xchg_open_peers(prs_xchg* x, const void* handles) {
    if (!x || !handles) { set_error("xchg_open_peers: bad arguments"); return PRS_EINVAL; }
    DeviceGuard g(x->device);
    for (int p = 0; p < x->G; ++p) {
        if (p == x->rank || x->opened[p]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)handles + (size_t)p * sizeof(h), sizeof(h));
        void* ptr = nullptr;
        PRS_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer_base[p] = ptr; x->opened[p] = true;
    }
    xchg_fill_view(x);
    return 0;
}

// the kernel's report lives in mapped host memory: reading it costs nothing; it is reset once reported
static int xchg_take_status(prs_xchg* x) {
    if (*(volatile int*)x->h_status == 0) return 0;
    *(volatile int*)x->h_status = 0;
    set_error("sharded search: timed out (%.1f s) waiting for a peer rank's candidate lists; the affected queries returned ids -1",
              (double)x->timeout_ns * 1e-9);
    return PRS_ECUDA;
}

// exchange buffers of ONE process on different devices (group.cu): peer access is enabled, so the peers' buffers
// are addressed directly instead of through CUDA IPC handles
int prs_xchg_link_local(prs_xchg** xs, int n) {
    if (!xs || n < 1 || n > XCHG_MAX_RANKS) { set_error("xchg_link_local: bad arguments"); return PRS_EINVAL; }
    for (int i = 0; i < n; ++i) if (!xs[i] || xs[i]->G != n || xs[i]->rank != i) { set_error("xchg_link_local: buffer %d does not belong to a group of %d", i, n); return PRS_EINVAL; }
    for (int i = 0; i < n; ++i) {
        for (int p = 0; p < n; ++p) xs[i]->peer_base[p] = xs[p]->base;
        xchg_fill_view(xs[i]);
    }
    return 0;
}

int prs_xchg_status(prs_xchg* x) {
    if (!x) { set_error("null exchange"); return PRS_EINVAL; }
    DeviceGuard g(x->device);
    PRS_CUDA(cudaDeviceSynchronize());
    return xchg_take_status(x);
}

int prs_xchg_set_timeout_ms(prs_xchg* x, int64_t ms) {
    if (!x || ms < 1) { set_error("xchg_set_timeout_ms: bad arguments"); return PRS_EINVAL; }
    x->timeout_ns = (unsigned long long)ms * 1000000ull;
    return 0;
}

void prs_xchg_free(prs_xchg* x) {
    if (!x) return;
    DeviceGuard g(x->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < x->G; ++p) if (x->opened[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
    if (x->base) cudaFree(x->base);
    if (x->h_status) cudaFreeHost(x->h_status);
    if (x->event) cudaEventDestroy(x->event);
    delete x;
}

int prs_index_search_sharded_device(prs_index* idx, prs_xchg* x, const void* q, int qdtype, int64_t nq, int k,
                                    float* D, int64_t* I, void* stream) {
    if (!idx || !x) { set_error("sharded search: null handle"); return PRS_EINVAL; }
    if (x->device != idx->device) { set_error("sharded search: index and exchange buffer live on different devices"); return PRS_EINVAL; }
    if (nq < 1 || nq > x->nq_cap || k < 1 || (long long)nq * k > x->cap) {
        set_error("sharded search: nq=%lld k=%d exceed the exchange buffer (nq_cap=%d, entries=%lld)", (long long)nq, k, x->nq_cap, x->cap);
        return PRS_EINVAL;
    }
    for (int p = 0; p < x->G; ++p) if (!x->peer_base[p]) { set_error("sharded search: peer %d not opened", p); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    // a timeout reported by an EARLIER search on this context surfaces here (cheap: mapped host word)
    if (int rc = xchg_take_status(x)) return rc;
    struct Scope { ~Scope() { t_xchg = nullptr; } } scope;
    t_xchg = x;               // read by launch_merge on this thread
    return search_device_impl(idx, q, qdtype, nq, k, D, I, (cudaStream_t)stream);
}

// ---- on-disk format: faiss IndexFlat (IxF2 / IxFI), SURVEY.md 8f-2 ----
#pragma pack(push, 1)
struct FlatHeader {
    char fourcc[4];
    int32_t d;
    int64_t ntotal;
    int64_t dummy1, dummy2;
    uint8_t is_trained;
    int32_t metric_type;
    uint64_t n_values;
};
#pragma pack(pop)
static_assert(sizeof(FlatHeader) == 45, "faiss IndexFlat header is 45 bytes");

int prs_index_write(prs_index* idx, const char* path) {
    if (!idx || !path) { set_error("write: bad arguments"); return PRS_EINVAL; }
    FlatHeader h;
    const char* cc = idx->storage == PRS_F32 ? (idx->metric == PRS_METRIC_L2 ? "IxF2" : "IxFI")
                                             : (idx->storage == PRS_F16 ? "PRSh" : "PRSb");
    memcpy(h.fourcc, cc, 4);
    h.d = idx->d; h.ntotal = idx->n; h.dummy1 = h.dummy2 = 1 << 20; h.is_trained = 1; h.metric_type = idx->metric;
    h.n_values = (uint64_t)idx->n * idx->d;
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("write: cannot open %s: %s", path, strerror(errno)); return PRS_EIO; }
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    const long long chunk = std::max<long long>(1, (64ll << 20) / ((long long)idx->d * 4));
    std::vector<float> host((size_t)std::min<long long>(chunk, std::max<long long>(idx->n, 1)) * idx->d);
    std::vector<uint16_t> half;
    for (long long o = 0; ok && o < idx->n; o += chunk) {
        const long long c = std::min<long long>(chunk, idx->n - o);
        int rc = prs_index_reconstruct_host(idx, o, c, host.data());
        if (rc) { fclose(f); return rc; }
        if (idx->storage == PRS_F32) {
            ok = fwrite(host.data(), 4, (size_t)c * idx->d, f) == (size_t)c * idx->d;
        } else {
            // values are exactly representable in the 16-bit type (they came from it)
            half.resize((size_t)c * idx->d);
            for (size_t i = 0; i < half.size(); ++i) {
                uint32_t u; memcpy(&u, &host[i], 4);
                if (idx->storage == PRS_BF16) half[i] = (uint16_t)(u >> 16);
                else { __half hv = __float2half_rn(host[i]); memcpy(&half[i], &hv, 2); }
            }
            ok = fwrite(half.data(), 2, half.size(), f) == half.size();
        }
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) { set_error("write: short write to %s", path); return PRS_EIO; }
    return 0;
}

int prs_index_read(const char* path, int storage, int device, prs_index** out) {
    if (!path || !out) { set_error("read: bad arguments"); return PRS_EINVAL; }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("read: cannot open %s: %s", path, strerror(errno)); return PRS_EIO; }
    FlatHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1) { fclose(f); set_error("read: %s is shorter than an IndexFlat header", path); return PRS_EIO; }
    int file_dt;
    if (!memcmp(h.fourcc, "IxF2", 4) || !memcmp(h.fourcc, "IxFI", 4)) file_dt = PRS_F32;
    else if (!memcmp(h.fourcc, "PRSh", 4)) file_dt = PRS_F16;
    else if (!memcmp(h.fourcc, "PRSb", 4)) file_dt = PRS_BF16;
    else {
        fclose(f);
        set_error("read: %s has fourcc '%.4s'; only flat indices (IxF2/IxFI/PRSh/PRSb) are supported", path, h.fourcc);
        return PRS_EUNSUP;
    }
    if (h.d < 1 || h.ntotal < 0 || h.n_values != (uint64_t)h.ntotal * (uint64_t)h.d ||
        (h.metric_type != PRS_METRIC_IP && h.metric_type != PRS_METRIC_L2)) {
        fclose(f); set_error("read: %s has an inconsistent header", path); return PRS_EIO;
    }
    prs_index* idx = nullptr;
    int rc = prs_index_create(h.d, h.metric_type, storage, device, &idx);
    if (rc) { fclose(f); return rc; }
    DeviceGuard g(device);
    if ((rc = prs_index_reserve(idx, h.ntotal))) { fclose(f); prs_index_free(idx); return rc; }
    const int fes = elem_size(file_dt);
    const long long chunk = std::max<long long>(1, (64ll << 20) / ((long long)h.d * fes));
    std::vector<unsigned char> host((size_t)std::min<long long>(chunk, std::max<long long>(h.ntotal, 1)) * h.d * fes);
    for (long long o = 0; o < h.ntotal; o += chunk) {
        const long long c = std::min<long long>(chunk, h.ntotal - o);
        const size_t bytes = (size_t)c * h.d * fes;
        if (fread(host.data(), 1, bytes, f) != bytes) { fclose(f); prs_index_free(idx); set_error("read: %s is truncated", path); return PRS_EIO; }
        if (file_dt == PRS_F32) rc = prs_index_add_host(idx, (const float*)host.data(), c);
        else {
            std::lock_guard<std::mutex> lock(idx->mu);
            rc = idx->stage.ensure(bytes);
            if (!rc) {
                if (cudaMemcpy(idx->stage.p, host.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("read: H2D copy failed"); rc = PRS_ECUDA; }
                else { rc = add_device_impl(idx, idx->stage.p, file_dt, c, 0); cudaStreamSynchronize(0); }
            }
        }
        if (rc) { fclose(f); prs_index_free(idx); return rc; }
    }
    fclose(f);
    *out = idx;
    return 0;
}

}  // extern "C"

// ---- sharded container (second half): one file per shard holding the HBM image itself ----
// A shard file is [128-byte header][pad to 4096][payload][pad to 4096][norms]:
//   payload = the corpus exactly as it sits in HBM (16-bit storage: T64 blocks, rows padded to a multiple of 64;
//             fp32: row-major [rows, pitch]), norms = float32 ||x||^2 of the stored rows.
// Loading is therefore a straight copy -- no conversion kernel, no re-computation of the norms -- streamed through
// two page-locked staging buffers with one cudaMemcpyAsync per 64 MB chunk, the file read of chunk i+1 overlapping
// the copy of chunk i.  Offsets are page aligned, so the payload can also be mmap'ed as is.  Which shard belongs to
// which rank is the manifest's business (container.py); every rank loads its own file(s) in parallel.
#pragma pack(push, 1)
struct ShardHeader {
    char fourcc[4];                  // "PRST"
    int32_t version, d, pitch, metric, storage;
    int64_t rows, rows_padded, id_offset;
    int64_t payload_offset, payload_bytes, norms_offset, norms_bytes;
    char pad[128 - 4 - 5 * 4 - 7 * 8];
};
#pragma pack(pop)
static_assert(sizeof(ShardHeader) == 128, "shard header is 128 bytes");
constexpr size_t SHARD_CHUNK = (size_t)64 << 20;

struct PinnedPair {
    void* buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    int init() {
        for (int i = 0; i < 2; ++i) {
            if (cudaHostAlloc(&buf[i], SHARD_CHUNK, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); set_error("shard I/O: cannot allocate page-locked staging buffers"); return PRS_ENOMEM; }
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); set_error("shard I/O: cudaEventCreate failed"); return PRS_ECUDA; }
        }
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); set_error("shard I/O: cudaStreamCreate failed"); return PRS_ECUDA; }
        return 0;
    }
    ~PinnedPair() {
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
        for (int i = 0; i < 2; ++i) { if (ev[i]) cudaEventDestroy(ev[i]); if (buf[i]) cudaFreeHost(buf[i]); }
    }
};

// file[off, off+bytes) -> device, double buffered
static int file_to_device(FILE* f, long long off, void* dptr, size_t bytes, PinnedPair& pp, const char* path) {
    if (fseeko(f, (off_t)off, SEEK_SET) != 0) { set_error("read_shard: seek failed in %s", path); return PRS_EIO; }
    int b = 0;
    bool used[2] = {false, false};
    for (size_t o = 0; o < bytes; o += SHARD_CHUNK, b ^= 1) {
        const size_t c = std::min(SHARD_CHUNK, bytes - o);
        if (used[b]) PRS_CUDA(cudaEventSynchronize(pp.ev[b]));             // the copy that last used this buffer is done
        if (fread(pp.buf[b], 1, c, f) != c) { set_error("read_shard: %s is truncated", path); return PRS_EIO; }
        PRS_CUDA(cudaMemcpyAsync((unsigned char*)dptr + o, pp.buf[b], c, cudaMemcpyHostToDevice, pp.st));
        PRS_CUDA(cudaEventRecord(pp.ev[b], pp.st));
        used[b] = true;
    }
    PRS_CUDA(cudaStreamSynchronize(pp.st));
    return 0;
}
// device -> file (appended at the current position), double buffered
static int device_to_file(FILE* f, const void* dptr, size_t bytes, PinnedPair& pp, const char* path) {
    int b = 0;
    size_t pending[2] = {0, 0};
    for (size_t o = 0; o < bytes || pending[0] || pending[1]; o += SHARD_CHUNK, b ^= 1) {
        if (pending[b]) {                                                   // flush what this buffer received two steps ago
            PRS_CUDA(cudaEventSynchronize(pp.ev[b]));
            if (fwrite(pp.buf[b], 1, pending[b], f) != pending[b]) { set_error("write_shard: short write to %s", path); return PRS_EIO; }
            pending[b] = 0;
        }
        if (o < bytes) {
            const size_t c = std::min(SHARD_CHUNK, bytes - o);
            PRS_CUDA(cudaMemcpyAsync(pp.buf[b], (const unsigned char*)dptr + o, c, cudaMemcpyDeviceToHost, pp.st));
            PRS_CUDA(cudaEventRecord(pp.ev[b], pp.st));
            pending[b] = c;
        }
        if (o >= bytes && !pending[b ^ 1]) break;
    }
    return 0;
}
static int pad_file_to(FILE* f, long long target, const char* path) {
    static const char zeros[4096] = {0};
    long long pos = (long long)ftello(f);
    while (pos < target) {
        const size_t c = (size_t)std::min<long long>(4096, target - pos);
        if (fwrite(zeros, 1, c, f) != c) { set_error("write_shard: short write to %s", path); return PRS_EIO; }
        pos += (long long)c;
    }
    return 0;
}
static inline long long align4k(long long v) { return (v + 4095) & ~4095ll; }

extern "C" {

int prs_index_write_shard(prs_index* idx, const char* path) {
    if (!idx || !path) { set_error("write_shard: bad arguments"); return PRS_EINVAL; }
    DeviceGuard g(idx->device);
    std::lock_guard<std::mutex> lock(idx->mu);
    PRS_CUDA(cudaDeviceSynchronize());
    const size_t es = elem_size(idx->storage);
    ShardHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.fourcc, "PRST", 4);
    h.version = 1; h.d = idx->d; h.pitch = idx->pitch; h.metric = idx->metric; h.storage = idx->storage;
    h.rows = idx->n;
    h.rows_padded = idx->storage != PRS_F32 ? (idx->n + BLK_ROWS - 1) / BLK_ROWS * BLK_ROWS : idx->n;
    h.id_offset = idx->id_offset;
    h.payload_offset = 4096;
    h.payload_bytes = (int64_t)((size_t)h.rows_padded * idx->pitch * es);
    h.norms_offset = align4k(h.payload_offset + h.payload_bytes);
    h.norms_bytes = (int64_t)idx->n * 4;
    PinnedPair pp;
    int rc = pp.init();
    if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) { set_error("write_shard: cannot open %s: %s", path, strerror(errno)); return PRS_EIO; }
    rc = fwrite(&h, sizeof(h), 1, f) == 1 ? 0 : PRS_EIO;
    if (rc) set_error("write_shard: short write to %s", path);
    if (!rc) rc = pad_file_to(f, h.payload_offset, path);
    if (!rc && h.payload_bytes) rc = device_to_file(f, idx->x, (size_t)h.payload_bytes, pp, path);
    if (!rc) rc = pad_file_to(f, h.norms_offset, path);
    if (!rc && h.norms_bytes) rc = device_to_file(f, idx->xnorm, (size_t)h.norms_bytes, pp, path);
    if (fclose(f) != 0 && !rc) { set_error("write_shard: close failed for %s", path); rc = PRS_EIO; }
    return rc;
}

// Loads a shard file into a new index, or -- when *out already holds an index -- appends it (consecutive shards of
// one rank); 16-bit images concatenate only at 64-row block boundaries.
int prs_index_read_shard(const char* path, int device, prs_index** out) {
    if (!path || !out) { set_error("read_shard: bad arguments"); return PRS_EINVAL; }
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("read_shard: cannot open %s: %s", path, strerror(errno)); return PRS_EIO; }
    ShardHeader h;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.fourcc, "PRST", 4) != 0 || h.version != 1) {
        fclose(f); set_error("read_shard: %s is not a version-1 PRST shard file", path); return PRS_EIO;
    }
    const size_t es = (size_t)elem_size(h.storage);
    const bool t64 = h.storage != PRS_F32;
    if (h.d < 1 || h.pitch != (h.d + 63) / 64 * 64 || h.rows < 0 || (h.storage != PRS_F32 && h.storage != PRS_F16 && h.storage != PRS_BF16) ||
        h.rows_padded != (t64 ? (h.rows + BLK_ROWS - 1) / BLK_ROWS * BLK_ROWS : h.rows) ||
        h.payload_bytes != (int64_t)((size_t)h.rows_padded * h.pitch * es) || h.norms_bytes != h.rows * 4 ||
        h.payload_offset < (int64_t)sizeof(h) || h.norms_offset < h.payload_offset + h.payload_bytes) {
        fclose(f); set_error("read_shard: %s has an inconsistent header", path); return PRS_EIO;
    }
    prs_index* idx = *out;
    const bool fresh = idx == nullptr;
    int rc = 0;
    if (fresh) {
        rc = prs_index_create(h.d, h.metric, h.storage, device, &idx);
        if (rc) { fclose(f); return rc; }
        idx->id_offset = h.id_offset;
    } else if (idx->d != h.d || idx->storage != h.storage || idx->metric != h.metric || idx->device != device) {
        fclose(f); set_error("read_shard: %s does not match the index it is appended to", path); return PRS_EINVAL;
    } else if (t64 && idx->n % BLK_ROWS != 0) {
        fclose(f); set_error("read_shard: cannot append %s: the index holds %lld rows, not a multiple of %d", path, idx->n, BLK_ROWS); return PRS_EINVAL;
    }
    DeviceGuard g(device);
    {
        std::lock_guard<std::mutex> lock(idx->mu);
        PRS_CUDA(cudaDeviceSynchronize());
        PinnedPair pp;
        rc = pp.init();
        if (!rc) rc = index_grow(idx, idx->n + h.rows);
        // index_grow zeroes / copies on the legacy stream (asynchronous to the host); the shard copies below run on
        // their own non-blocking stream and must not overtake them
        if (!rc && cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); set_error("read_shard: device synchronisation failed"); rc = PRS_ECUDA; }
        if (!rc && h.payload_bytes)
            rc = file_to_device(f, h.payload_offset, (unsigned char*)idx->x + (size_t)idx->n * idx->pitch * es, (size_t)h.payload_bytes, pp, path);
        if (!rc && h.norms_bytes) rc = file_to_device(f, h.norms_offset, idx->xnorm + idx->n, (size_t)h.norms_bytes, pp, path);
        if (!rc) idx->n += h.rows;
    }
    fclose(f);
    if (rc) { if (fresh) prs_index_free(idx); return rc; }
    *out = idx;
    return 0;
}

}  // extern "C"
