// flat_simt.cuh -- bandwidth-bound flat scan on CUDA cores with a fused per-query top-k.
//
// Replaces faiss IndexFlat::search for small query batches (reference call site
// src/retrieval.py:102 always has nq = 1) and is the exact-parity fp32 path: squared L2 is
// computed in the direct form sum (q - x)^2 like faiss does for nq < 20, so near-duplicate
// rows (fine-tuned indices, L2^2 ~ 5e-5) do not suffer the cancellation of the expanded form.
//
// Data flow per CTA (persistent, one per SM):
//   producer warp : cp.async.bulk (TMA, 1-D) streams tiles of `tile_rows` contiguous corpus rows
//                   HBM -> shared memory through an mbarrier ring of `stages` buffers;
//   8 consumer warps: each takes R rows of the tile at a time; lanes stride the row in 16-byte
//                   pieces (LDS.128, conflict free), multiply against up to QB queries held in
//                   shared memory (fp32), accumulate in fp32; a transposed butterfly reduction
//                   leaves ONE (row, query) total per lane group; survivors (score >= the warp's
//                   current k-th best for that query) are appended as 64-bit keys to a per-warp
//                   candidate list, which is compacted by a warp-level bitonic sort when it fills;
//   epilogue      : the CTA merges its warps' lists into one sorted top-k per query
//                   (cand[cta][query][k]); merge_cand_kernel then merges the CTAs.
// No score ever reaches HBM except the few survivors.
#pragma once
#include "common.cuh"

namespace prs {

constexpr int SIMT_NW = 8;                       // consumer warps
constexpr int SIMT_THREADS = SIMT_NW * 32 + 32;  // + producer warp
constexpr int SIMT_MAX_STAGES = 8;

struct SimtParams {
    const void* x;       // [n_rows, pitch] storage dtype
    const float* q;      // first query of this group, fp32, row stride q_stride
    long long n_rows;
    int d, pitch, q_stride;
    int nq;              // valid queries in this group (1..QB)
    int k, cap;          // cap: per-warp list capacity (power of two, >= k + R)
    int tile_rows, stages;
    u64* lists;          // [grid*SIMT_NW][QB][cap]
    u64* cand;           // [grid][nq_total][k]
    int* cand_cnt;       // [grid][nq_total]
    int nq_total, q0;    // q0: index of this group's first query in the whole batch
    int sortn;           // power of two >= k + SIMT_THREADS
    // single-CTA searches (an index of a few tiles: the reference's own 125-row indices) write the result themselves
    // instead of leaving one candidate list for merge_cand_kernel: one launch per search
    float* D;            // [nq_total, k]
    long long* I;
    long long id_offset;
    int direct;          // 1: grid == 1, write D / I here;  out_mode 0: D = score (IP), 1: D = -score (direct-form L2)
    int out_mode;
};

template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void load(const unsigned char* p, float (&f)[4]) {
        float4 v = *reinterpret_cast<const float4*>(p);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
};
template <> struct Vec16<__half> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void load(const unsigned char* p, float (&f)[8]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void load(const unsigned char* p, float (&f)[8]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
};

// transposed butterfly: N per-lane partial sums -> lane L ends with the warp total of
// accumulator index (L >> (5 - log2 N)).  N-1 + (5 - log2 N) shuffles instead of 5*N.
template <int N, int OFF>
struct XReduce {
    __device__ static __forceinline__ float run(float (&v)[N], int lane) {
        constexpr int H = N / 2;
        float w[H];
        const bool upper = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float send = upper ? v[i] : v[i + H];
            const float keep = upper ? v[i + H] : v[i];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        return XReduce<H, OFF / 2>::run(w, lane);
    }
};
template <int OFF>
struct XReduce<1, OFF> {
    __device__ static __forceinline__ float run(float (&v)[1], int lane) {
        float x = v[0];
#pragma unroll
        for (int o = OFF; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        return x;
    }
};
template <int N> struct Log2 { static constexpr int v = 1 + Log2<N / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// in-place warp bitonic sort (descending) of n (power of two) keys in global/shared memory
__device__ __forceinline__ void warp_sort_mem_desc(u64* s, int n, int lane) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int p = lane; p < (n >> 1); p += 32) {
                const int i = ((p & ~(stride - 1)) << 1) | (p & (stride - 1));
                const int j = i + stride;
                const bool desc = ((i & size) == 0);
                u64 a = s[i], b = s[j];
                const bool sw = desc ? (a < b) : (a > b);
                if (sw) { s[i] = b; s[j] = a; }
            }
            __syncwarp();
        }
    }
}

// compacts a warp-private candidate list to its k best (sorted descending, in place).
// Returns the new admission threshold (score of the k-th best), or -inf if fewer than k.
__device__ __forceinline__ float warp_compact(u64* list, int& cnt, int k, int cap, int lane) {
    __syncwarp();
    u64 kth = 0;
    if (cap == 128) {
        u64 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int e = lane * 4 + i; v[i] = (e < cnt) ? list[e] : 0ull; }
        warp_sort_desc<4>(v, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int e = lane * 4 + i; if (e < k) list[e] = v[i]; }
        const int r = (k - 1) & 3;
        u64 mine = r == 0 ? v[0] : (r == 1 ? v[1] : (r == 2 ? v[2] : v[3]));
        kth = shfl_u64(mine, (k - 1) >> 2);
    } else {
        for (int e = cnt + lane; e < cap; e += 32) list[e] = 0ull;
        __syncwarp();
        warp_sort_mem_desc(list, cap, lane);
        kth = list[k - 1];
    }
    __syncwarp();
    if (cnt > k) cnt = k;
    return (cnt == k) ? key_score(kth) : -INFINITY;
}

template <typename T, int QB, int R, bool L2>
__global__ void __launch_bounds__(SIMT_THREADS, 1) flat_scan_simt_kernel(const SimtParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int VEC = Vec16<T>::N;
    constexpr bool TILED = (VEC == 8);                // 16-bit corpora are stored in the T64 layout (common.cuh)
    constexpr int NACC = R * QB;
    constexpr int P = Log2<NACC>::v;
    static_assert(NACC <= 32, "at most 32 (row, query) accumulators per lane");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row_bytes = p.pitch * (int)sizeof(T);
    const int tile_bytes = p.tile_rows * row_bytes;

    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + SIMT_MAX_STAGES;
    int* s_cnts = reinterpret_cast<int*>(smem + 128);                 // [SIMT_NW][QB] + 1
    float* sq = reinterpret_cast<float*>(smem + 512);                 // QB * pitch floats
    unsigned char* tiles = smem + 512 + (((size_t)QB * p.pitch * 4 + 127) & ~(size_t)127);

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SIMT_NW); }
        mbar_fence_init();
    }
    // queries -> shared memory (fp32).  For 16-bit storage a lane consumes 8 consecutive
    // elements per step as two float4; they are stored in two planes so that each LDS.128 has
    // the 32 lanes at a 16-byte stride (no bank conflicts).
    for (int i = tid; i < QB * p.pitch; i += SIMT_THREADS) {
        const int qq = i / p.pitch, c = i - qq * p.pitch;
        const float v = (qq < p.nq && c < p.d) ? p.q[(size_t)qq * p.q_stride + c] : 0.f;
        int pos;
        if (VEC == 8) pos = ((c & 7) >> 2) * (p.pitch >> 1) + ((c >> 3) << 2) + (c & 3);
        else pos = c;
        sq[qq * p.pitch + pos] = v;
    }
    __syncthreads();

    const long long n_tiles = (p.n_rows + p.tile_rows - 1) / p.tile_rows;
    const unsigned char* xg = reinterpret_cast<const unsigned char*>(p.x);

    int cnt[QB];
    float thr[QB];
#pragma unroll
    for (int i = 0; i < QB; ++i) { cnt[i] = 0; thr[i] = -INFINITY; }
    const int part = blockIdx.x * SIMT_NW + warp;

    if (warp == SIMT_NW) {
        // ---------------- producer: one elected lane drives the TMA ring ----------------
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&empty[s], ph ^ 1u);
                const long long row0 = t * p.tile_rows;
                long long rows = (p.n_rows - row0 < p.tile_rows) ? (p.n_rows - row0) : p.tile_rows;
                if (TILED) rows = (rows + BLK_ROWS - 1) / BLK_ROWS * BLK_ROWS;        // whole row blocks
                const uint32_t bytes = (uint32_t)(rows * row_bytes);
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(tiles + (size_t)s * tile_bytes, xg + (size_t)row0 * row_bytes, bytes, &full[s]);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int a = lane >> (5 - P);                      // accumulator this lane ends up owning
        const int my_r = a / QB, my_q = a % QB;
        const bool leader = (lane & ((1 << (5 - P)) - 1)) == 0;
        float my_thr = -INFINITY;
        u64* my_lists = p.lists + (size_t)part * QB * p.cap;
        const int steps = (p.pitch + 32 * VEC - 1) / (32 * VEC);
        const int half_pitch = p.pitch >> 1;
        int s = 0;
        uint32_t ph = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(&full[s], ph);
            const unsigned char* tile = tiles + (size_t)s * tile_bytes;
            const long long row0 = t * p.tile_rows;
            for (int rg = warp; rg * R < p.tile_rows; rg += SIMT_NW) {
                float acc[NACC];
#pragma unroll
                for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
                for (int st = 0; st < steps; ++st) {
                    const int chunk = st * 32 + lane;
                    const int col = chunk * VEC;
                    if (col < p.pitch) {
                        float xv[R][VEC];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            if (TILED) Vec16<T>::load(tile + t64_offset(rg * R + r, chunk, p.pitch), xv[r]);
                            else Vec16<T>::load(tile + (size_t)(rg * R + r) * row_bytes + (size_t)chunk * 16, xv[r]);
                        }
#pragma unroll
                        for (int qq = 0; qq < QB; ++qq) {
                            float qv[VEC];
                            if constexpr (VEC == 8) {
                                const float4 lo = *reinterpret_cast<const float4*>(sq + qq * p.pitch + chunk * 4);
                                const float4 hi = *reinterpret_cast<const float4*>(sq + qq * p.pitch + half_pitch + chunk * 4);
                                qv[0] = lo.x; qv[1] = lo.y; qv[2] = lo.z; qv[3] = lo.w;
                                qv[4] = hi.x; qv[5] = hi.y; qv[6] = hi.z; qv[7] = hi.w;
                            } else {
                                const float4 lo = *reinterpret_cast<const float4*>(sq + qq * p.pitch + col);
                                qv[0] = lo.x; qv[1] = lo.y; qv[2] = lo.z; qv[3] = lo.w;
                            }
#pragma unroll
                            for (int r = 0; r < R; ++r) {
#pragma unroll
                                for (int e = 0; e < VEC; ++e) {
                                    if (L2) { const float dlt = qv[e] - xv[r][e]; acc[r * QB + qq] = fmaf(dlt, dlt, acc[r * QB + qq]); }
                                    else acc[r * QB + qq] = fmaf(xv[r][e], qv[e], acc[r * QB + qq]);
                                }
                            }
                        }
                    }
                }
                const float tot = XReduce<NACC, 16>::run(acc, lane);
                const long long row = row0 + (long long)rg * R + my_r;
                const float sc = sanitize(L2 ? -tot : tot);
                const bool pass = leader && (row < p.n_rows) && (my_q < p.nq) && (sc >= my_thr);
#pragma unroll
                for (int qq = 0; qq < QB; ++qq) {
                    const unsigned m = __ballot_sync(0xffffffffu, pass && my_q == qq);
                    if (m) {
                        if (pass && my_q == qq) {
                            const int slot = cnt[qq] + __popc(m & ((1u << lane) - 1u));
                            my_lists[(size_t)qq * p.cap + slot] = make_key<PRS_TIE_LOW_ID>(sc, (uint32_t)row);
                        }
                        cnt[qq] += __popc(m);
                        if (cnt[qq] > p.cap - R) {
                            thr[qq] = warp_compact(my_lists + (size_t)qq * p.cap, cnt[qq], p.k, p.cap, lane);
                            if (my_q == qq) my_thr = thr[qq];
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        if (lane == 0) {
#pragma unroll
            for (int qq = 0; qq < QB; ++qq) s_cnts[warp * QB + qq] = cnt[qq];
        }
    }
    __syncthreads();   // all copies landed and consumed; tile memory is reused below

    // ---------------- CTA epilogue: merge the 8 warps' lists, one sorted top-k per query ------
    u64* buf = reinterpret_cast<u64*>(tiles);
    int* s_n = s_cnts + SIMT_NW * QB;
    for (int qq = 0; qq < p.nq; ++qq) {
        const u64* base = p.lists + ((size_t)blockIdx.x * SIMT_NW * QB + qq) * p.cap;
        const int cap = p.cap;
        const size_t o = ((size_t)blockIdx.x * p.nq_total + p.q0 + qq);
        int total = 0;
#pragma unroll
        for (int w = 0; w < SIMT_NW; ++w) total += s_cnts[w * QB + qq];
        auto emit = [&](int j, u64 key) {            // slot j of this query's result (key 0 = empty)
            if (p.direct) {
                const size_t r = (size_t)(p.q0 + qq) * p.k + j;
                if (key) {
                    const float sc = key_score(key);
                    p.D[r] = p.out_mode == 0 ? sc : -sc;
                    p.I[r] = (long long)key_id<PRS_TIE_LOW_ID>(key) + p.id_offset;
                } else {
                    p.D[r] = p.out_mode == 0 ? -3.402823466e+38f : 3.402823466e+38f;
                    p.I[r] = -1;
                }
            } else {
                p.cand[o * p.k + j] = key;             // all k slots, 0 = empty
            }
        };
        if (total <= 128) {
            // few candidates (small indices): one warp gathers them into registers and sorts them, no block-wide rounds
            if (warp == 0) {
                u64 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = lane * 4 + i;
                    int off = 0;
                    u64 val = 0ull;
#pragma unroll
                    for (int w = 0; w < SIMT_NW; ++w) {
                        const int c = s_cnts[w * QB + qq];
                        if (e >= off && e < off + c) val = base[(size_t)w * QB * cap + (e - off)];
                        off += c;
                    }
                    v[i] = val;
                }
                warp_sort_desc<4>(v, lane);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const int e = lane * 4 + i; if (e < p.k) emit(e, v[i]); }
                if (lane == 0 && !p.direct) p.cand_cnt[o] = total < p.k ? total : p.k;
            }
            for (int j = 128 + tid; j < p.k; j += SIMT_THREADS) emit(j, 0ull);      // k > 128: the remaining slots are empty
            __syncthreads();
            continue;
        }
        auto fetch = [&](long long i) -> u64 {
            const int w = (int)(i / cap), j = (int)(i - (long long)w * cap);
            return (j < s_cnts[w * QB + qq]) ? base[(size_t)w * QB * cap + j] : 0ull;
        };
        const int n = block_topk_stream(fetch, (long long)SIMT_NW * cap, p.k, buf, p.sortn, s_n, tid, SIMT_THREADS, 1);
        for (int j = tid; j < p.k; j += SIMT_THREADS) emit(j, j < n ? buf[j] : 0ull);
        if (tid == 0 && !p.direct) p.cand_cnt[o] = n;
        __syncthreads();
    }
}

}  // namespace prs
