// pair_probe.cu -- does a cta_group::2 TS-mode tcgen05.mma do what the pair variant of the scan needs?
//   A (queries) in each CTA's own tensor memory (128 rows each, M = 256 over the pair), B (corpus rows) split over the two
//   CTAs' shared memories (64 rows each, N = 128), D in each CTA's tensor memory (its 128 rows x 128 columns).
// nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -o pair_probe tools/probes/pair_probe.cu && ./pair_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../persian-rag-system_b200/csrc/common.cuh"
using namespace prs;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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

constexpr int K = 64, NROWS = 64;        // one k-block; 64 corpus rows per CTA
// q: [2][128][K] half, x: [2][64][K] half (CTA r holds x[r]), out: [2][128][128] float, ticks: cycles of the timed loop
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_probe(const __half* q, const __half* x, float* out, int reps, long long* ticks, int commit_every, int walk_a) {
    __shared__ __align__(1024) unsigned char sB[NROWS * 128];
    __shared__ uint64_t bar_done;
    __shared__ uint64_t bar_scratch;     // target of the extra commits (nobody waits on it)
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int rank = (int)cluster_ctarank();
    if (tid == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_scratch, 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // B half of this CTA: T64 image (128-byte rows, 16-byte chunks swizzled by row & 7)
    for (int i = tid; i < NROWS * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(x + ((size_t)rank * NROWS + r) * K + c * 8);
        *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_ptr;
    // A: this thread's query row -> TMEM lane tid, columns [0, 32)
    {
        uint32_t v[32];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(q + ((size_t)rank * 128 + tid) * K);
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = src[c];
        tmem_st32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + 384;          // accumulator: columns [384, 512)
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // f16 x f16 -> f32, N = 128, M = 256
        const uint64_t bdesc0 = (uint64_t)((smem_u32(sB) >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const long long t0 = clock64();
        for (int it = 0; it < reps; ++it) {
            // walk_a: the A operand moves over three more 32-column groups (garbage data, timing only) before the checked one
            const int na = (walk_a && reps > 1) ? 3 : 0;
            for (int g = na; g >= 0; --g) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const uint32_t acc = k4 ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                        ::"r"(d_tmem), "r"(tmem_base + (uint32_t)(g * 32 + k4 * 8)), "l"(bdesc0 + (uint64_t)(k4 * 2)), "r"(idesc), "r"(acc) : "memory");
                }
            }
            if (commit_every > 0 && (it % commit_every) == commit_every - 1)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&bar_scratch)), "h"((uint16_t)3) : "memory");
        }
        // variant timings (results of these MMAs are discarded: they run before the checked ones only when reps < 0)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar_done)), "h"((uint16_t)3) : "memory");
        // bounded wait so that a wrong guess cannot hang the box
        long long spins = 0;
        while (!mbar_try_wait(&bar_done, 0) && ++spins < 50000000ll) { }
        ticks[0] = clock64() - t0;
        ticks[1] = spins;
    }
    {
        long long spins = 0;
        while (!mbar_try_wait(&bar_done, 0) && ++spins < 50000000ll) { }
    }
    tc_fence_after();
    for (int h = 0; h < 4; ++h) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 384 + h * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; ++c) out[((size_t)rank * 128 + tid) * 128 + h * 32 + c] = __uint_as_float(v[c]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
    std::vector<__half> hq(2 * 128 * K), hx(2 * NROWS * K);
    for (int i = 0; i < (int)hq.size(); ++i) hq[i] = __float2half((float)((i * 7 + (i >> 6)) % 5 - 2));
    for (int i = 0; i < (int)hx.size(); ++i) hx[i] = __float2half((float)((i * 3 + (i >> 5)) % 7 - 3));
    __half *dq, *dx; float* dout; long long* dt;
    cudaMalloc(&dq, hq.size() * 2); cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dout, 2 * 128 * 128 * 4); cudaMalloc(&dt, 16);
    cudaMemcpy(dq, hq.data(), hq.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
    struct V { int reps, commit_every, walk_a; };
    for (V v : {V{1, 0, 0}, V{1000, 0, 0}, V{1000, 6, 0}, V{1000, 1, 0}, V{1000, 0, 1}, V{1000, 2, 1}}) {
        const int reps = v.reps;
        cudaMemset(dout, 0xff, 2 * 128 * 128 * 4);
        pair_probe<<<2, 128>>>(dq, dx, dout, reps, dt, v.commit_every, v.walk_a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> ho(2 * 128 * 128);
        long long ht[2];
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(ht, dt, 16, cudaMemcpyDeviceToHost);
        int bad = 0, badswap = 0;
        for (int r = 0; r < 2; ++r) for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
            float ref = 0.f;
            for (int k = 0; k < K; ++k) ref += __half2float(hq[((size_t)r * 128 + m) * K + k]) * __half2float(hx[(size_t)n * K + k]);   // x rows 0..63 in CTA 0, 64..127 in CTA 1
            const float got = ho[((size_t)r * 128 + m) * 128 + n];
            if (got != ref) { if (bad < 5) printf("  mismatch cta %d m %d n %d: got %g want %g\n", r, m, n, got, ref); ++bad; }
        }
        const double nmma = 4.0 * reps * ((v.walk_a && reps > 1) ? 4 : 1);
        printf("reps %d commit every %d iterations, A walk %d: %d mismatches of %d; timed loop %lld cycles (%lld spins) -> %.1f cycles per MMA (M=256 over the pair, N=128, K=16)\n",
               reps, v.commit_every, v.walk_a, bad, 2 * 128 * 128, ht[0], ht[1], (double)ht[0] / nmma);
        (void)badswap;
    }
    return 0;
}
