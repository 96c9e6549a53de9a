// hop_probe.cu -- latency of signalling between the two CTAs of a cluster: remote mbarrier arrive + try_wait, and a plain
// distributed-shared-memory flag (st.release.cluster / ld.acquire.cluster).  1000 ping-pongs each.
#include <cstdio>
#include "../../persian-rag-system_b200/csrc/common.cuh"
using namespace prs;
__device__ __forceinline__ uint32_t mapa_u32(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32) hop_probe(long long* out, int mode) {
    __shared__ uint64_t bar;
    __shared__ uint32_t flag;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); flag = 0; }
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x == 0) {
        const uint32_t peer_bar = mapa_u32(smem_u32(&bar), rank ^ 1u), peer_flag = mapa_u32(smem_u32(&flag), rank ^ 1u);
        const int N = 1000;
        const long long t0 = clock64();
        if (mode == 0) {
            // rank 0 arrives on rank 1's barrier, rank 1 answers on rank 0's: N round trips
            uint32_t ph = 0;
            for (int i = 0; i < N; ++i) {
                if (rank == 0) {
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(peer_bar) : "memory");
                    while (!mbar_try_wait(&bar, ph)) { }
                } else {
                    while (!mbar_try_wait(&bar, ph)) { }
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(peer_bar) : "memory");
                }
                ph ^= 1u;
            }
        } else {
            for (int i = 1; i <= N; ++i) {
                uint32_t v;
                if (rank == 0) {
                    asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(peer_flag), "r"((uint32_t)i) : "memory");
                    do { asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory"); } while (v != (uint32_t)i);
                } else {
                    do { asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory"); } while (v != (uint32_t)i);
                    asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(peer_flag), "r"((uint32_t)i) : "memory");
                }
            }
        }
        out[rank] = clock64() - t0;
    }
    __syncthreads();
    cluster_sync_all();
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    for (int mode = 0; mode < 2; ++mode) {
        hop_probe<<<2, 32>>>(d, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s: %.0f cycles per round trip (two hops)\n", mode == 0 ? "remote mbarrier arrive + try_wait" : "DSMEM flag st.release / ld.acquire", (double)h[0] / 1000.0);
    }
    return 0;
}
