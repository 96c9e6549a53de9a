// flat_simt_inst.cuh -- template dispatch shared by the three per-type translation units.
#pragma once
#include "flat_simt_launch.h"

namespace prs {
template <typename T, int QB, int R, bool L2>
static int launch_simt_one(const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kfn = flat_scan_simt_kernel<T, QB, R, L2>;
    PRS_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<grid, SIMT_THREADS, smem, st>>>(p);
    PRS_LAUNCH_CHECK();
    return 0;
}
template <typename T, int QB, bool L2>
static int launch_simt_r(int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (R) {
        case 4: return launch_simt_one<T, QB, 4, L2>(p, grid, smem, st);
        case 2: return launch_simt_one<T, QB, 2, L2>(p, grid, smem, st);
        default: return launch_simt_one<T, QB, 1, L2>(p, grid, smem, st);
    }
}
template <typename T, bool L2>
static int launch_simt_qb(int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    switch (QB) {
        case 8: return launch_simt_r<T, 8, L2>(R, p, grid, smem, st);
        case 4: return launch_simt_r<T, 4, L2>(R, p, grid, smem, st);
        case 2: return launch_simt_r<T, 2, L2>(R, p, grid, smem, st);
        default: return launch_simt_r<T, 1, L2>(R, p, grid, smem, st);
    }
}
template <typename T>
static int launch_simt_t(bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    return l2 ? launch_simt_qb<T, true>(QB, R, p, grid, smem, st) : launch_simt_qb<T, false>(QB, R, p, grid, smem, st);
}
}  // namespace prs
