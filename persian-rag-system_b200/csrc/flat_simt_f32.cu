// flat_simt_f32.cu -- flat_scan_simt_kernel instantiations for float storage
#include "flat_simt_inst.cuh"
namespace prs {
int launch_simt_f32(bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    return launch_simt_t<float>(l2, QB, R, p, grid, smem, st);
}
}
