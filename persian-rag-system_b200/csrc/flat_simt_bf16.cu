// flat_simt_bf16.cu -- flat_scan_simt_kernel instantiations for __nv_bfloat16 storage
#include "flat_simt_inst.cuh"
namespace prs {
int launch_simt_bf16(bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    return launch_simt_t<__nv_bfloat16>(l2, QB, R, p, grid, smem, st);
}
}
