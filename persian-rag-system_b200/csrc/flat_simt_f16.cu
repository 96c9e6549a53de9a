// flat_simt_f16.cu -- flat_scan_simt_kernel instantiations for __half storage
#include "flat_simt_inst.cuh"
namespace prs {
int launch_simt_f16(bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    return launch_simt_t<__half>(l2, QB, R, p, grid, smem, st);
}
}
