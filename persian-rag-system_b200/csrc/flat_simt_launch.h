// flat_simt_launch.h -- per-storage-type launchers of flat_scan_simt_kernel (one TU each so the
// 24 instantiations per type build in parallel).
#pragma once
#include "flat_simt.cuh"
#include "host_common.h"

namespace prs {
int launch_simt_f32(bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st);
int launch_simt_f16(bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st);
int launch_simt_bf16(bool l2, int QB, int R, const SimtParams& p, int grid, size_t smem, cudaStream_t st);
}
