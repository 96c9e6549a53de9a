// ivf.cu -- the one piece of the reference's IVF-Flat branch (scripts/phase3_pdf_chunking.py:45-57,
// faiss.IndexIVFFlat(IndexFlatL2(d), d, nlist); index.train(first 10 000 rows)) that the flat-scan kernels do not
// already provide: the centroid update of k-means.  Assignment (nearest centroid), the coarse probe and the scan of
// an inverted list are all exact flat searches and run on the kernels of flat_simt.cuh / flat_umma.cuh (ivf.py).
//
// faiss 1.7.4 Clustering::train -> compute_centroids sums the points of every centroid in POINT ORDER in float32
// (its OpenMP split is over centroid ranges, every thread walks all points), then divides by the count.  One thread
// per (centroid, dimension) walking the assignment array reproduces exactly that sum order, so the centroids are
// bit-identical to a sequential float32 restatement (oracle/oracle.py::IVFFlatOracle) given the same assignment.
#include "common.cuh"
#include "host_common.h"

namespace prs {

__global__ void centroid_update_kernel(const float* __restrict__ x, long long n, int d, const long long* __restrict__ assign, int k,
                                       float* __restrict__ centroids, long long* __restrict__ counts) {
    const int c = blockIdx.x;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        float s = 0.f;
        long long cnt = 0;
        for (long long i = 0; i < n; ++i) {
            if (__ldg(assign + i) == c) { s += __ldg(x + (size_t)i * d + j); ++cnt; }
        }
        // an empty cluster keeps its previous centroid (faiss re-seeds it by splitting a big cluster; ivf.py reports it)
        if (cnt > 0) centroids[(size_t)c * d + j] = s * (1.f / (float)cnt);
        if (j == 0) counts[c] = cnt;
    }
}

}  // namespace prs

using namespace prs;

extern "C" int prs_centroid_update_device(const float* x, int64_t n, int d, const int64_t* assign, int k, float* centroids,
                                          int64_t* counts, int device, void* stream) {
    if (!x || !assign || !centroids || !counts || n < 0 || d < 1 || k < 1) { set_error("centroid_update: bad arguments"); return PRS_EINVAL; }
    DeviceGuard g(device);
    centroid_update_kernel<<<(unsigned)k, 128, 0, (cudaStream_t)stream>>>(x, n, d, (const long long*)assign, k, centroids, (long long*)counts);
    PRS_LAUNCH_CHECK();
    return 0;
}
