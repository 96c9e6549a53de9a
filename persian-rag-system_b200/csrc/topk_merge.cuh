// topk_merge.cuh -- final merge of the per-CTA candidate lists (one CTA per query).
#pragma once
#include "common.cuh"

namespace prs {

// ---------------------------------------------------------------------------------------------
// Final merge over CTAs ("parts"): one CTA per query.
// out_mode 0: D = score (IP)   1: D = -score (direct L2)   2: D = max(0, qnorm - score) (expanded L2)
// ---------------------------------------------------------------------------------------------
constexpr int MERGE_THREADS = 256;

__global__ void __launch_bounds__(MERGE_THREADS) merge_cand_kernel(
    const u64* __restrict__ cand, const int* __restrict__ cand_cnt, int parts, int nq, int k, int sortn,
    int out_mode, const float* __restrict__ qnorm, long long id_offset, float* __restrict__ D,
    long long* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char msm[];
    u64* buf = reinterpret_cast<u64*>(msm);
    int* s_n = reinterpret_cast<int*>(msm + (size_t)sortn * 8);
    const int q = blockIdx.x, tid = threadIdx.x;
    auto fetch = [&](long long i) -> u64 {
        const int part = (int)(i / k), j = (int)(i - (long long)part * k);
        const size_t o = (size_t)part * nq + q;
        return (j < cand_cnt[o]) ? cand[o * k + j] : 0ull;
    };
    const int n = block_topk_stream(fetch, (long long)parts * k, k, buf, sortn, s_n, tid, MERGE_THREADS, 1);
    for (int j = tid; j < k; j += MERGE_THREADS) {
        float dv;
        long long iv;
        if (j < n) {
            const u64 key = buf[j];
            const float s = key_score(key);
            dv = out_mode == 0 ? s : (out_mode == 1 ? -s : fmaxf(0.f, qnorm[q] - s));
            iv = (long long)key_id<PRS_TIE_LOW_ID>(key) + id_offset;
        } else {
            dv = out_mode == 0 ? -3.402823466e+38f : 3.402823466e+38f;
            iv = -1;
        }
        D[(size_t)q * k + j] = dv;
        I[(size_t)q * k + j] = iv;
    }
}


}  // namespace prs
