#!/bin/bash
# compute-sanitizer over the smoke shapes of every kernel family (scan, merge, sparse, pool, hybrid, wide k): memcheck + racecheck.
# usage (GPU box): bash tools/sanitize.sh > gpurun_out/sanitizer.log
set -u
cat > /tmp/san_smoke.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import persian_rag_system_b200 as P
rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)
# tcgen05 scan, one-launch search (L2 re-rank) and the three-kernel sequence
x = rng.standard_normal((3000, 256)).astype(np.float32)
q = rng.standard_normal((9, 256)).astype(np.float32)
for fused in (True, False):
    for metric in (P.METRIC_L2, P.METRIC_INNER_PRODUCT):
        idx = P.FlatIndex(256, metric, "fp16"); idx.add(x); idx.set_fused(fused)
        idx.search(torch.from_numpy(q).to(dev), 5); idx.search(q, 5)
# clusters (nq > 128), wide k, fp32 CUDA-core scan
idx = P.FlatIndex(128, P.METRIC_L2, "bf16"); idx.add(rng.standard_normal((20000, 128)).astype(np.float32))
idx.search(rng.standard_normal((130, 128)).astype(np.float32), 4)
idx.search(rng.standard_normal((3, 128)).astype(np.float32), 40)
f32 = P.IndexFlatL2(96); f32.add(rng.standard_normal((700, 96)).astype(np.float32)); f32.search(rng.standard_normal((3, 96)).astype(np.float32), 5)
# sparse (both kernels), hybrid fusion, pooling
docs = [[f"w{int(t)}" for t in rng.integers(0, 80, size=int(rng.integers(1, 30)))] for _ in range(5000)]
for mode in ("exact", "throughput"):
    bm = P.BM25Index(docs, mode=mode)
    S, I = bm.search_device([docs[3][:4], docs[9][:2], ["zz"]], 5)
D = torch.rand(3, 10, device=dev); Id = torch.randint(0, 5000, (3, 10), device=dev)
P.hybrid_fuse(D, Id, S.repeat(1, 2)[:, :10].contiguous(), I.repeat(1, 2)[:, :10].contiguous(), 5000, 5)
h = torch.randn(4, 16, 384, device=dev).half(); m = torch.ones(4, 16, dtype=torch.int64, device=dev)
P.mean_pool_normalize(h, m, True)
torch.cuda.synchronize()
print("smoke done")
PY
for tool in memcheck racecheck; do
  echo "===== compute-sanitizer --tool $tool ====="
  timeout 900 compute-sanitizer --tool $tool --kernel-name-exclude regex:"at::|cub::|elementwise|reduce_kernel|distribution" python /tmp/san_smoke.py 2>&1 | grep -v "^$" | tail -25
done
