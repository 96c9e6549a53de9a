"""Stress: the one-launch search must equal the three-kernel sequence bit for bit, every time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import persian_rag_system_b200 as P
dev = torch.device("cuda", 0)
bad = 0
for (n, d, nq, k, storage, metric) in [(1000, 200, 9, 7, "bf16", 1), (1000, 200, 9, 7, "fp16", 1), (20000, 768, 64, 10, "fp16", 0), (100000, 384, 128, 16, "bf16", 1),
                                       (125, 384, 1, 5, "fp16", 1), (300000, 768, 64, 10, "fp16", 0), (5000, 64, 33, 3, "fp16", 0)]:
    rng = np.random.default_rng(n + d)
    x = rng.standard_normal((n, d)).astype(np.float32)
    idx = P.FlatIndex(d, metric, storage); idx.add(x)
    ref = P.FlatIndex(d, metric, storage); ref.add(x); ref.set_fused(False)
    fails = 0
    for rep in range(200):
        q = torch.from_numpy(rng.standard_normal((nq, d)).astype(np.float32)).to(dev)
        D, I = idx.search(q, k)
        Dr, Ir = ref.search(q, k)
        assert idx.last_fused and not ref.last_fused
        if not (torch.equal(I, Ir) and torch.equal(D, Dr)):
            fails += 1
            if fails == 1:
                r = int((I != Ir).any(1).nonzero()[0]) if (I != Ir).any() else 0
                print("  first mismatch rep", rep, "row", r, I[r].tolist(), Ir[r].tolist(), D[r].tolist(), Dr[r].tolist())
    print(f"n={n} d={d} nq={nq} k={k} {storage} metric={metric}: {fails}/200 mismatches", flush=True)
    bad += fails
# timing A/B at the headline shape
n, d, nq, k = 1_000_000, 768, 64, 10
idx = P.IndexFlatIP(d, storage="fp16"); idx.reserve(n)
g = torch.Generator(device=dev).manual_seed(1)
for _ in range(4):
    xb = torch.randn(n // 4, d, generator=g, device=dev); xb /= xb.norm(dim=1, keepdim=True); idx.add(xb.half())
q = torch.randn(nq, d, generator=g, device=dev)
for fused in (True, False, True, False):
    idx.set_fused(fused)
    for _ in range(5): idx.search(q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): idx.search(q, k)
    e1.record(); torch.cuda.synchronize()
    print(f"fused={fused}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per search")
print("TOTAL BAD", bad)
