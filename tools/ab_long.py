"""A/B of two builds on long corpora: python tools/ab_long.py  (PRS_LIB_PATH selects the build)
50M x 384 fp16 (configs[4] share): B x k cases; then 20M x 512 k = 100 (wide k)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import persian_rag_system_b200 as P
dev = torch.device("cuda", 0)
tag = os.path.basename(os.environ.get("PRS_LIB_PATH", "libprs.so"))


def build(n, d):
    idx = P.IndexFlatIP(d, storage="fp16"); idx.reserve(n)
    g = torch.Generator(device=dev).manual_seed(1)
    step = 2_000_000
    for o in range(0, n, step):
        xb = torch.randn(min(step, n - o), d, generator=g, device=dev, dtype=torch.float16)
        idx.add(xb)
    return idx, g


def timeit(idx, q, k, reps):
    for _ in range(2): idx.search(q, k)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx.search(q, k); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


idx, g = build(50_000_000, 384)
for B, k in ((1, 10), (16, 1), (16, 10), (64, 10), (256, 10), (4096, 10), (16, 100), (256, 100)):
    q = torch.randn(B, 384, generator=g, device=dev)
    med, best = timeit(idx, q, k, 3 if B >= 4096 else 7)
    print(f"{tag} 50M x 384 B={B} k={k}: median {med:.3f} ms (best {best:.3f})", flush=True)
del idx
torch.cuda.empty_cache()
idx, g = build(20_000_000, 512)
for B, k in ((64, 100), (1024, 100)):
    q = torch.randn(B, 512, generator=g, device=dev)
    med, best = timeit(idx, q, k, 5)
    print(f"{tag} 20M x 512 B={B} k={k}: median {med:.3f} ms (best {best:.3f})", flush=True)
