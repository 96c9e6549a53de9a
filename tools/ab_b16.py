"""50M x 384 fp16: the B = 16, k = 10 point of the capacity sweep measured 25-40 % slower than its neighbours -- which (B, k) and why?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import persian_rag_system_b200 as P
dev = torch.device("cuda", 0)
n, d = 50_000_000, 384
idx = P.IndexFlatIP(d, storage="fp16"); idx.reserve(n)
g = torch.Generator(device=dev).manual_seed(1)
for o in range(0, n, 2_000_000):
    idx.add(torch.randn(2_000_000, d, generator=g, device=dev, dtype=torch.float16))
def t(q, k, reps=5):
    for _ in range(2): idx.search(q, k)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx.search(q, k); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2]
for seed in (1, 2):
    gq = torch.Generator(device=dev).manual_seed(100 + seed)
    for B in (4, 8, 16, 32, 64):
        q = torch.randn(B, d, generator=gq, device=dev)
        print(f"seed {seed} B={B}: " + "  ".join(f"k={k}: {t(q, k):.3f}" for k in (1, 5, 10, 16)), flush=True)
q = torch.randn(16, d, generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
print("unit-norm queries B=16: " + "  ".join(f"k={k}: {t(q, k):.3f}" for k in (1, 10)))
