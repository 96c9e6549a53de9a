import os, sys, torch
sys.path.insert(0, os.getcwd())
import persian_rag_system_b200 as P
dev = torch.device("cuda", 0)
d, n, B = int(os.environ.get("DD", "768")), 1_000_000, 256
idx = P.IndexFlatIP(d, storage="fp16"); idx.reserve(n)
g = torch.Generator(device=dev).manual_seed(1)
for _ in range(4):
    xb = torch.randn(n // 4, d, generator=g, device=dev); xb /= xb.norm(dim=1, keepdim=True); idx.add(xb.half())
q = torch.randn(B, d, generator=g, device=dev)
for _ in range(2): idx.search(q, 10)
torch.cuda.synchronize()
