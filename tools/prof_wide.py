"""One wide-k search for an ncu launch list: python tools/prof_wide.py ROWS D BATCH K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import persian_rag_system_b200 as P
rows, d, B, k = (int(v) for v in sys.argv[1:5])
dev = torch.device("cuda", 0)
idx = P.IndexFlatIP(d, storage="fp16")
idx.reserve(rows)
g = torch.Generator(device=dev).manual_seed(1)
done = 0
while done < rows:
    c = min(1 << 20, rows - done)
    x = torch.randn(c, d, generator=g, device=dev); x /= x.norm(dim=1, keepdim=True)
    idx.add(x.half()); done += c
q = torch.randn(B, d, generator=g, device=dev); q /= q.norm(dim=1, keepdim=True)
for _ in range(2):
    idx.search(q, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); D, I = idx.search(q, k); e1.record(); torch.cuda.synchronize()
print(f"rows={rows} d={d} B={B} k={k}: {e0.elapsed_time(e1):.3f} ms")
