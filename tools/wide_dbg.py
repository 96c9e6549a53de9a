import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import persian_rag_system_b200 as P
rng = np.random.default_rng(0)
n, d = 200000, 768
x = rng.standard_normal((n, d)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True)
q = rng.standard_normal((4, d)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
idx = P.FlatIndex(d, P.METRIC_IP, "fp16"); idx.add(x)
D, I = idx.search(q, 100)
print(idx.last_path, D[0, :3], D[0, 97:])
S = (torch.from_numpy(x).half().float().numpy() @ torch.from_numpy(q).half().float().numpy().T)
print("true 100th:", np.sort(S[:, 0])[::-1][99], "16th", np.sort(S[:, 0])[::-1][15])
