"""Debug helper (GPU box): exercises the tcgen05 path on small shapes with verbose diagnostics."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import persian_rag_system_b200 as P
from oracle import oracle as O

def run(n, d, nq, k, metric, storage, path="tcgen05"):
    rng = np.random.default_rng(n + d)
    x = rng.standard_normal((n, d)).astype(np.float32); x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = rng.standard_normal((nq, d)).astype(np.float32); q /= np.linalg.norm(q, axis=1, keepdims=True)
    idx = P.FlatIndex(d, metric, storage); idx.add(x); idx.set_path(path)
    t = time.time(); D, I = idx.search(q, k); dt = time.time() - t
    dt16 = torch.float16 if storage == "fp16" else torch.bfloat16
    xs = torch.from_numpy(x).to(dt16).float().numpy()
    qs = torch.from_numpy(q).to(dt16).float().numpy() if path == "tcgen05" else q
    S = O.flat_scores_f64(xs, qs, metric)
    bad = 0
    for r in range(nq):
        try:
            O.check_topk_against_scores(I[r], D[r], S[r], k, metric == O.METRIC_IP, rtol=1e-3, atol=2e-5)
        except AssertionError as e:
            bad += 1
            if bad <= 3:
                order = np.argsort(-S[r] if metric == O.METRIC_IP else S[r])[:k]
                print("  MISMATCH q", r, "got", I[r][:6], D[r][:6], "want", order[:6], S[r][order][:6], str(e)[:200])
    print(f"n={n} d={d} nq={nq} k={k} metric={metric} {storage} {path}: bad={bad}/{nq} time={dt*1e3:.1f}ms path={idx.last_path}", flush=True)
    return bad

if __name__ == "__main__":
    tot = 0
    tot += run(64, 64, 8, 4, O.METRIC_IP, "fp16")
    tot += run(1000, 64, 8, 4, O.METRIC_IP, "fp16")
    tot += run(1000, 128, 32, 10, O.METRIC_IP, "fp16")
    tot += run(5000, 384, 64, 10, O.METRIC_IP, "fp16")
    tot += run(5000, 768, 64, 10, O.METRIC_IP, "fp16")
    tot += run(5000, 768, 128, 10, O.METRIC_L2, "fp16")
    tot += run(5000, 512, 100, 16, O.METRIC_L2, "bf16")
    tot += run(200000, 768, 64, 10, O.METRIC_IP, "fp16")
    tot += run(200000, 768, 300, 10, O.METRIC_L2, "bf16")
    tot += run(30000, 384, 129, 10, O.METRIC_IP, "fp16")       # cluster of 2, second block nearly empty
    tot += run(30000, 512, 256, 16, O.METRIC_L2, "fp16")       # cluster of 2, full
    tot += run(30000, 768, 257, 10, O.METRIC_IP, "bf16")       # cluster of 4, blocks 3 and 4 (almost) empty
    tot += run(100000, 768, 1024, 10, O.METRIC_IP, "fp16")     # cluster of 4, two passes
    tot += run(100, 64, 600, 5, O.METRIC_L2, "fp16")           # fewer tiles than clusters
    # wide k (16 < k <= 1024): sample -> threshold -> collect -> select
    tot += run(200000, 768, 64, 100, O.METRIC_IP, "fp16")
    tot += run(200000, 384, 1, 17, O.METRIC_L2, "fp16")
    tot += run(100000, 512, 300, 100, O.METRIC_L2, "bf16")
    tot += run(300000, 384, 5, 1024, O.METRIC_IP, "fp16")
    tot += run(40000, 768, 130, 33, O.METRIC_L2, "fp16")
    print("TOTAL BAD", tot)
    sys.exit(1 if tot else 0)
