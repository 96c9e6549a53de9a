"""The reference's own call shape: 125 x 384 fp32 IndexFlatL2, nq = 1, k = 5 (src/retrieval.py:102).
python tools/prof_small.py [rows]  -- latency of FlatIndex.search(numpy) and of the raw C-ABI call."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import persian_rag_system_b200 as P
from persian_rag_system_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125
d, k = 384, 5
rng = np.random.default_rng(0)
x = rng.standard_normal((n, d)).astype(np.float32)
idx = P.IndexFlatL2(d); idx.add(x)
qs = (x[rng.integers(0, n, 2000)] + 0.01 * rng.standard_normal((2000, d))).astype(np.float32)
for i in range(200): idx.search(qs[i:i + 1], k)
ts = []
for i in range(2000):
    t0 = time.perf_counter(); D, I = idx.search(qs[i:i + 1], k); ts.append(time.perf_counter() - t0)
ts = np.sort(np.array(ts)) * 1e6
print(f"{n} x {d} fp32, nq=1, k={k}: FlatIndex.search(numpy) p50 {ts[1000]:.1f} us p99 {ts[1980]:.1f} us")
L = _lib.lib()
D = np.empty((1, k), np.float32); I = np.empty((1, k), np.int64)
ts = []
for i in range(2000):
    q = qs[i:i + 1]
    t0 = time.perf_counter()
    L.prs_index_search_host(idx._h, q.ctypes.data_as(ctypes.c_void_p), 1, k, D.ctypes.data_as(ctypes.c_void_p), I.ctypes.data_as(ctypes.c_void_p))
    ts.append(time.perf_counter() - t0)
ts = np.sort(np.array(ts)) * 1e6
print(f"raw prs_index_search_host p50 {ts[1000]:.1f} us p99 {ts[1980]:.1f} us")
