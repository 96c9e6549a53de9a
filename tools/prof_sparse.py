"""One sparse search of the C4 shape for ncu (tools/bench_aux.py generates the same data)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import persian_rag_system_b200 as P
from tools.bench_aux import gen_sparse

docs, queries, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
dev = torch.device("cuda", 0)
indptr, indices, values, cdf, df = gen_sparse(docs, 200_000, 7, dev)
idx = P.SparseIndex(indptr, indices, values, 200_000, mode=mode)
rng = np.random.default_rng(11)
qlen = 1 + rng.poisson(6, size=queries)
q_indptr = np.zeros(queries + 1, np.int64)
q_indptr[1:] = np.cumsum(qlen)
u = torch.from_numpy(rng.random(int(q_indptr[-1]))).to(dev)
q_terms = torch.searchsorted(cdf, u).clamp_(max=199_999).to(torch.int32)
d_ip = torch.from_numpy(q_indptr).to(dev)
d_qw = torch.ones(q_terms.shape[0], dtype=torch.float64, device=dev)
for _ in range(2):
    S, I = idx.search_device(d_ip, q_terms, d_qw, 10)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); S, I = idx.search_device(d_ip, q_terms, d_qw, 10); e1.record(); torch.cuda.synchronize()
print(f"{mode}: {e0.elapsed_time(e1):.3f} ms for {queries} queries on {docs} docs; postings {idx.last_postings}")
