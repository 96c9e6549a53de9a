#!/bin/bash
# large-batch decomposition on the experiments build: which pipeline stage bounds B >= 256?
# PRS_UMMA_DEBUG bits: 1 no bootstrap / no insertions, 2 epilogue releases the accumulator without reading it
# usage: tools/bigbatch_decomp.sh ["0 1 2"]   (debug values to walk)
export PRS_LIB_PATH=$PWD/persian-rag-system_b200/libprs_x.so
DBGS="${1:-0 1 2}" python - <<'PY'
import os, sys, subprocess
code = r'''
import os, torch, sys
sys.path.insert(0, os.getcwd())
import persian_rag_system_b200 as P
dev = torch.device("cuda", 0)
d, n = int(os.environ.get("DD", "768")), 1_000_000
idx = P.IndexFlatIP(d, storage="fp16"); idx.reserve(n)
g = torch.Generator(device=dev).manual_seed(1)
for _ in range(4):
    xb = torch.randn(n // 4, d, generator=g, device=dev); xb /= xb.norm(dim=1, keepdim=True); idx.add(xb.half())
out = []
for B in (64, 256, 1024):
    q = torch.randn(B, d, generator=g, device=dev)
    for _ in range(3): idx.search(q, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): idx.search(q, 10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    out.append(f"B={B}: {ms:.4f} ms {2.0*n*d*B/(ms*1e-3)/1e12:.0f} TF")
print("DEBUG", os.environ.get("PRS_UMMA_DEBUG", "0"), "NB", os.environ.get("PRS_UMMA_NB", "-"), "CL", os.environ.get("PRS_UMMA_CLUSTER", "-"), "d=%d" % d, " | ".join(out), flush=True)
'''
for dd in ("768", "384"):
    for nb in ("1", "2"):
        for dbg in os.environ["DBGS"].split():
            env = dict(os.environ, PRS_UMMA_DEBUG=dbg, PRS_UMMA_NB=nb, DD=dd)
            subprocess.run([sys.executable, "-c", code], env=env)
PY
