#!/bin/bash
# large-batch experiments on the experiments build (make EXPERIMENTS=1 OBJDIR=build_x OUT=../libprs_x.so): tile width x cluster size x bound refresh
export PRS_LIB_PATH=$PWD/persian-rag-system_b200/libprs_x.so
run() { env "$@" python - <<'PY'
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
for B in (256, 512, 1024, 4096):
    q = torch.randn(B, d, generator=g, device=dev)
    for _ in range(3): idx.search(q, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8): idx.search(q, 10)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    out.append(f"B={B}: {ms:.3f} ms {2.0*n*d*B/(ms*1e-3)/1e12:.0f} TF")
print("NB", os.environ.get("PRS_UMMA_NB", "-"), "CL", os.environ.get("PRS_UMMA_CLUSTER", "-"), "REBOOT", os.environ.get("PRS_UMMA_REBOOT", "-"), "d=%d" % d, " | ".join(out), flush=True)
PY
}
for nb in 0 2; do for cl in 2 4; do run PRS_UMMA_NB=$nb PRS_UMMA_CLUSTER=$cl; done; done
run PRS_UMMA_NB=2 PRS_UMMA_CLUSTER=2 PRS_UMMA_REBOOT=0
run PRS_UMMA_NB=0 PRS_UMMA_CLUSTER=2 PRS_UMMA_REBOOT=0
run PRS_UMMA_NB=0 PRS_UMMA_CLUSTER=2 DD=384
