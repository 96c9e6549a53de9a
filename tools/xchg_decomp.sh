#!/bin/bash
# row-sharded step decomposition on the experiments build (N ranks on one box): what does the merge + exchange kernel wait for?
# PRS_XCHG_DBG bits: 1 do not wait for the peers' flags, 2 do not store into the peers (results WRONG, timing only)
N=${1:-2}
export PRS_LIB_PATH=$PWD/persian-rag-system_b200/libprs_x.so
for dbg in 0 1 2 3; do
  PRS_XCHG_DBG=$dbg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 \
      --no-sweep --no-cpu-baseline --extras none --capacity-rows 0 > gpurun_out/xd_$dbg.json 2> gpurun_out/xd_$dbg.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/xd_$dbg.json") if l.startswith("{")][-1])
    print("N=$N PRS_XCHG_DBG=$dbg", "ms/step", round(d["ms_per_step"], 4), {k[:5]: round(v, 4) for k, v in d["step_breakdown_ms"].items()})
except Exception as e:
    print("N=$N PRS_XCHG_DBG=$dbg failed", e)
PY
done
