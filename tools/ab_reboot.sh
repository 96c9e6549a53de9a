#!/bin/bash
# A/B on ONE box: cross-CTA bound refresh off / on (experiments build), headline shape and large batches
export PRS_LIB_PATH=$PWD/persian-rag-system_b200/libprs_x.so
for rb in 0 1 0 1; do
  PRS_UMMA_REBOOT=$rb python bench.py --steps 30 --warmup 5 --capacity-rows 0 --extras none --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read())
print('reboot=$rb', 'B64 step', round(j['ms_per_step'],4), 'kernel', round(j['roofline']['avg_launch_ms'],4), [(r['batch'], r['ms']) for r in j['sweep'] if r['batch'] in (1,16,128,256,1024)])"
done
