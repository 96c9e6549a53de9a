#!/bin/bash
B="python bench.py --steps 10 --warmup 3 --no-sweep --no-cpu-baseline --capacity-rows 0 --inflight 1"
run() { echo "== $*"; env $1 $2 $B $3 $4 $5 $6 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print('scan_ms',round(r['avg_launch_ms'],4),'launches',r['launches_timed'],'step_ms',round(j['ms_per_step'],4))
    elif 'rror' in l: print(l[:200])
"; }
run PRS_UMMA_DEBUG=2 PRS_UMMA_CLUSTER=2 --batch 256
run PRS_UMMA_DEBUG=130 PRS_UMMA_CLUSTER=2 --batch 256
run PRS_UMMA_DEBUG=2 PRS_UMMA_CLUSTER=2 --batch 256 --d 384
run PRS_UMMA_DEBUG=130 PRS_UMMA_CLUSTER=2 --batch 256 --d 384
