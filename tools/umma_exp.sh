#!/bin/bash
# kernel-variant experiments for the tcgen05 scan (timing only; results are wrong in debug modes)
B="python bench.py --steps 20 --warmup 3 --no-sweep --no-cpu-baseline"
for cfg in "0 0" "8 0" "1 0"; do
  set -- $cfg
  echo "== dbg=$1 kbs=$2"
  PRS_UMMA_DEBUG=$1 PRS_UMMA_KBS=$2 $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); r=j['roofline']; print('scan_ms',round(r['avg_launch_ms'],4),'GB/s',round(r['achieved']),'frac',round(r['frac'],3),'step_ms',round(j['ms_per_step'],4))
"
done
