#!/bin/bash
# extra shapes for the record (profiles/): the reference's own mode (fp32 L2), other dims, wide k
out=gpurun_out/extra_shapes.jsonl; : > $out
run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline --capacity-rows 0 --inflight 1 "$@" 2>/dev/null | grep '^{' | python -c "
import sys,json
j=json.loads(sys.stdin.readline())
keep={k:j[k] for k in ('value','ms_per_step','kernel_path','roofline','e2e','step_breakdown_ms','config')}
keep['sweep']=j.get('sweep')
print(json.dumps(keep))" >> $out; }
run --storage fp32 --metric l2 --batch 1
run --storage fp32 --metric l2 --batch 8 --no-sweep
run --d 384
run --d 512
run --k 100 --no-sweep
run --k 1000 --no-sweep
run --metric l2 --no-sweep
run --storage bf16 --no-sweep
python - <<'PY'
import json
for l in open('gpurun_out/extra_shapes.jsonl'):
    j=json.loads(l); c=j['config']
    print(c['storage'], c['metric'], 'd',c['d'],'k',c['k'],'B',c['batch'], j['kernel_path'], 'qps',round(j['value']), 'ms',round(j['ms_per_step'],4), 'frac',round(j['roofline']['frac'],3))
PY
