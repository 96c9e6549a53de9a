#!/bin/bash
# the driver's scaling series on one 8-GPU box: N = 1, 2, 4, 8 back to back, default flags
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --no-sweep --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "N=1 rc=$?"
for N in 2 4 8; do
  timeout 900 $TR --nproc-per-node $N bench.py --gpus $N > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "N=$N rc=$?"
done
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        j=json.loads([l for l in open(f'gpurun_out/scale_n{N}.json') if l.startswith('{')][-1])
        c=j.get('capacity_scaling',{})
        print(N, 'qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']), 'pipelined', round(j.get('pipelined',{}).get('value',0)),
              '| capacity ms', round(c.get('ms_per_batch',0),3), 'agg GB/s', round(c.get('aggregate_scan_gbs',0)), 'frac', round(c.get('frac_hbm',0),3), c.get('error',''))
    except Exception as e:
        print(N, 'failed', e)
PY
