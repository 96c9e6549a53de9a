#!/bin/bash
# final-code scaling series on one 8-GPU box: the multi-GPU pytest, then the headline at N = 1, 2, 4, 8 (N = 8 with the
# 400M x 384 capacity measurement at B = 64; the other extras of the 8-GPU run are in r2_scaling_series.json)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_multigpu.py -m gpu -q -x 2>&1 | tail -3
FLAGS="--extras none --no-sweep --no-cpu-baseline"
timeout 600 python bench.py --gpus 1 $FLAGS --capacity-rows 0 > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err; echo "N=1 rc=$?"
for N in 2 4; do
  timeout 900 $TR --nproc-per-node $N bench.py --gpus $N $FLAGS --capacity-rows 0 > gpurun_out/scale2_n$N.json 2> gpurun_out/scale2_n$N.err; echo "N=$N rc=$?"
done
timeout 1200 $TR --nproc-per-node 8 bench.py --gpus 8 $FLAGS > gpurun_out/scale2_n8.json 2> gpurun_out/scale2_n8.err; echo "N=8 rc=$?"
python - <<'PY'
import json
for N in (1,2,4,8):
    try:
        j=json.loads([l for l in open(f'gpurun_out/scale2_n{N}.json') if l.startswith('{')][-1])
        c=j.get('capacity_scaling') or {}
        print(N, 'qps', round(j['value']), 'ms', round(j['ms_per_step'],4), 'e2e', round(j['e2e']['value']), 'pipelined', round((j.get('pipelined') or {}).get('value',0)),
              'breakdown', {k[:5]: round(v,4) for k,v in j['step_breakdown_ms'].items()},
              '| capacity ms', round(c.get('ms_per_batch',0),3), 'agg GB/s', round(c.get('aggregate_scan_gbs',0)), 'frac', round(c.get('frac_hbm',0),3), c.get('error',''))
    except Exception as e:
        print(N, 'failed', e)
PY
tail -3 gpurun_out/scale2_n8.err
