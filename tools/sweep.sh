#!/bin/bash
# batch sweep only (device resident), prints batch / ms / qps / TFLOP/s
python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l)
        print('headline', round(j['value']), j['ms_per_step'], j['roofline']['frac'])
        for s in j['sweep']: print(s['batch'], s['ms'], s['qps'], s['tflops'], s['path'])
    elif 'rror' in l: print(l[:300])
"
