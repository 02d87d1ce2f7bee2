#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 100 --warmup 10 2>gpurun_out/b.err > gpurun_out/bench_r1d.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r1d.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d.get('ops',{}).items(): print(k, v if not isinstance(v,dict) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','frac','mpix_s','gbs')})
print(d.get('cpu_baseline'))"
