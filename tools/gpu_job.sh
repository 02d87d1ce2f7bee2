#!/bin/bash
# Scratch driver for one gpurun call (edited per experiment; the durable scripts are tools/run_op.py and bench.py).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_r1.json 2>gpurun_out/bench_ref.err; cut -c1-160 gpurun_out/bench_ref_r1.json
python bench.py --steps 100 --warmup 10 2>gpurun_out/b.err > gpurun_out/bench_r1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r1.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_of_traffic'), d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
for k,v in d.get('ops',{}).items(): print(k, v if not isinstance(v,dict) else {a:b for a,b in v.items() if a in ('ms','frac','mpix_s')})
print(d.get('cpu_baseline'))"
