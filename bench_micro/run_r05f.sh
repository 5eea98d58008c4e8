#!/bin/bash
# round 2, final build: the full GPU suite, the bench line (1 GPU), smoke()
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r05f_gpu_tests.log 2>&1; tail -2 gpurun_out/r05f_gpu_tests.log
python bench.py > gpurun_out/r05f_bench_full.json 2> gpurun_out/r05f_bench_full.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r05f_bench_full.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['value_single_stream']), round(d['e2e']['value']), d['e2e']['h2d_copies_alone']['gb_per_s_per_gpu'], d['roofline']['frac'], d['roofline_step']['frac'], d['clocks'])
for k,v in d['secondary'].items():
    if 'edge' in k or 'C3' in k: print(k, json.dumps(v)[:260])
"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
