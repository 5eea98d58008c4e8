#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python bench.py --steps 1000 --warmup 100 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), [round(v) for v in d['e2e']['passes_pairs_per_s']])" | tee -a gpurun_out/r04c_e2e.txt
done
timeout 900 python bench.py --steps 1000 --warmup 100 > gpurun_out/r04c_bench.json 2>gpurun_out/r04c_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r04c_bench.json')); print('full: value', round(d['value']), 'e2e', round(d['e2e']['value']), [round(v) for v in d['e2e']['passes_pairs_per_s']])"
