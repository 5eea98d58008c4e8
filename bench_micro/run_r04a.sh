#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench_micro/e2e_depth.py > gpurun_out/r04a_e2e_depth.json 2> gpurun_out/r04a.err; cat gpurun_out/r04a_e2e_depth.json; tail -3 gpurun_out/r04a.err
for d in 4 6 8 4 6 8; do
  FPSG_BENCH_DEPTH=$d timeout 300 python bench.py --steps 1000 --warmup 100 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('depth $d value', round(d['value']), 'e2e', round(d['e2e']['value']))" | tee -a gpurun_out/r04a_depth.txt
done
