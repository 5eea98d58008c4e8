#!/bin/bash
mkdir -p gpurun_out
for d in 3 4 3 4; do
  FPSG_BENCH_DEPTH=$d timeout 300 python bench.py --steps 1000 --warmup 100 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('depth $d value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ag', round(d['e2e_autograd']['value']))" | tee -a gpurun_out/r03m_depth.txt
done
