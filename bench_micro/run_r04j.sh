#!/bin/bash
# round 2 (final): full GPU suite, bench on 1 GPU (4 lanes), smoke
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r04j_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r04j_gpu_tests.log
tail -4 gpurun_out/r04j_gpu_tests.log
timeout 900 python bench.py --steps 1000 --warmup 100 > gpurun_out/r04j_bench.json 2> gpurun_out/r04j_bench.err
tail -3 gpurun_out/r04j_bench.err; cut -c1-300 gpurun_out/r04j_bench.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r04j_smoke.log 2>&1; tail -2 gpurun_out/r04j_smoke.log
