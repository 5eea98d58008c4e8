#!/bin/bash
# round 2: full GPU test suite + bench (1 GPU) with lanes and the tensor-core kNN
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r03d_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r03d_gpu_tests.log
tail -5 gpurun_out/r03d_gpu_tests.log
timeout 900 python bench.py --steps 1000 --warmup 100 > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err
tail -3 gpurun_out/r03d_bench.err; cut -c1-400 gpurun_out/r03d_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03d_bench_reference.json 2>> gpurun_out/r03d_bench.err
cut -c1-300 gpurun_out/r03d_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r03d_smoke.log 2>&1; tail -2 gpurun_out/r03d_smoke.log
