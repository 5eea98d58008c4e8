#!/bin/bash
# round 2: lanes (per-slot streams in the pipeline, bench value over 3 lanes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_host_gpu.py tests/test_sharded_gpu.py tests/test_bench_contract.py -x -q > gpurun_out/r02n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02n_tests.log
tail -8 gpurun_out/r02n_tests.log
timeout 900 python bench.py --steps 1000 --warmup 100 > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
tail -5 gpurun_out/r02n_bench.err; cut -c1-1200 gpurun_out/r02n_bench.json
