#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 1000 --warmup 100 > gpurun_out/r03n_bench.json 2> gpurun_out/r03n_bench.err
tail -3 gpurun_out/r03n_bench.err; cut -c1-200 gpurun_out/r03n_bench.json
timeout 600 python -m pytest tests/test_pipeline_host_gpu.py tests/test_chamfer_lanes_gpu.py -x -q 2>&1 | tail -2
