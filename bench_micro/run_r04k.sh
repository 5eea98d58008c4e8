#!/bin/bash
# round 2: two GPUs -- sharded tests and bench with lanes (lag 3 exchange)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py tests/test_pipeline_host_gpu.py -x -q > gpurun_out/r04k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r04k_tests.log
tail -4 gpurun_out/r04k_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 100 > gpurun_out/r04k_bench_2gpu.json 2> gpurun_out/r04k_bench_2gpu.err
tail -5 gpurun_out/r04k_bench_2gpu.err; cut -c1-600 gpurun_out/r04k_bench_2gpu.json
