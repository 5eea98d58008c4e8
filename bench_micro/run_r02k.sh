#!/bin/bash
# round 2: two GPUs -- sharded step with the in-kernel all-reduce, bench at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q > gpurun_out/r02k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_tests.log
tail -15 gpurun_out/r02k_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 100 > gpurun_out/r02k_bench_2gpu.json 2> gpurun_out/r02k_bench_2gpu.err
tail -8 gpurun_out/r02k_bench_2gpu.err; cut -c1-1500 gpurun_out/r02k_bench_2gpu.json
