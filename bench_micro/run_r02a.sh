#!/bin/bash
# round 2, first GPU pass: fused finish kernel parity, full GPU suite, step anatomy, sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
timeout 600 python -m pytest tests/test_chamfer_fused_gpu.py -x -q > gpurun_out/r02a_fused_tests.log 2>&1; echo "fused rc=$?" >> gpurun_out/r02a_fused_tests.log
timeout 200 python bench_micro/step_c1.py > gpurun_out/r02a_step_c1.jsonl 2>&1
timeout 200 python bench_micro/step_c1.py 37 2048 2048 >> gpurun_out/r02a_step_c1.jsonl 2>&1
timeout 200 python bench_micro/step_c1.py 5 2048 2048 >> gpurun_out/r02a_step_c1.jsonl 2>&1
timeout 200 python bench_micro/step_c1.py 256 2048 2048 >> gpurun_out/r02a_step_c1.jsonl 2>&1
timeout 200 python bench_micro/step_c1.py 64 4096 4096 >> gpurun_out/r02a_step_c1.jsonl 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02a_gpu_tests.log 2>&1; echo "all rc=$?" >> gpurun_out/r02a_gpu_tests.log
timeout 900 compute-sanitizer --tool memcheck python bench_micro/sanitizer_workload.py > gpurun_out/r02a_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02a_memcheck.log
tail -5 gpurun_out/r02a_fused_tests.log; cat gpurun_out/r02a_step_c1.jsonl; tail -5 gpurun_out/r02a_gpu_tests.log; tail -5 gpurun_out/r02a_memcheck.log
