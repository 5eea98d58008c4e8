#!/bin/bash
# round 2: finish kernel polling per-pair sweep progress; parity, step times, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chamfer_fused_gpu.py tests/test_chamfer_gpu.py tests/test_chamfer_multi_gpu.py tests/test_pipeline_host_gpu.py -x -q > gpurun_out/r02g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_tests.log
for s in "32 2048 2048" "37 2048 2048" "5 2048 2048" "256 2048 2048" "64 4096 4096"; do timeout 200 python bench_micro/step_c1.py $s >> gpurun_out/r02g_step.jsonl 2>&1; done
timeout 200 python bench_micro/timeline.py > gpurun_out/r02g_timeline.json 2>&1
timeout 200 python bench_micro/timeline.py 256 2048 2048 > gpurun_out/r02g_timeline_b256.json 2>&1
tail -3 gpurun_out/r02g_tests.log; cat gpurun_out/r02g_step.jsonl; cut -c1-3000 gpurun_out/r02g_timeline.json
