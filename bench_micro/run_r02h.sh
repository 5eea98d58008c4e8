#!/bin/bash
# round 2: finish kernel, cluster size / one-CTA-per-SM variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chamfer_fused_gpu.py tests/test_chamfer_multi_gpu.py tests/test_pipeline_host_gpu.py -x -q > gpurun_out/r02h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_tests.log
for s in "32 2048 2048" "37 2048 2048" "5 2048 2048" "256 2048 2048" "64 4096 4096" "32 1024 1024"; do timeout 200 python bench_micro/step_c1.py $s >> gpurun_out/r02h_step.jsonl 2>&1; done
for v in "4 1" "4 0" "8 1" "8 0"; do set -- $v
  FPSG_TUNING_LIB=1 FPSG_FINISH_CL=$1 FPSG_FINISH_PAD=$2 timeout 200 python bench_micro/step_c1.py 32 2048 2048 >> gpurun_out/r02h_variants.jsonl 2>&1
  FPSG_TUNING_LIB=1 FPSG_FINISH_CL=$1 FPSG_FINISH_PAD=$2 timeout 200 python bench_micro/step_c1.py 37 2048 2048 >> gpurun_out/r02h_variants.jsonl 2>&1
done
timeout 200 python bench_micro/timeline.py > gpurun_out/r02h_timeline.json 2>&1
tail -3 gpurun_out/r02h_tests.log; cat gpurun_out/r02h_step.jsonl gpurun_out/r02h_variants.jsonl; cut -c1-1500 gpurun_out/r02h_timeline.json
