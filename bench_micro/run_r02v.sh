#!/bin/bash
# round 2: tensor-core kNN with BF16 operands, one stage per candidate block
mkdir -p gpurun_out
for cfg in "2 64 1024 20" "1 72 700 20" "1 128 2048 20" "1 32 4100 32" "2 16 300 5" "32 64 1024 20" "32 128 1024 20"; do
  timeout 120 python bench_micro/knn_tc_check.py $cfg >> gpurun_out/r02v_knn_tc.jsonl 2>> gpurun_out/r02v_knn_tc.err; echo "cfg $cfg rc=$?" >> gpurun_out/r02v_knn_tc.err
done
cat gpurun_out/r02v_knn_tc.jsonl; tail -8 gpurun_out/r02v_knn_tc.err
timeout 120 python bench_micro/knn_tc_ablate.py 32 64 1024 20 > gpurun_out/r02v_ablate.jsonl 2>>gpurun_out/r02v_knn_tc.err; head -1 gpurun_out/r02v_ablate.jsonl
timeout 120 python bench_micro/knn_tc_ablate.py 32 128 1024 20 > gpurun_out/r02v_ablate128.jsonl 2>>gpurun_out/r02v_knn_tc.err; head -1 gpurun_out/r02v_ablate128.jsonl
