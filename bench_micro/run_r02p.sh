#!/bin/bash
# round 2: first runs of the tensor-core kNN pre-filter
mkdir -p gpurun_out
for cfg in "2 64 1024 20" "1 128 520 20" "2 64 300 20" "32 64 1024 20" "32 128 1024 20"; do
  timeout 120 python bench_micro/knn_tc_check.py $cfg >> gpurun_out/r02p_knn_tc.jsonl 2>> gpurun_out/r02p_knn_tc.err; echo "cfg $cfg rc=$?" >> gpurun_out/r02p_knn_tc.err
done
cat gpurun_out/r02p_knn_tc.jsonl; tail -20 gpurun_out/r02p_knn_tc.err
