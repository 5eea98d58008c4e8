#!/bin/bash
# round 2: ncu full capture of the four kNN tensor-core kernels (final build), C=64
mkdir -p gpurun_out
timeout 120 python bench_micro/knn_one.py 32 64 1024 20 > /dev/null 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn_tc' -s 8 -c 4 -o gpurun_out/r03f_knn_tc_c64 python bench_micro/knn_one.py 32 64 1024 20 > gpurun_out/r03f_ncu_full.log 2>&1
tail -3 gpurun_out/r03f_ncu_full.log
