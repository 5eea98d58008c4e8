#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench_micro/e2e_depth.py > gpurun_out/r03j_e2e_depth.json 2> gpurun_out/r03j.err; cat gpurun_out/r03j_e2e_depth.json; tail -3 gpurun_out/r03j.err
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_pipeline_host_gpu.py tests/test_graph_feature.py -x -q > gpurun_out/r03j_tests.log 2>&1; tail -3 gpurun_out/r03j_tests.log
