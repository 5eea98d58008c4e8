#!/bin/bash
# round 2 (final): lanes test, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chamfer_lanes_gpu.py -x -q > gpurun_out/r03h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r03h_tests.log; tail -3 gpurun_out/r03h_tests.log
CMD="python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline"
$CMD > gpurun_out/r03h_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 600 --csv --log-file gpurun_out/r03h_launches_bench.csv $CMD > gpurun_out/r03h_ncu_launch.log 2>&1
tail -2 gpurun_out/r03h_ncu_launch.log | cut -c1-300
python bench_micro/launch_summary.py gpurun_out/r03h_launches_bench.csv > gpurun_out/r03h_launches_summary.txt 2>&1; cat gpurun_out/r03h_launches_summary.txt | head -20
