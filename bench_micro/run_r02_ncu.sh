#!/bin/bash
# round 2: ncu launch list of the bench command and full captures of the two kernels of the step
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launch.log 2>&1
$CMD > gpurun_out/r02_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'chamfer_(sweep|finish)' -s 20 -c 4 -o gpurun_out/r02_step_kernels $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -3 gpurun_out/r02_ncu_launch.log gpurun_out/r02_ncu_full.log; ls -la gpurun_out | tail -8
