#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'chamfer_finish' -s 30 -c 2 -o gpurun_out/r04l_finish $CMD > gpurun_out/r04l_ncu.log 2>&1
tail -2 gpurun_out/r04l_ncu.log
