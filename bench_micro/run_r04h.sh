#!/bin/bash
mkdir -p gpurun_out
export FPSG_TUNING_LIB=1
timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04h_pdl.jsonl 2>>gpurun_out/r04h.err
FPSG_FIN_NOPDL=1 timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04h_pdl.jsonl 2>>gpurun_out/r04h.err
FPSG_SWEEP_NOPDL=1 timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04h_pdl.jsonl 2>>gpurun_out/r04h.err
cat gpurun_out/r04h_pdl.jsonl | cut -c1-330; tail -3 gpurun_out/r04h.err
