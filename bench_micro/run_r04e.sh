#!/bin/bash
mkdir -p gpurun_out
export FPSG_TUNING_LIB=1
timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04e_pdl.jsonl 2>>gpurun_out/r04e.err
FPSG_FIN_NOPDL=1 timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04e_pdl.jsonl 2>>gpurun_out/r04e.err
FPSG_SWEEP_NOPDL=1 timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04e_pdl.jsonl 2>>gpurun_out/r04e.err
FPSG_FIN_NOPDL=1 FPSG_SWEEP_NOPDL=1 timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04e_pdl.jsonl 2>>gpurun_out/r04e.err
cat gpurun_out/r04e_pdl.jsonl; tail -3 gpurun_out/r04e.err
