#!/bin/bash
mkdir -p gpurun_out
export FPSG_TUNING_LIB=1
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04i_pdl.jsonl 2>>gpurun_out/r04i.err
FPSG_FIN_NOPDL=1 timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04i_pdl.jsonl 2>>gpurun_out/r04i.err
FPSG_FIN_NOPDL=1 FPSG_SWEEP_NOPDL=1 timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04i_pdl.jsonl 2>>gpurun_out/r04i.err
cat gpurun_out/r04i_pdl.jsonl | cut -c1-420; tail -3 gpurun_out/r04i.err
