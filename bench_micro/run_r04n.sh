#!/bin/bash
mkdir -p gpurun_out
export FPSG_TUNING_LIB=1
export CUDA_DEVICE_MAX_CONNECTIONS=32
for pct in 100 60 40 25 10 0; do
FPSG_SWEEP_TRIGGER_PCT=$pct timeout 200 python bench_micro/overlap_probe.py 32 2048 2048 >> gpurun_out/r04n_trigger.jsonl 2>>gpurun_out/r04n.err
done
cat gpurun_out/r04n_trigger.jsonl | cut -c1-420; tail -3 gpurun_out/r04n.err
