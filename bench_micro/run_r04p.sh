#!/bin/bash
# e2e of the pipeline with the sweep's completion trigger at 100 / 60 / 25 % of the run left (tuning build)
mkdir -p gpurun_out
for rep in 1 2; do for pct in 100 60 25; do
  FPSG_TUNING_LIB=1 FPSG_SWEEP_TRIGGER_PCT=$pct python bench_micro/e2e_depth.py | sed "s/^/{\"pct\": $pct, \"res\": /; s/$/}/" >> gpurun_out/r04p_e2e_trigger.jsonl
done; done
cat gpurun_out/r04p_e2e_trigger.jsonl
