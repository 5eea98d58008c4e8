#!/bin/bash
# round 2: cross-step overlap probe (independent batches round-robin over streams)
mkdir -p gpurun_out
for cfg in "32 2048 2048" "37 2048 2048" "256 2048 2048"; do
  timeout 300 python bench_micro/overlap_probe.py $cfg >> gpurun_out/r02m_overlap.jsonl 2>> gpurun_out/r02m_overlap.err
done
cat gpurun_out/r02m_overlap.jsonl; tail -5 gpurun_out/r02m_overlap.err
