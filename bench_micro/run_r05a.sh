#!/bin/bash
# round 2, final build: bench line (1 GPU), ncu launch list of the bench command, ncu full capture of the two step kernels, smoke()
mkdir -p gpurun_out
python bench.py > gpurun_out/r05a_bench_full.json 2> gpurun_out/r05a_bench_full.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r05a_bench_full.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['value_single_stream']), round(d['e2e']['value']), d['roofline']['frac'], d['roofline_step']['frac'], d['clocks'])
for k,v in d['secondary'].items(): print(k, json.dumps(v)[:200])
"
CMD="python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline"
$CMD > gpurun_out/r05a_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 600 --csv --log-file gpurun_out/r05a_launches_bench.csv $CMD > gpurun_out/r05a_ncu_launch.log 2>&1
python bench_micro/launch_summary.py gpurun_out/r05a_launches_bench.csv > gpurun_out/r05a_launches_summary.txt 2>&1; head -12 gpurun_out/r05a_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:'chamfer_sweep|chamfer_finish' -s 60 -c 4 -f -o gpurun_out/r05a_step_kernels $CMD > gpurun_out/r05a_ncu_full.log 2>&1; tail -1 gpurun_out/r05a_ncu_full.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
