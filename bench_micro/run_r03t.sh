#!/bin/bash
# round 2: tensor-core kNN: 16 epilogue warps, re-scoring with 32-channel tiles
mkdir -p gpurun_out
for cfg in "2 64 1024 20" "1 72 700 20" "1 128 2048 20" "1 32 4000 32" "2 16 300 5" "32 64 1024 20" "32 128 1024 20"; do
  timeout 120 python bench_micro/knn_tc_check.py $cfg >> gpurun_out/r03t_knn_tc.jsonl 2>> gpurun_out/r03t_knn_tc.err; echo "cfg $cfg rc=$?" >> gpurun_out/r03t_knn_tc.err
done
cat gpurun_out/r03t_knn_tc.jsonl | cut -c1-400; tail -8 gpurun_out/r03t_knn_tc.err
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_knn_tc_gpu.py tests/test_property_gpu.py tests/test_reference_harness_gpu.py tests/test_graph_feature.py -x -q > gpurun_out/r03t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r03t_tests.log
tail -3 gpurun_out/r03t_tests.log
for cfg in "32 64 1024 20" "32 128 1024 20"; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r03t_launch_$(echo $cfg | tr ' ' '_').csv python bench_micro/knn_one.py $cfg > /dev/null 2>&1
done
python - <<'PY'
import csv, glob, collections
for f in sorted(glob.glob("gpurun_out/r03t_launch_*.csv")):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        d[r[kn][:60]].append(float(r[mv].replace(",", "")))
    print(f)
    for k_, v in d.items():
        v.sort(); print("  ", k_, "n=%d median=%.1f us" % (len(v), v[len(v)//2] / 1e3))
PY
