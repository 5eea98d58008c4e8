#!/bin/bash
# round 2: tensor-core kNN, re-scoring with coalesced staging; per-kernel times; full ncu capture of the three kernels
mkdir -p gpurun_out
for cfg in "2 64 1024 20" "1 72 700 20" "1 64 2048 20" "1 32 4100 32" "32 64 1024 20" "32 128 1024 20"; do
  timeout 120 python bench_micro/knn_tc_check.py $cfg >> gpurun_out/r02s_knn_tc.jsonl 2>> gpurun_out/r02s_knn_tc.err; echo "cfg $cfg rc=$?" >> gpurun_out/r02s_knn_tc.err
done
cat gpurun_out/r02s_knn_tc.jsonl; tail -6 gpurun_out/r02s_knn_tc.err
timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_property_gpu.py -x -q > gpurun_out/r02s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_tests.log
tail -4 gpurun_out/r02s_tests.log
for cfg in "32 64 1024 20" "32 128 1024 20"; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_launch_$(echo $cfg | tr ' ' '_').csv python bench_micro/knn_one.py $cfg > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn_tc' -s 6 -c 3 -o gpurun_out/r02s_knn_tc_c64 python bench_micro/knn_one.py 32 64 1024 20 > gpurun_out/r02s_ncu_full.log 2>&1
python - <<'PY'
import csv, glob, collections
for f in sorted(glob.glob("gpurun_out/r02s_launch_*.csv")):
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
