#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_chamfer_fused_gpu.py tests/test_chamfer_multi_gpu.py tests/test_chamfer_gpu.py tests/test_knn_gpu.py tests/test_knn_tc_gpu.py tests/test_property_gpu.py tests/test_integration_gpu.py tests/test_chamfer_lanes_gpu.py -x -q > gpurun_out/r03r_tests.log 2>&1; tail -3 gpurun_out/r03r_tests.log
timeout 900 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/r03r_bench.json 2> gpurun_out/r03r_bench.err; tail -2 gpurun_out/r03r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03r_bench.json'))
print({k: d.get(k) for k in ("value","ms_per_step","value_single_stream","latency_us_per_step")}, "e2e", d["e2e"]["value"])
print({k:(round(v["us_per_call"],1) if isinstance(v,dict) and "us_per_call" in v else None) for k,v in d["secondary"].items()})
PY
