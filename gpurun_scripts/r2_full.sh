#!/bin/bash
# full-shape config tests, then the default bench line (with cpu baseline + pipeline section)
TAG=${1:-r2h}
timeout 700 python -m pytest tests/test_gpu_configs.py -x -q 2>&1 | tail -15
timeout 400 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG}.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", d.get("cpu_baseline",{}).get("value"))
print(d.get("mixed_pipeline"))
PY
