#!/bin/bash
# inflate kernel at each lane-group size: parity tests, then bench (corpus A) and text throughput
for g in 8 16 32; do
  export BDF_INFLATE_GROUP=$g
  echo "== G=$g"
  timeout 300 python -m pytest tests/test_gpu_inflate.py -x -q 2>&1 | tail -2
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('corpusA value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1))"
  timeout 300 python gpurun_scripts/gpu_quick.py 2>&1 | tail -1
done
