#!/bin/bash
# call 20: lane table placement by batch size: pipeline number, inflate / determinism / resume tests, profiles of the changed kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_inflate.py tests/test_gpu_engines.py tests/test_gpu_configs.py tests/test_gpu_resume.py -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --compress-levels '' 2>gpurun_out/bench_r3t.err | tee gpurun_out/bench_r3t.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); mp=d['mixed_pipeline']
print('value',round(d['value'],1),'pipeline',round(mp['value'],2),'compress',round(mp['compress'],2),'decompress',round(mp['decompress'],2))"
bash gpurun_scripts/r3s_profiles.sh
