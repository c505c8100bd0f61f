#!/bin/bash
# quick inflate iteration: parity net, then the headline at G=16
timeout 400 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --pipeline-streams 8192 2>gpurun_out/bench_i3.err | tee gpurun_out/bench_i3.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('corpusA value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'mixed decompress', round(d['mixed_pipeline']['decompress'],1))"
