#!/bin/bash
# inflate iteration on a thin budget: every test that runs the inflate kernel, then the bench line
timeout 300 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py tests/test_gpu_api_stream.py tests/test_gpu_scale.py tests/test_gpu_guard.py -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 --pipeline-streams 32768 2>gpurun_out/bench_i4.err | tee gpurun_out/bench_i4.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('corpusA value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],4), 'mixed decompress', round(d['mixed_pipeline']['decompress'],1), 'compress', round(d['mixed_pipeline']['compress'],2))"
