#!/bin/bash
# first run of the lane-per-stream inflate engine: parity in every mode, then throughput per corpus kind
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_engines.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_guard.py tests/test_gpu_reuse.py -x -q 2>&1 | tail -8
timeout 400 python -u gpurun_scripts/inflate_modes.py 16384 2>&1 | tee gpurun_out/inflate_modes_r2b.txt | tail -12
