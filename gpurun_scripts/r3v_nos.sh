#!/bin/bash
# call 22: cost kernel with the literal-only fast path
mkdir -p gpurun_out
LEVELS=10,12 timeout 600 python -u gpurun_scripts/nos_probe.py 2048 2>&1 | tee gpurun_out/nos_probe_r3v.txt | tail -10
timeout 600 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_checksum_compress.py -x -q 2>&1 | tail -3
