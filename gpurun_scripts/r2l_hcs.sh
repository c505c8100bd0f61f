#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
timeout 600 python -u gpurun_scripts/deflate_probe.py 2,6 8192 text,mixedB,binary 2>&1 | tee gpurun_out/deflate_probe_$1.txt | tail -8
