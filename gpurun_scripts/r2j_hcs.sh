#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py tests/test_gpu_size.py -x -q 2>&1 | tail -5
timeout 600 python -u gpurun_scripts/deflate_probe.py 2,6 8192 2>&1 | tee gpurun_out/deflate_probe_$1.txt | tail -16
echo "== all streams through the new kernel"
BDF_HC_KERNEL=new timeout 600 python -u gpurun_scripts/deflate_probe.py 2,6 8192 text,corpusA 2>&1 | tail -5
