#!/bin/bash
# deferred long-codeword dispatch in the lane kernel, the -DBDF_CHECK build, the determinism test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_engines.py tests/test_gpu_check_build.py tests/test_gpu_determinism.py tests/test_gpu_inflate.py -x -q 2>&1 | tail -8
PRODUCERS=1 KINDS=text,binary,mixedB timeout 400 python -u gpurun_scripts/inflate_modes.py 65536 lane0 lane2 auto 2>&1 | tee gpurun_out/inflate_modes_$1.txt | tail -4
