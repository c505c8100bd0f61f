#!/bin/bash
# call 23: large randomised differential batches through the large-batch decompress configuration
mkdir -p gpurun_out
timeout 600 python -u gpurun_scripts/stress_inflate.py 40000 1 2>&1 | tee gpurun_out/stress_r3w.txt | tail -8
BDF_LIBRARY=$PWD/libdeflate_rsx_b200/libbdeflate_check.so timeout 600 python -u gpurun_scripts/stress_inflate.py 36000 2 2>&1 | tee -a gpurun_out/stress_r3w.txt | tail -6
