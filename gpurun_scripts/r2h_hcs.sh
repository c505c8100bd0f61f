#!/bin/bash
# hcs kernel: parity, throughput, one ncu capture on text level 6
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_checksum_compress.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -5
timeout 600 python -u gpurun_scripts/deflate_probe.py 2,6 8192 2>&1 | tee gpurun_out/deflate_probe_$1.txt | tail -12
timeout 120 python -u gpurun_scripts/deflate_probe.py 6 1216 text > gpurun_out/plain_$1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deflate_hcs -s 1 -c 1 -f -o gpurun_out/prof_hcs_$1 python -u gpurun_scripts/deflate_probe.py 6 1216 text > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log | cut -c1-200
