#!/bin/bash
# call 18: adaptive 8 / 9-bit litlen tables in global memory (BDF_LANE_CFG=5): throughput, then parity under that config
mkdir -p gpurun_out
PRODUCERS=1 KINDS=text,binary,mixedB,lowent,corpusA timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 lane0 lane3 lane5 auto auto5 2>&1 | tee gpurun_out/inflate_modes_r3q.txt | tail -6
BDF_LANE_CFG=5 timeout 900 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_engines.py tests/test_gpu_configs.py tests/test_gpu_reuse.py tests/test_gpu_guard.py tests/test_gpu_api_stream.py -x -q 2>&1 | tail -4
