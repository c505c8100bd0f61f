#!/bin/bash
# call 10: pre-pass with the header bytes staged in shared memory
mkdir -p gpurun_out
timeout 300 python -u gpurun_scripts/prehdr_probe.py 65536 2>&1 | tee gpurun_out/prehdr_probe_r3j.txt | tail -4
timeout 900 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_engines.py tests/test_gpu_configs.py tests/test_gpu_reuse.py tests/test_gpu_check_build.py tests/test_gpu_determinism.py tests/test_gpu_guard.py tests/test_gpu_api_stream.py -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r3j.csv python gpurun_scripts/prehdr_probe.py 65536 > /dev/null 2>&1; grep -i "prehdr" gpurun_out/launches_r3j.csv | head -3 | cut -c1-200
