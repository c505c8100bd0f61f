#!/bin/bash
# call 25: larger block buffer for the bulk-store replay (2.5 KB instead of 1.9 KB per lane group)
mkdir -p gpurun_out
timeout 300 python -u gpurun_scripts/prehdr_probe.py 65536 2>&1 | tee gpurun_out/prehdr_probe_r3y.txt | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_r3y.csv python gpurun_scripts/prehdr_probe.py 65536 > /dev/null 2>&1; grep "inflate_kernel" gpurun_out/launches_r3y.csv | tail -2 | cut -d, -f5,8,9,15-
timeout 900 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_engines.py tests/test_gpu_configs.py tests/test_gpu_guard.py tests/test_gpu_check_build.py -x -q 2>&1 | tail -3
