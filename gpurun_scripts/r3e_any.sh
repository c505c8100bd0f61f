#!/bin/bash
# call 5: batch-step parity of the resumable decoder, any-length device compress
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resume.py tests/test_gpu_device_any.py -x -q 2>&1 | tail -15
