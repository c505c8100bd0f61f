#!/bin/bash
# call 4: resumable inflate (new tests), C++ mirror, then the whole suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_resume.py tests/test_gpu_cpp_host.py tests/test_gpu_api_stream.py -x -q 2>&1 | tail -15
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
PRODUCERS=1 KINDS=corpusA,mixedB timeout 300 python -u gpurun_scripts/inflate_modes.py 65536 auto concurrent 2>&1 | tee gpurun_out/inflate_modes_r3d.txt | tail -3
