#!/bin/bash
# call 7: section H again (helpers on copies), A/B, then the inflate tests
mkdir -p gpurun_out
PRODUCERS=1 KINDS=text,binary,mixedB,corpusA timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 auto auto_nolanehdr 2>&1 | tee gpurun_out/inflate_modes_r3g.txt | tail -5
timeout 900 python -m pytest tests/test_gpu_inflate.py tests/test_gpu_fuzz.py tests/test_gpu_engines.py tests/test_gpu_configs.py tests/test_gpu_reuse.py tests/test_gpu_check_build.py tests/test_gpu_determinism.py -x -q 2>&1 | tail -4
