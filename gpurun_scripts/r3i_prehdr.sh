#!/bin/bash
# call 9: L1 / shared-memory split of the pre-pass kernel
mkdir -p gpurun_out
timeout 300 python -u gpurun_scripts/prehdr_probe.py 65536 2>&1 | tee gpurun_out/prehdr_probe_r3i.txt | tail -8
