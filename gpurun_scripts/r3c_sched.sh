#!/bin/bash
# call 3: how the two inflate engines share the GPU (concurrent / one after the other / stream priority)
mkdir -p gpurun_out
PRODUCERS=1 KINDS=binary,mixedB,text timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 lane0 auto serial1 serial2 prio auto_nopre 2>&1 | tee gpurun_out/inflate_modes_r3c.txt | tail -4
PRODUCERS=1 KINDS=binary,mixedB timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 prio serial1 auto lane0 2>&1 | tee -a gpurun_out/inflate_modes_r3c.txt | tail -3
