#!/bin/bash
# call 6: lanes read the headers of their later blocks inside the decode loop (inflate_lane.cuh, section H)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
PRODUCERS=2 KINDS=text,binary,lowent,mixedB,corpusA timeout 600 python -u gpurun_scripts/inflate_modes.py 65536 auto auto_nolanehdr 2>&1 | tee gpurun_out/inflate_modes_r3f.txt | tail -12
