#!/bin/bash
# call 11: staged pre-pass, loads in flight
mkdir -p gpurun_out
timeout 300 python -u gpurun_scripts/prehdr_probe.py 65536 2>&1 | tee gpurun_out/prehdr_probe_r3k.txt | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r3k.csv python gpurun_scripts/prehdr_probe.py 65536 > /dev/null 2>&1; grep -i "prehdr" gpurun_out/launches_r3k.csv | tail -3 | cut -d, -f5,15-
