#!/bin/bash
# ncu --set full of the gzip inflate instance on the mixed corpus (second launch of the pipeline section)
timeout 250 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:inflate_kernel<\(int\)2' -s 1 -c 1 -f -o gpurun_out/prof_inflate_mixed_r3 \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --pipeline-streams 16384 > gpurun_out/prof_inflate_mixed_r3.log 2>&1
tail -3 gpurun_out/prof_inflate_mixed_r3.log | cut -c1-300
ls -la gpurun_out/prof_inflate_mixed_r3.ncu-rep
