#!/bin/bash
for c in 8 12; do
  echo "== L1 ctas/SM $c"; BDF_L1_CTAS_PER_SM=$c timeout 200 python gpurun_scripts/gpu_compress.py 1 16384 2>&1 | grep "L1:" 
done
