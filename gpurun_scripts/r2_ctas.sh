#!/bin/bash
# resident-CTA sweep for the level-1 and hash-chain compress kernels (GB/s of uncompressed input)
for c in 8 12 16; do
  echo "== L1 ctas/SM $c"; BDF_L1_CTAS_PER_SM=$c timeout 200 python gpurun_scripts/gpu_compress.py 1 16384 2>&1 | grep "L1:" 
done
for c in 8 12 16; do
  echo "== HC ctas/SM $c"; BDF_HC_CTAS_PER_SM=$c timeout 300 python gpurun_scripts/gpu_compress.py 6 8192 2>&1 | grep "L6:"
done
