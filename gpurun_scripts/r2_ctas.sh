#!/bin/bash
# resident-CTA sweep for the level-1 compress kernel (table slabs vs the 126 MB L2)
for c in 2 3 4 6; do
  echo "== L1 ctas/SM $c"; BDF_L1_CTAS_PER_SM=$c timeout 200 python gpurun_scripts/gpu_compress.py 1 16384 2>&1 | grep "L1:" 
done
for c in 2 4; do
  echo "== HC ctas/SM $c"; BDF_HC_CTAS_PER_SM=$c timeout 300 python gpurun_scripts/gpu_compress.py 6 8192 2>&1 | grep "L6:"
done
