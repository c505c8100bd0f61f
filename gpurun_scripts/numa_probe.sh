#!/bin/bash
# 8-GPU end-to-end probe: where the GPUs sit, and the bench line with / without NUMA-local pinned slabs
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/topo.txt
for d in /sys/bus/pci/devices/*; do
  if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-6)" = "0x0302" ]; then
    echo "$(basename $d) numa_node=$(cat $d/numa_node)" >> gpurun_out/topo.txt
  fi
done
lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/topo.txt
for mode in 0 1; do
  BDF_HOST_ALLOC_NUMA=$mode timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$mode \
    bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 4 --no-cpu-baseline --pipeline-streams 0 > gpurun_out/bench_n8_numa$mode.json 2> gpurun_out/bench_n8_numa$mode.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_n8_numa$mode.json').read().strip().splitlines()[-1])
print('numa mode $mode: value', round(d['value']), 'e2e', d['e2e']['value'])"
done
tail -3 gpurun_out/bench_n8_numa1.err
