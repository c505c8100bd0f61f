#!/bin/bash
# whole GPU suite, smoke, the bench line and the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_$1.err | tee gpurun_out/bench_$1.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('config2 value',round(d['value'],1),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1),'cpu',round(d.get('cpu_baseline',{}).get('value',0),1))
for k in ('compress_l1','compress_l6','compress_l12','mixed_pipeline'):
    s=d.get(k)
    if s: print(k,'value',round(s['value'],2),'ms',round(s.get('kernel_ms',0),2),'frac',round(s['roofline']['frac'],4),'traffic_x',round((s['roofline']['traffic'] or 0)/s['roofline']['algorithmic_bytes_per_launch'],2),'e2e',round(s['e2e']['value'],2),'cpu',round(s.get('cpu_baseline',{}).get('value',0),3), {x:round(s[x],1) for x in ('compress','decompress') if x in s})
"
tail -3 gpurun_out/bench_$1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$1.json; cut -c1-300 gpurun_out/bench_ref_$1.json
