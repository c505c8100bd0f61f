#!/bin/bash
# call 14: where the time of the three-kernel near-optimal tier goes (launch list), wave size sweep
mkdir -p gpurun_out
KINDS=corpusA,text LEVELS=12 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_nos_r3m.csv python gpurun_scripts/nos_probe.py 2048 > /dev/null 2>&1
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/launches_nos_r3m.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for r in csv.DictReader(lines):
    try: agg.setdefault(r['Kernel Name'][:50],[]).append(round(float(r['Metric Value'])/1e6,2))
    except: pass
for k,v in agg.items():
    if 'nos' in k: print(k, v[:16])
PY
for w in 4 14; do BDF_NOS_WAVE=$w KINDS=corpusA,text LEVELS=12 timeout 300 python -u gpurun_scripts/nos_probe.py 2048 2>&1 | sed "s/^/wave$w /" | tee -a gpurun_out/nos_probe_r3m_waves.txt | tail -2; done
