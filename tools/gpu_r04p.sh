#!/bin/bash
TAG=r04p
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
for c in C1 C2 C3; do timeout 600 python tools/ab_probe.py $c 2>&1 | cut -c1-200 | tee -a gpurun_out/${TAG}_ab.log; done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_depth -c 21 --csv --log-file gpurun_out/${TAG}_depth.csv python tools/prof_step.py C2 3 > /dev/null 2>&1
python - <<PY
import csv,re
rows=list(csv.reader(open('gpurun_out/${TAG}_depth.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); im=h.index('Metric Name'); iid=h.index('ID')
d={}
for r in rows[hdr+1:]:
    if len(r)<=iv: continue
    d.setdefault((int(r[iid]), re.sub(r'\(.*','',r[ik]).split('::')[-1]),{})[r[im]]=float(r[iv].replace(',',''))
for k in sorted(d)[-7:]: print(k[1], d[k])
PY
