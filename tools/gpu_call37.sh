#!/bin/bash
# late-phase kernels at 128 / 256 / 512 crops per launch (streams=1): per-crop cost vs working-set size (L2 residency of E and D)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for n in 128 256 512; do
NPROF=$n FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c37_kt_n$n.log 2>&1
done
python - <<'PY'
import re
for n in (128,256,512):
    rows={}
    for l in open('gpurun_out/c37_kt_n%d.log'%n):
        m=re.match(r"\s+(b\d\d\.\w+|stem|head\.\w+)\s+([\d.]+) ms",l)
        if m: rows[m.group(1)]=float(m.group(2))
    fam={}
    for k,v in rows.items():
        if k[0]=='b' and int(k[1:3])>=7:
            fam[k.split('.')[1]]=fam.get(k.split('.')[1],0)+v
    tot=sum(rows.values())
    print(n, 'total %.3f (x%d = %.3f)'%(tot, 512//n, tot*512/n), 'late per 512 crops:', {k:round(v*512/n,3) for k,v in fam.items()}, 'b10:', {k:round(v*512/n,3) for k,v in rows.items() if k.startswith('b10')})
PY
