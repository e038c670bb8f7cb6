#!/bin/bash
# KD route: does the late phase become L2-resident at smaller batches?  per-kernel times at N = 512 / 256 / 128 / 64, gate placement
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for n in 256 128 64; do
NPROF=$n FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,k1_split_ctas=0 timeout 300 python tools/gpu_check.py > gpurun_out/c17_kt_n$n.log 2>&1
NPROF=$n FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,k1_split_ctas=0,kd_from=0 timeout 300 python tools/gpu_check.py > gpurun_out/c17_kt_k1_n$n.log 2>&1
done
NPROF=512 FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,se_scale_out=0 timeout 300 python tools/gpu_check.py > gpurun_out/c17_kt_n512_noscale.log 2>&1
NPROF=128 FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,se_scale_out=0,k1_split_ctas=0 timeout 300 python tools/gpu_check.py > gpurun_out/c17_kt_n128_noscale.log 2>&1
NPROF=128 FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,k1_split_ctas=400 timeout 300 python tools/gpu_check.py > gpurun_out/c17_kt_n128_split.log 2>&1
python - <<'PY'
import re,glob
for f in sorted(glob.glob('gpurun_out/c17_kt_*.log')):
    rows={}
    tot=None
    for l in open(f):
        m=re.match(r"\s+(b\d\d\.\w+|stem|head\.\w+)\s+([\d.]+) ms",l)
        if m: rows[m.group(1)]=float(m.group(2))
        if 'total kernel' in l: tot=l.strip()
    late=sum(v for k,v in rows.items() if k[0]=='b' and int(k[1:3])>=7)
    print(f, tot, 'late(b07-16) ms %.3f'%late)
    print('   ', ' '.join('%s=%.3f'%(k,v) for k,v in rows.items() if k[0]=='b' and int(k[1:3]) in (7,10,13)))
PY
