#!/bin/bash
# ncu (full set + source) of the early K1 launches (b01.dw, b02, b03, b04) in the current HFMA2 build, N = 128
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=128 REPS=1 OPTS=streams=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_expand_dw -c 4 -o /tmp/c23_k1 python tools/prof_run.py > gpurun_out/c23_ncu_k1.log 2>&1
python tools/ncu_summary.py /tmp/c23_k1.ncu-rep gpurun_out/c23_k1_summary.txt >> gpurun_out/c23_ncu_k1.log 2>&1
python tools/ncu_source.py /tmp/c23_k1.ncu-rep gpurun_out/c23_k1_source.txt 70 >> gpurun_out/c23_ncu_k1.log 2>&1
ncu -i /tmp/c23_k1.ncu-rep --page source --csv --print-source sass > /tmp/c23_sass.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('/tmp/c23_sass.csv')))
out = open('gpurun_out/c23_k1_opcodes.txt', 'w')
hdr = None; fn = None; agg = None
def flush():
    if agg:
        tot = sum(agg.values()) or 1
        out.write("## %s\n   warp instructions %.0f\n" % (fn, tot))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
            out.write("   %-28s %6.2f%%\n" % (k, 100 * v / tot))
        out.write("\n")
for r in rows:
    if not r: continue
    if r[0] == 'Function Name':
        flush(); fn = r[1][:140]; agg = collections.defaultdict(float); hdr = None
    elif 'Instructions Executed' in r and 'Source' in r:
        hdr = r
    elif hdr and agg is not None:
        try:
            src = r[hdr.index('Source')].strip(); n = float(r[hdr.index('Instructions Executed')] or 0)
        except (ValueError, IndexError):
            continue
        toks = src.split()
        if not toks: continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        agg[op.split('.')[0] + ('.' + op.split('.')[1] if '.' in op and op.split('.')[0] in ('LDS','STS','LDG','STG','MUFU','LDGSTS') else '')] += n
flush(); out.close()
PY
head -50 gpurun_out/c23_k1_opcodes.txt; head -5 /tmp/c23_sass.csv | cut -c1-300
