#!/bin/bash
# ncu (full + source): K2 gated project of block 10 (8th k2 launch of a forward) and block 13 (14th), expand of block 10 (7th)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=512 REPS=1 OPTS=streams=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_kernel -s 6 -c 2 -o /tmp/c25_a python tools/prof_run.py > gpurun_out/c25_ncu.log 2>&1
python tools/ncu_summary.py /tmp/c25_a.ncu-rep gpurun_out/c25_k2_b10_summary.txt >> gpurun_out/c25_ncu.log 2>&1
python tools/ncu_source.py /tmp/c25_a.ncu-rep gpurun_out/c25_k2_b10_source.txt 45 >> gpurun_out/c25_ncu.log 2>&1
ncu -i /tmp/c25_a.ncu-rep --page raw --csv > /tmp/c25_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/c25_raw.csv')))
hdr = rows[0]
want = [h for h in hdr if any(k in h for k in ("l1tex__m_xbar2l1tex_read_bytes", "lts__t_sectors_srcunit_tex_op_read", "lts__t_bytes.sum", "lts__t_sector_hit_rate", "sm__inst_executed_pipe_uniform", "smsp__warps_issue_stalled", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__throughput", "l1tex__throughput", "gpu__compute_memory_throughput"))]
with open('gpurun_out/c25_k2_b10_extra.txt', 'w') as f:
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        f.write("## %s\n" % d.get("Kernel Name", "?")[:100])
        for k in want:
            f.write("%-80s %s\n" % (k, d[k]))
PY
tail -3 gpurun_out/c25_ncu.log
