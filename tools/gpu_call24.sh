#!/bin/bash
# ncu (full + source): KD without tail (b10, b13), gated K2 project (b10), batched SE (b13); per-kernel times with SE at 512 threads
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c24_kt.log 2>&1
grep -E "total kernel|\.se " gpurun_out/c24_kt.log | head -20
# launch order per forward (bf16, N=512, streams=1): dwse launches are b07..b16 -> b10 = 4th, b13 = 7th
N=512 REPS=1 OPTS=streams=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dwse_kernel -s 3 -c 4 -o /tmp/c24_kd python tools/prof_run.py > gpurun_out/c24_ncu.log 2>&1
python tools/ncu_summary.py /tmp/c24_kd.ncu-rep gpurun_out/c24_kd_summary.txt >> gpurun_out/c24_ncu.log 2>&1
python tools/ncu_source.py /tmp/c24_kd.ncu-rep gpurun_out/c24_kd_source.txt 45 >> gpurun_out/c24_ncu.log 2>&1
N=512 REPS=1 OPTS=streams=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k2_kernel<__nv_bfloat16, false, true|k2_kernel.*0, 1, 1|se_gate_batch" -s 6 -c 4 -o /tmp/c24_pj python tools/prof_run.py >> gpurun_out/c24_ncu.log 2>&1
python tools/ncu_summary.py /tmp/c24_pj.ncu-rep gpurun_out/c24_pj_summary.txt >> gpurun_out/c24_ncu.log 2>&1
python tools/ncu_source.py /tmp/c24_pj.ncu-rep gpurun_out/c24_pj_source.txt 45 >> gpurun_out/c24_ncu.log 2>&1
tail -5 gpurun_out/c24_ncu.log
