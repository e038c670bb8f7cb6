#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NT=3 N=256 timeout 300 python tools/k1w_check.py > gpurun_out/c4_k1w_check.log 2>&1
N=256 timeout 300 python tools/k1w_trace.py > gpurun_out/c4_k1w_trace.log 2>&1
N=256 timeout 900 python tools/tune_k1w.py > gpurun_out/c4_tune_k1w.log 2>&1
N=128 REPS=2 OPTS=k1_variant=4,streams=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1w_kernel -s 15 -c 6 -o gpurun_out/c4_k1w python tools/prof_run.py > gpurun_out/c4_ncu_k1w.log 2>&1
python tools/ncu_summary.py gpurun_out/c4_k1w.ncu-rep gpurun_out/c4_k1w_summary.txt >> gpurun_out/c4_ncu_k1w.log 2>&1
python tools/ncu_source.py gpurun_out/c4_k1w.ncu-rep gpurun_out/c4_k1w_source.txt 60 >> gpurun_out/c4_ncu_k1w.log 2>&1
rm -f gpurun_out/c2_k1w.ncu-rep
tail -22 gpurun_out/c4_k1w_check.log
cat gpurun_out/c4_k1w_trace.log
cat gpurun_out/c4_tune_k1w.log
