#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NT=3 N=256 timeout 300 python tools/k1w_check.py > gpurun_out/c5_k1w_check.log 2>&1
N=256 timeout 300 python tools/k1w_trace.py > gpurun_out/c5_k1w_trace.log 2>&1
N=256 timeout 900 python tools/tune_k1w.py > gpurun_out/c5_tune_k1w.log 2>&1
rm -f gpurun_out/c4_k1w.ncu-rep
tail -22 gpurun_out/c5_k1w_check.log
cat gpurun_out/c5_k1w_trace.log
cat gpurun_out/c5_tune_k1w.log
