#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 tools/bin/microbench > gpurun_out/c6_microbench.txt 2>&1
N=256 timeout 300 python tools/k1w_trace.py > gpurun_out/c6_k1w_trace.log 2>&1
NT=3 N=256 timeout 300 python tools/k1w_check.py > gpurun_out/c6_k1w_check.log 2>&1
tail -18 gpurun_out/c6_k1w_check.log
cat gpurun_out/c6_k1w_trace.log
tail -8 gpurun_out/c6_microbench.txt
