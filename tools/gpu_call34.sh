#!/bin/bash
# re-tune the K1 plans of blocks 2-6 (HFMA2 build) at 256 crops per launch
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=256 REPS=3 BLOCKS=2,3,4,5,6 timeout 900 python tools/tune_k1.py > gpurun_out/c34_tune_k1.log 2>&1
tail -30 gpurun_out/c34_tune_k1.log
