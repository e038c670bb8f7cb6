#!/bin/bash
# L2-residency probe for the early blocks: unfused route (expand GEMM -> E -> depthwise) at small batches
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for n in 16 32 64 128; do
NPROF=$n FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,fused=0,k1_split_ctas=0 timeout 300 python tools/gpu_check.py > gpurun_out/c22_unfused_n$n.log 2>&1
echo "== N=$n"; grep -E "total kernel|b0[2-6]\.(expand|dw|k1)" gpurun_out/c22_unfused_n$n.log
done
NPROF=32 FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,k1_split_ctas=0 timeout 300 python tools/gpu_check.py > gpurun_out/c22_fused_n32.log 2>&1
echo "== fused N=32"; grep -E "total kernel|b0[2-6]\.(expand|dw|k1)" gpurun_out/c22_fused_n32.log
