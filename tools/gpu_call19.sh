#!/bin/bash
# K2 with the bias row in smem + two epilogue groups: conv1x1 tests, per-kernel times (expand on K2 / everything on K2)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv1x1.py tests/test_gpu_parity.py -m gpu -x -q -k "conv1x1 or k2 or K2 or kd_route or pw_variant" > gpurun_out/c19_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c19_pytest.log
tail -8 gpurun_out/c19_pytest.log
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,kd_expand_k2=1 timeout 300 python tools/gpu_check.py > gpurun_out/c19_kt_expk2.log 2>&1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,kd_expand_k2=1,pw_variant=3 timeout 300 python tools/gpu_check.py > gpurun_out/c19_kt_allk2.log 2>&1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,kd_expand_k2=1,pw_variant=3,se_scale_out=0 timeout 300 python tools/gpu_check.py > gpurun_out/c19_kt_allk2_gated.log 2>&1
for f in c19_kt_expk2 c19_kt_allk2 c19_kt_allk2_gated; do echo == $f; grep -E "angles|total kernel|expand|\.kd|project|head" gpurun_out/$f.log | head -60; done
