#!/bin/bash
# round-2 call 1: microbench, GPU tests, bench, K1 vs K1P timings, ncu of the K1P kernels and of the late K1 blocks
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 120 tools/bin/microbench > gpurun_out/c1_microbench.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 300 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
timeout 300 python tools/k1p_check.py > gpurun_out/c1_k1p_check.log 2>&1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 200 python tools/gpu_check.py > gpurun_out/c1_kernel_times.log 2>&1
# ncu: K1P kernels (blocks 2-6) at 128 crops
N=128 REPS=2 OPTS=k1_variant=3,streams=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1p_kernel -s 5 -c 5 -o gpurun_out/c1_k1p python tools/prof_run.py > gpurun_out/c1_ncu_k1p.log 2>&1
python tools/ncu_summary.py gpurun_out/c1_k1p.ncu-rep gpurun_out/c1_k1p_summary.txt >> gpurun_out/c1_ncu_k1p.log 2>&1
python tools/ncu_source.py gpurun_out/c1_k1p.ncu-rep gpurun_out/c1_k1p_source.txt 60 >> gpurun_out/c1_ncu_k1p.log 2>&1
# ncu: the K1 launches of blocks 7-16 (second forward)
N=128 REPS=2 OPTS=streams=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_expand_dw -s 16 -c 16 -o gpurun_out/c1_k1 python tools/prof_run.py > gpurun_out/c1_ncu_k1.log 2>&1
python tools/ncu_summary.py gpurun_out/c1_k1.ncu-rep gpurun_out/c1_k1_summary.txt >> gpurun_out/c1_ncu_k1.log 2>&1
python tools/ncu_source.py gpurun_out/c1_k1.ncu-rep gpurun_out/c1_k1_source.txt 50 >> gpurun_out/c1_ncu_k1.log 2>&1
ls -la gpurun_out/ | tail -30
tail -5 gpurun_out/c1_pytest.log
cat gpurun_out/c1_microbench.txt
cat gpurun_out/c1_k1p_check.log | tail -25
