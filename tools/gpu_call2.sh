#!/bin/bash
# round-2 call 2: K1W first contact - taps vs K1, per-block times, tests, ncu of the K1W launches
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 120 tools/bin/microbench > gpurun_out/c2_microbench.txt 2>&1
NT=3 N=256 timeout 300 python tools/k1w_check.py > gpurun_out/c2_k1w_check.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity_wide.py -m gpu -x -q -k "k1w" > gpurun_out/c2_pytest_k1w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest_k1w.log
timeout 300 python bench.py --no-cpu --opt k1_variant=4 > gpurun_out/c2_bench_k1w.json 2> gpurun_out/c2_bench_k1w.err
N=128 REPS=2 OPTS=k1_variant=4,streams=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1w_kernel -s 15 -c 15 -o gpurun_out/c2_k1w python tools/prof_run.py > gpurun_out/c2_ncu_k1w.log 2>&1
python tools/ncu_summary.py gpurun_out/c2_k1w.ncu-rep gpurun_out/c2_k1w_summary.txt >> gpurun_out/c2_ncu_k1w.log 2>&1
python tools/ncu_source.py gpurun_out/c2_k1w.ncu-rep gpurun_out/c2_k1w_source.txt 60 >> gpurun_out/c2_ncu_k1w.log 2>&1
rm -f gpurun_out/c1_k1.ncu-rep gpurun_out/c1_k1p.ncu-rep
cat gpurun_out/c2_k1w_check.log | tail -70
tail -15 gpurun_out/c2_pytest_k1w.log
cat gpurun_out/c2_bench_k1w.err | tail -5
