#!/bin/bash
# ncu: why are the 1x1 GEMM kernels slow on the late expand shapes?  K2 (persistent) and pw_tc2, plus KD
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=512 REPS=1 OPTS=streams=1,kd_expand_k2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_kernel -s 3 -c 4 -o /tmp/c18_k2 python tools/prof_run.py > gpurun_out/c18_ncu_k2.log 2>&1
python tools/ncu_summary.py /tmp/c18_k2.ncu-rep gpurun_out/c18_k2_summary.txt >> gpurun_out/c18_ncu_k2.log 2>&1
python tools/ncu_source.py /tmp/c18_k2.ncu-rep gpurun_out/c18_k2_source.txt 40 >> gpurun_out/c18_ncu_k2.log 2>&1
N=512 REPS=1 OPTS=streams=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pw_tc2_kernel|dwse_kernel" -s 15 -c 12 -o /tmp/c18_pw python tools/prof_run.py > gpurun_out/c18_ncu_pw.log 2>&1
python tools/ncu_summary.py /tmp/c18_pw.ncu-rep gpurun_out/c18_pw_summary.txt >> gpurun_out/c18_ncu_pw.log 2>&1
python tools/ncu_source.py /tmp/c18_pw.ncu-rep gpurun_out/c18_pw_source.txt 40 >> gpurun_out/c18_ncu_pw.log 2>&1
ls -la gpurun_out /tmp/*.ncu-rep | tail -12; tail -n 3 gpurun_out/c18_ncu_k2.log gpurun_out/c18_ncu_pw.log
