#!/bin/bash
# compute-sanitizer over one bf16 forward at 70 crops (batched SE / head kernels, K2 + pw_tc2 mix, KD with chunk split off and on, pw_tc3)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=70 REPS=1 OPTS=streams=1 timeout 900 compute-sanitizer --tool memcheck python tools/prof_run.py > gpurun_out/c36_memcheck.log 2>&1; tail -4 gpurun_out/c36_memcheck.log
N=70 REPS=1 OPTS=streams=1,k1_split_ctas=0 timeout 900 compute-sanitizer --tool memcheck python tools/prof_run.py > gpurun_out/c36_memcheck_nosplit.log 2>&1; tail -4 gpurun_out/c36_memcheck_nosplit.log
N=70 REPS=1 OPTS=streams=1 timeout 1200 compute-sanitizer --tool racecheck python tools/prof_run.py > gpurun_out/c36_racecheck.log 2>&1; tail -6 gpurun_out/c36_racecheck.log
