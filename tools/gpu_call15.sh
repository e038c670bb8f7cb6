#!/bin/bash
# round-2 re-entry call: state of HEAD (tests, bench), unfused late blocks vs K1, launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c15_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest.log
timeout 400 python bench.py > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c15_kernel_times.log 2>&1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,fused_max_block=6 timeout 300 python tools/gpu_check.py > gpurun_out/c15_kernel_times_unfused_late.log 2>&1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,fused=0 timeout 300 python tools/gpu_check.py > gpurun_out/c15_kernel_times_unfused.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c15_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/c15_ncu_bench.log 2>&1
tail -6 gpurun_out/c15_pytest.log
tail -3 gpurun_out/c15_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/c15_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})
print(d['cpu_baseline'])
"
for f in c15_kernel_times c15_kernel_times_unfused_late c15_kernel_times_unfused; do echo == $f; grep -E "angles|total kernel|expand|\.dw|k1 " gpurun_out/$f.log | head -50; done
