#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest.log
timeout 300 python bench.py > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
FULL=1 PRECS=bf16,fp16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c13_kernel_times.log 2>&1
NT=3 N=256 timeout 300 python tools/k1w_check.py > gpurun_out/c13_k1w_check.log 2>&1
tail -12 gpurun_out/c13_pytest.log
tail -3 gpurun_out/c13_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/c13_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['get_angle_value'], d['gpu_launches'], d['self_check_max_deg_vs_simt_path'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})
print(d['cpu_baseline'])
"
grep -E "angles|k1 |total kernel" gpurun_out/c13_kernel_times.log | head -40
tail -20 gpurun_out/c13_k1w_check.log
