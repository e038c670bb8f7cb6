#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=256 timeout 300 python tools/k1w_trace.py > gpurun_out/c8_k1w_trace.log 2>&1
NT=3 N=256 timeout 300 python tools/k1w_check.py > gpurun_out/c8_k1w_check.log 2>&1
timeout 600 python -m pytest tests/test_gpu_conv1x1.py -m gpu -x -q > gpurun_out/c8_pytest_conv.log 2>&1; echo "rc=$?" >> gpurun_out/c8_pytest_conv.log
timeout 300 python bench.py --no-cpu --opt pw_variant=3 > gpurun_out/c8_bench_k2.json 2> gpurun_out/c8_bench_k2.err
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,pw_variant=3 timeout 200 python tools/gpu_check.py > gpurun_out/c8_kernel_times_k2.log 2>&1
tail -18 gpurun_out/c8_k1w_check.log
cat gpurun_out/c8_k1w_trace.log
tail -15 gpurun_out/c8_pytest_conv.log
tail -3 gpurun_out/c8_bench_k2.err
grep -E "project|head.conv|total kernel" gpurun_out/c8_kernel_times_k2.log
python -c "
import json
d=json.loads(open('gpurun_out/c8_bench_k2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})
"
