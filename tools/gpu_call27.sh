#!/bin/bash
# KD on TMA + block-1 depthwise on KD (fp16 stem output) + batched head: full GPU tests, bench, per-kernel times
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c27_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c27_pytest.log
tail -6 gpurun_out/c27_pytest.log
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c27_kt.log 2>&1
grep -E "angles|total kernel|head" gpurun_out/c27_kt.log
timeout 400 python bench.py > gpurun_out/c27_bench.json 2> gpurun_out/c27_bench.err
tail -2 gpurun_out/c27_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/c27_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['get_angle_value'], d['gpu_launches'], d['self_check_max_deg_vs_simt_path'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['share_of_step'])
print(d['cpu_baseline'])
"
timeout 300 python tools/latency.py > gpurun_out/c27_latency.json 2> gpurun_out/c27_latency.err; tail -3 gpurun_out/c27_latency.json | cut -c1-600
