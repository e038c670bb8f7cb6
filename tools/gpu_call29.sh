#!/bin/bash
# state check: full GPU tests, bench (with CPU baseline), reference arm, launch list, per-kernel times, stream + latency tools
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c29_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c29_pytest.log
tail -5 gpurun_out/c29_pytest.log
timeout 400 python bench.py > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
tail -2 gpurun_out/c29_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/c29_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['get_angle_value'], d['gpu_launches'], d['self_check_max_deg_vs_simt_path'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['share_of_step'], d['clocks'])
"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c29_bench_ref.json 2> gpurun_out/c29_bench_ref.err; tail -1 gpurun_out/c29_bench_ref.json | cut -c1-300
FULL=1 PRECS=bf16,fp16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c29_kt.log 2>&1
grep -E "angles|total kernel" gpurun_out/c29_kt.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c29_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/c29_ncu_bench.log 2>&1
timeout 300 python tools/latency.py > gpurun_out/c29_latency.json 2> gpurun_out/c29_latency.err
timeout 300 python tools/stream_bench.py > gpurun_out/c29_stream.json 2> gpurun_out/c29_stream.err; tail -1 gpurun_out/c29_stream.json | cut -c1-300
timeout 300 python bench.py --batch 32 --no-cpu > gpurun_out/c29_bench_b32.json 2> gpurun_out/c29_bench_b32.err; python -c "
import json
d=json.loads(open('gpurun_out/c29_bench_b32.json').read().strip().splitlines()[-1])
print('batch32', d['value'], d['ms_per_step'])
"
