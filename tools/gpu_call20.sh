#!/bin/bash
# K2 256-bit stores, batched SE gate, KD without tail + gated K2 projects: full GPU tests, per-kernel times, bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c20_pytest.log
tail -8 gpurun_out/c20_pytest.log
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c20_kt.log 2>&1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,kd_tail=1 timeout 300 python tools/gpu_check.py > gpurun_out/c20_kt_tail.log 2>&1
for f in c20_kt c20_kt_tail; do echo == $f; grep -E "angles|total kernel|expand|\.kd|project|head|\.se" gpurun_out/$f.log | head -70; done
timeout 300 python bench.py --no-cpu > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
timeout 300 python bench.py --no-cpu --opt kd_tail=1 > gpurun_out/c20_bench_tail.json 2> gpurun_out/c20_bench_tail.err
timeout 300 python bench.py --no-cpu --opt streams=1 > gpurun_out/c20_bench_s1.json 2> gpurun_out/c20_bench_s1.err
for f in c20_bench c20_bench_tail c20_bench_s1; do tail -2 gpurun_out/$f.err; python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
"; done
