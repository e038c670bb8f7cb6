#!/bin/bash
# float4 SE order + fast batched SE, K2 resident weights up to 154 KB: full GPU tests, per-kernel times, bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c21_pytest.log
tail -8 gpurun_out/c21_pytest.log
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c21_kt.log 2>&1
grep -E "angles|total kernel|expand|\.kd|project|head|\.se|stem|\.k1|\.dw" gpurun_out/c21_kt.log | head -70
timeout 300 python bench.py --no-cpu > gpurun_out/c21_bench.json 2> gpurun_out/c21_bench.err
timeout 300 python bench.py --no-cpu --opt kd_tail=1 > gpurun_out/c21_bench_tail.json 2> gpurun_out/c21_bench_tail.err
timeout 300 python bench.py --no-cpu --opt streams=1 > gpurun_out/c21_bench_s1.json 2> gpurun_out/c21_bench_s1.err
for f in c21_bench c21_bench_tail c21_bench_s1; do tail -2 gpurun_out/$f.err; python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
"; done
