#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/repro_pw3.py > gpurun_out/c33_repro.log 2>&1; cat gpurun_out/c33_repro.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c33_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c33_pytest.log
tail -6 gpurun_out/c33_pytest.log
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c33_kt.log 2>&1
grep -E "angles|total kernel|b0[1-6].project" gpurun_out/c33_kt.log
timeout 300 python bench.py --no-cpu > gpurun_out/c33_bench.json 2> gpurun_out/c33_bench.err
tail -2 gpurun_out/c33_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/c33_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
"
