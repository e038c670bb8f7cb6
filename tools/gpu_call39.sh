#!/bin/bash
# tensor-core stem: taps vs oracle, tests that touch the stem, per-kernel times, bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c39_kt.log 2>&1
grep -E "angles|rel-|total kernel|stem|b01" gpurun_out/c39_kt.log
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,stem_tc=0 timeout 300 python tools/gpu_check.py > gpurun_out/c39_kt_off.log 2>&1
grep -E "angles|total kernel|stem " gpurun_out/c39_kt_off.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_wide.py -m gpu -x -q > gpurun_out/c39_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c39_pytest.log
tail -6 gpurun_out/c39_pytest.log
timeout 300 python bench.py --no-cpu > gpurun_out/c39_bench.json 2> gpurun_out/c39_bench.err
tail -2 gpurun_out/c39_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/c39_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['self_check_max_deg_vs_simt_path'])
"
