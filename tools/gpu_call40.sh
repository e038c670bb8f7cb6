#!/bin/bash
# final state: full GPU tests, bench (+CPU baseline), reference arm, launch list, per-kernel tables, stream-count A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c40_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c40_pytest.log
tail -5 gpurun_out/c40_pytest.log
timeout 400 python bench.py > gpurun_out/c40_bench.json 2> gpurun_out/c40_bench.err
tail -2 gpurun_out/c40_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/c40_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['get_angle_value'], d['gpu_launches'], d['self_check_max_deg_vs_simt_path'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['share_of_step'], d['clocks'])
print(d['cpu_baseline'])
"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c40_bench_ref.json 2> gpurun_out/c40_bench_ref.err
for st in 1 3 4; do timeout 300 python bench.py --no-cpu --opt streams=$st > gpurun_out/c40_bench_s$st.json 2> gpurun_out/c40_bench_s$st.err; python -c "
import json
d=json.loads(open('gpurun_out/c40_bench_s$st.json').read().strip().splitlines()[-1])
print('streams=$st', d['value'], d['ms_per_step'], d['e2e']['value'])
"; done
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c40_kt.log 2>&1
grep -E "angles|total kernel" gpurun_out/c40_kt.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c40_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/c40_ncu_bench.log 2>&1
