#!/bin/bash
# staged upload of pageable inputs: tests + bench (get_angle figure)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c41_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c41_pytest.log
tail -5 gpurun_out/c41_pytest.log
timeout 400 python bench.py > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err
tail -2 gpurun_out/c41_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/c41_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['get_angle_value'], d['gpu_launches'])
"
timeout 300 python bench.py --no-cpu --opt stage_threads=0 > gpurun_out/c41_bench_nostage.json 2> gpurun_out/c41_bench_nostage.err; python -c "
import json
d=json.loads(open('gpurun_out/c41_bench_nostage.json').read().strip().splitlines()[-1])
print('stage_threads=0', d['value'], d['e2e']['value'], d['e2e']['get_angle_value'])
"
timeout 300 python bench.py --no-cpu --opt stage_threads=16 > gpurun_out/c41_bench_st16.json 2> gpurun_out/c41_bench_st16.err; python -c "
import json
d=json.loads(open('gpurun_out/c41_bench_st16.json').read().strip().splitlines()[-1])
print('stage_threads=16', d['value'], d['e2e']['value'], d['e2e']['get_angle_value'])
"
