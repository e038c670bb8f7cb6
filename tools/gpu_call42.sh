#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_conv1x1.py tests/test_gpu_parity.py -m gpu -q -k "pw_tc3 or se_batch_and_k2 or pageable or kd_route" > gpurun_out/c42_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c42_pytest.log
tail -4 gpurun_out/c42_pytest.log
timeout 200 python bench.py --no-cpu --steps 20 > gpurun_out/c42_bench.json 2> gpurun_out/c42_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/c42_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['get_angle_value'], d['gpu_launches'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})
"
