#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=256 timeout 900 python tools/tune_k1.py > gpurun_out/c14_tune_k1.log 2>&1
timeout 300 python bench.py --no-cpu > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
timeout 300 python bench.py --no-cpu --opt streams=1 > gpurun_out/c14_bench_s1.json 2> gpurun_out/c14_bench_s1.err
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "16bit or fused_k1_taps or sample" > gpurun_out/c14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c14_pytest.log
tail -40 gpurun_out/c14_tune_k1.log
tail -4 gpurun_out/c14_pytest.log
for f in c14_bench c14_bench_s1; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['e2e']['value'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})
"; done
