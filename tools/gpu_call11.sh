#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv1x1.py tests/test_gpu_parity_wide.py -m gpu -x -q -k "fp32 or packed" -s > gpurun_out/c11_pytest_tc32.log 2>&1; echo "rc=$?" >> gpurun_out/c11_pytest_tc32.log
timeout 300 python bench.py --no-cpu --precision fp32 > gpurun_out/c11_bench_fp32_simt.json 2> gpurun_out/c11_bench_fp32_simt.err
timeout 300 python bench.py --no-cpu --precision fp32 --opt tensor_cores=1 > gpurun_out/c11_bench_fp32_tc.json 2> gpurun_out/c11_bench_fp32_tc.err
tail -25 gpurun_out/c11_pytest_tc32.log
for f in gpurun_out/c11_bench_fp32_simt gpurun_out/c11_bench_fp32_tc; do tail -2 $f.err; python -c "
import json,sys
d=json.loads(open('$f.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['self_check_max_deg_vs_simt_path'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})
"; done
