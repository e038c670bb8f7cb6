#!/bin/bash
# KD route (late blocks: expand GEMM + depthwise/SE kernel): tests, per-kernel times, bench A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kd_route or se_tail or batch_invariance or fused_k1_taps or two_stream or sample_angles_16bit" > gpurun_out/c16_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c16_pytest.log
tail -15 gpurun_out/c16_pytest.log
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1 timeout 300 python tools/gpu_check.py > gpurun_out/c16_kt_kd.log 2>&1
FULL=1 PRECS=bf16 TCS=1 OPTS=streams=1,kd_expand_k2=1 timeout 300 python tools/gpu_check.py > gpurun_out/c16_kt_kd_k2.log 2>&1
for f in c16_kt_kd c16_kt_kd_k2; do echo == $f; grep -E "angles|total kernel|expand|\.kd|k1 |project" gpurun_out/$f.log | head -60; done
timeout 300 python bench.py --no-cpu > gpurun_out/c16_bench_kd.json 2> gpurun_out/c16_bench_kd.err
timeout 300 python bench.py --no-cpu --opt kd_expand_k2=1 > gpurun_out/c16_bench_kd_k2.json 2> gpurun_out/c16_bench_kd_k2.err
timeout 300 python bench.py --no-cpu --opt kd_from=0 > gpurun_out/c16_bench_k1.json 2> gpurun_out/c16_bench_k1.err
for f in c16_bench_kd c16_bench_kd_k2 c16_bench_k1; do tail -2 gpurun_out/$f.err; python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
print({k:round(v['ms_per_step'],3) for k,v in d['roofline']['families'].items()})
"; done
