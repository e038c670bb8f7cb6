#!/bin/bash
# 2 x B200: multi-GPU tests (sharded == unsharded bitwise, NCCL gather) and the weak-scaling bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/c35_smi.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/c35_pytest_multi_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c35_pytest_multi_gpu.log
tail -5 gpurun_out/c35_pytest_multi_gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c35_bench_n2.json 2> gpurun_out/c35_bench_n2.err
tail -2 gpurun_out/c35_bench_n2.err; python -c "
import json
d=json.loads(open('gpurun_out/c35_bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
"
