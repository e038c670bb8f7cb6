#!/bin/bash
# two GPUs: sharded == unsharded tests, weak-scaling bench at N=2 (both arms), strong scaling at fixed 1024 crops
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L > gpurun_out/c12_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s > gpurun_out/c12_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/c12_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c12_bench_n2.json 2> gpurun_out/c12_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --batch 256 > gpurun_out/c12_bench_n2_strong512.json 2> gpurun_out/c12_bench_n2_strong512.err
cat gpurun_out/c12_gpus.txt
tail -12 gpurun_out/c12_pytest_multi.log
tail -3 gpurun_out/c12_bench_n2.err
python -c "
import json
for f in ('c12_bench_n2','c12_bench_n2_strong512'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
"
