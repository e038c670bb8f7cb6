#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/repro_pw.py 3232322 > gpurun_out/c28_repro.log 2>&1; cat gpurun_out/c28_repro.log
timeout 300 python tools/repro_pw.py 2222 >> gpurun_out/c28_repro.log 2>&1; tail -4 gpurun_out/c28_repro.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c28_pytest.log
tail -12 gpurun_out/c28_pytest.log
