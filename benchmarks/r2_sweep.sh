#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python benchmarks/sweep_functions.py --json gpurun_out/sweep_all.json --md gpurun_out/sweep_all.md > gpurun_out/sweep_all.log 2>&1
tail -2 gpurun_out/sweep_all.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
