#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
tail -4 gpurun_out/pytest_gpu2.log
timeout 900 python benchmarks/variants.py fewbit_b200/libfewbit_b200.so fewbit_b200/libfewbit_b200_*.so > gpurun_out/variants.txt 2>&1
cat gpurun_out/variants.txt
