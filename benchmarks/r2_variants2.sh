#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python benchmarks/variants.py fewbit_b200/libfewbit_b200.so fewbit_b200/libfewbit_b200_*.so > gpurun_out/variants.txt 2>&1
cat gpurun_out/variants.txt
export VARIANT_CELLS="gelu:bf16:4,hardswish:bf16:4,silu:bf16:4,gelu:bf16:6,celu:bf16:2,selu:bf16:2,mish:bf16:5,tanh:bf16:8,hardswish:bf16:8,sigmoid:bf16:3,softsign:bf16:7,softsign:bf16:8,softsign:f32:8,mish:bf16:3"
timeout 900 python benchmarks/variants.py fewbit_b200/libfewbit_b200.so fewbit_b200/libfewbit_b200_*.so > gpurun_out/variants_b.txt 2>&1
cat gpurun_out/variants_b.txt
