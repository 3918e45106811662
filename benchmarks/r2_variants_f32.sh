#!/bin/bash
mkdir -p gpurun_out
export VARIANT_CELLS="gelu:f32:3,gelu:f32:7,logsigmoid:f32:3,softplus:f32:3,softplus:f32:7,selu:f32:3,selu:f32:7,hardswish:f32:3,silu:f32:3,sigmoid:f32:3,mish:f32:3,tanh:f32:3,tanhshrink:f32:3,softsign:f32:3,celu:f32:7,gelu:f32:1"
timeout 900 python benchmarks/variants.py fewbit_b200/libfewbit_b200.so fewbit_b200/libfewbit_b200_*.so > gpurun_out/variants_f32.txt 2>&1
cat gpurun_out/variants_f32.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
