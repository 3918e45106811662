#!/bin/bash
mkdir -p gpurun_out
for bits in 3 7 1 5; do
export VARIANT_CELLS="celu:bf16:$bits,elu:bf16:$bits,gelu:bf16:$bits,hardswish:bf16:$bits,logsigmoid:bf16:$bits,mish:bf16:$bits,selu:bf16:$bits,sigmoid:bf16:$bits,silu:bf16:$bits,softplus:bf16:$bits,softsign:bf16:$bits,tanh:bf16:$bits,tanhshrink:bf16:$bits"
timeout 600 python benchmarks/variants.py fewbit_b200/libfewbit_b200.so fewbit_b200/libfewbit_b200_pf0.so >> gpurun_out/variants_fn.txt 2>&1
done
cat gpurun_out/variants_fn.txt
