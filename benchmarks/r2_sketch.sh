#!/bin/bash
mkdir -p gpurun_out
timeout 600 python benchmarks/sketch_variants.py fewbit_b200/libfewbit_b200.so fewbit_b200/libfewbit_b200_sk*.so > gpurun_out/sketch_variants.txt 2>&1
cat gpurun_out/sketch_variants.txt
