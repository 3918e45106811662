#!/bin/bash
mkdir -p gpurun_out
FEWBIT_B200_SKETCH_TRACE=1 timeout 120 python benchmarks/sketch_trace.py 768,3072 gaussian,rademacher > gpurun_out/sketch_trace.txt 2>&1
cat gpurun_out/sketch_trace.txt
for mask in 0 1 4 5; do echo "debug mask $mask"; FEWBIT_B200_SKETCH_DEBUG=$mask timeout 120 python benchmarks/sketch_variants.py fewbit_b200/libfewbit_b200.so 2>&1 | cut -c1-200; done
