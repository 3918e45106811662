#!/bin/bash
# projection kernel with the unit ring / slot ring: correctness first, then the knobs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sketch.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python benchmarks/sketch_sweep.py - UNITS=9 UNITS=9,SLOTS=2 SLOTS=2 SLOTS=4 PREFETCH=2 PREFETCH=4 PREFETCH=8 DEBUG=1 DEBUG=4 DEBUG=5 PAIR=0 2>&1 | tee gpurun_out/sketch_sweep.txt
FEWBIT_B200_SKETCH_TRACE=1 timeout 120 python benchmarks/sketch_trace.py 768,3072 gaussian,rademacher 2>&1 | tee gpurun_out/sketch_trace3.txt | cut -c1-700
