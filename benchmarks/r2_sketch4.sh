#!/bin/bash
# projection kernel: quick A/B (each process under its own timeout: a wrong barrier protocol shows up as a hang)
mkdir -p gpurun_out
export SWEEP_FEATURES=${SWEEP_FEATURES:-768,3072}
timeout 300 python benchmarks/sketch_sweep.py ${SWEEP_CONFIGS:-- SLOTS=3 SLOTS=2 DEBUG=1 DEBUG=4 DEBUG=5 PAIR=0} 2>&1 | tee gpurun_out/sketch_sweep4.txt
echo "exit $?"
FEWBIT_B200_SKETCH_TRACE=1 timeout 100 python benchmarks/sketch_trace.py 768 gaussian,rademacher 2>&1 | awk 'NR%4<2' | tee gpurun_out/sketch_trace4.txt | cut -c1-700
timeout 400 python -m pytest tests/test_gpu_sketch.py -x -q -m gpu 2>&1 | tail -5
