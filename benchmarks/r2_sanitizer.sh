#!/bin/bash
# compute-sanitizer over every kernel family + the pair-mode soak test (round-2 kernels)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python benchmarks/sanitizer_workload.py > gpurun_out/memcheck.log 2>&1; tail -3 gpurun_out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python benchmarks/sanitizer_workload.py > gpurun_out/racecheck.log 2>&1; tail -3 gpurun_out/racecheck.log
grep -c "hazard" gpurun_out/racecheck.log
timeout 600 python benchmarks/sketch_stress.py 600 2>&1 | tail -2
