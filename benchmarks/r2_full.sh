#!/bin/bash
# Round-2 record run on one B200: GPU tests, smoke, function sweep, bench line, ncu launch list of the bench
# command, ncu --set full rows of the kernels the bench reports rooflines for.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python benchmarks/sweep_functions.py --json gpurun_out/sweep_all.json --md gpurun_out/sweep_all.md > gpurun_out/sweep_all.log 2>&1
tail -1 gpurun_out/sweep_all.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'tiles_kernel|sketch_kernel' -c 14 -o gpurun_out/prof_r02 python benchmarks/profile_kernels.py 1 all > gpurun_out/ncu_all.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_tiles_kernel -c 6 -o gpurun_out/prof_r02_fwd python benchmarks/profile_kernels.py 1 r2fwd gelu:3,gelu:7,hardswish:7,gelu:8,hardswish:3,tanh:3 > gpurun_out/ncu_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
