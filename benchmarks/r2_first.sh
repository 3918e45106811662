#!/bin/bash
# Round-2 check on one B200: parity tests, bf16/f32 forward sweep of the reworked kernels, ncu of three of them.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python benchmarks/sweep_functions.py --bits 1,2,3,4,5,6,7,8 --dtypes bf16 --json gpurun_out/sweep_bf16.json --md gpurun_out/sweep_bf16.md > gpurun_out/sweep_bf16.log 2>&1
tail -3 gpurun_out/sweep_bf16.log
timeout 600 python benchmarks/sweep_functions.py --bits 5,6,7,8 --dtypes bf16 --tables synthetic --json gpurun_out/sweep_bf16_synth.json --md gpurun_out/sweep_bf16_synth.md > gpurun_out/sweep_bf16_synth.log 2>&1
timeout 600 python benchmarks/sweep_functions.py --bits 3,5,7 --dtypes f32 --json gpurun_out/sweep_f32.json --md gpurun_out/sweep_f32.md > gpurun_out/sweep_f32.log 2>&1
kill $SMI
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_tiles_kernel -c 6 -o gpurun_out/prof_fwd python benchmarks/profile_kernels.py 1 r2fwd > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
