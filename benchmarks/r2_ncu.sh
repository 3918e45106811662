#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:forward_tiles_kernel -c 4 -o gpurun_out/prof_fwd2 python benchmarks/profile_kernels.py 1 r2fwd gelu:3,hardswish:3,softplus:3,gelu:1 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
