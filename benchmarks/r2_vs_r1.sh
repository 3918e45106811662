#!/bin/bash
# Same-method comparison of the round-1 and round-2 forward kernels (CUDA-graph replayed back-to-back launches),
# then the bench line with graph-replayed cold groups.
mkdir -p gpurun_out
export VARIANT_CELLS="gelu:bf16:3,gelu:bf16:5,gelu:bf16:7,gelu:bf16:8,hardswish:bf16:3,hardswish:bf16:7,tanh:bf16:3,tanhshrink:bf16:3,selu:bf16:3,softplus:bf16:3,silu:bf16:3,gelu:bf16:1,gelu:f32:3,softplus:f32:3,gelu:f32:7,softsign:bf16:7"
timeout 900 python benchmarks/variants.py fewbit_b200/libfewbit_b200_r01.so fewbit_b200/libfewbit_b200.so > gpurun_out/r1_vs_r2.txt 2>&1
cat gpurun_out/r1_vs_r2.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'])
for r in d['rooflines']: print(r['kernel'], round(r['achieved'],1), round(r['frac'],3))
PY
