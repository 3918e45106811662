#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
    for r in d['rooflines']: print(r['kernel'], round(r['achieved'],1), r['unit'], round(r['frac'],3), r['traffic'])
    print(json.dumps(d['extra'].get('roberta'),indent=0)[:1500])
    print(d['cpu_baseline'])
except Exception as e: print('bench parse failed',e)
PY
