#!/bin/bash
# bench.py under torchrun at N GPUs, both arms (the driver's launch line).
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
tail -c 400 gpurun_out/bench_ref_n$N.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 300 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['roofline']['frac'],3),'cross',d.get('cross_device_check'),'launches',d['gpu_launches'])
r=json.loads(open('gpurun_out/bench_ref_n$N.json').read().strip().splitlines()[-1]); print('ref',round(r['value'],2),r['cpu_baseline']['cores'])
PY
