#!/bin/bash
# RoBERTa-base step under DistributedDataParallel at N GPUs (round-2 kernels)
N=${1:-2}
mkdir -p gpurun_out; rm -f gpurun_out/roberta_ddp_n$N.jsonl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for v in "both bf16" "vanilla bf16" "both fp32" "vanilla fp32"; do set -- $v; timeout 300 $TR benchmarks/roberta_step.py --ddp --variant $1 --dtype $2 --steps 5 2>/dev/null | tail -1 | tee -a gpurun_out/roberta_ddp_n$N.jsonl; done
