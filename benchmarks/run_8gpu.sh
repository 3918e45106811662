set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
timeout 300 $TR bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 600 gpurun_out/bench_n8.json
timeout 400 $TR benchmarks/sweep_functions.py --bits 1,2,3,4,5,6,7,8 --md gpurun_out/sweep_n8.md --json gpurun_out/sweep_n8.json > gpurun_out/sweep_n8.txt 2>&1; tail -3 gpurun_out/sweep_n8.txt
for v in "both bf16" "vanilla bf16" "both fp32" "vanilla fp32"; do set -- $v; timeout 300 $TR benchmarks/roberta_step.py --ddp --variant $1 --dtype $2 --steps 5 2>/dev/null | tail -1 | tee -a gpurun_out/roberta_ddp_n8.jsonl; done
