"""RoBERTa-base training step: time and peak memory with FewBit components swapped in
(BASELINE.json configs[3] and [4]; shapes of reference benchmark/bench-roberta.py and
bench-linear.py, measured the reference's way: benchmark/benchmark.py:165-188).

    python benchmarks/roberta_step.py [--batch 128] [--seq 128] [--steps 5] [--dtype fp32|bf16]
                                      [--variants vanilla,gelu3,rand0.2,both] [--json out.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           benchmarks/roberta_step.py --ddp --variant both        # configs[4]: data parallel

Random-init `RobertaForSequenceClassification` (no network: no checkpoint, no GLUE), synthetic
`input_ids`, AdamW lr 2e-5 wd 0.01 (bench-roberta.py:84-93).  Each variant runs in a child
process so that peak memory is isolated (the reference forks per case for the same reason).
Reported per variant: median step ms (CUDA events), peak GiB = max_memory_allocated after the
steps minus memory_allocated before the model is built, and the reduction against vanilla next
to the reference's published reduction (README.md:18-27: -13.8 %, -18.6 %, -32.7 %).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
PUBLISHED = {'gelu3': -13.8, 'rand0.2': -18.6, 'both': -32.7}   # reference README.md:18-27, percent; 'both-shared' = both + q/k/v share one sketch


def build_model(variant: str, dtype):
    import torch
    from transformers import RobertaConfig, RobertaForSequenceClassification
    from transformers.activations import GELUActivation

    import fewbit_b200 as fewbit
    config = RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1,
                           hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, num_labels=2)
    model = RobertaForSequenceClassification(config).to('cuda', dtype)
    swapped = {'gelu': 0, 'linear': 0}
    if variant in ('gelu3', 'both', 'both-shared'):
        def swap_gelu(module, path):
            if isinstance(module, (GELUActivation, torch.nn.GELU)):
                swapped['gelu'] += 1
                return fewbit.GELU(bits=3)
            return module
        model = fewbit.map_module(model, swap_gelu)
    if variant in ('rand0.2', 'both', 'both-shared'):
        def swap_linear(module, path):     # benchmark/bench-linear.py:138-144
            out = fewbit.convert_linear(module, fewbit.RandomizedLinear, proj_dim_ratio=0.2,
                                        proj_dim_min=3, share_sketch=variant == 'both-shared')
            swapped['linear'] += out is not module
            return out
        model = fewbit.map_module(model, swap_linear)
    return model, swapped


def child(args):
    import torch
    sys.path.insert(0, str(ROOT))
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if args.ddp:
        # Data parallel: every rank runs the same per-GPU batch (weak scaling); the quantized
        # activations and the sketches stay local to the GPU that produced them, the only
        # communication is DDP's gradient all-reduce over NCCL (plumbing).
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        dist.init_process_group('nccl')
    torch.manual_seed(rank)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    torch.cuda.init()
    base = torch.cuda.memory_allocated()
    model, swapped = build_model(args.variant, dtype)
    model.train()
    if args.ddp:
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[torch.cuda.current_device()])
    opt = torch.optim.AdamW(model.parameters(), lr=2e-5, weight_decay=0.01)
    ids = torch.randint(0, 50265, (args.batch, args.seq), device='cuda')
    labels = torch.randint(0, 2, (args.batch, ), device='cuda')
    mask = torch.ones_like(ids)
    times, losses = [], []
    for step in range(args.steps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        loss = model(input_ids=ids, attention_mask=mask, labels=labels).loss
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        b.record()
        torch.cuda.synchronize()
        losses.append(float(loss))
        if step >= 2:
            times.append(a.elapsed_time(b))
    peak = torch.cuda.max_memory_allocated() - base
    params = sum(p.numel() for p in model.parameters())
    step_ms = statistics.median(times)
    if args.ddp:
        import torch.distributed as dist
        t = torch.tensor([step_ms, peak / 2 ** 30], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, peak = t[0].item(), t[1].item() * 2 ** 30
    if rank == 0:
        print(json.dumps({'variant': args.variant, 'step_ms': step_ms, 'peak_gib': peak / 2 ** 30,
                          'params_m': params / 1e6, 'swapped': swapped, 'loss_first': losses[0],
                          'loss_last': losses[-1], 'world_size': world,
                          'tokens_per_s': world * args.batch * args.seq / (step_ms / 1e3)}))
    if args.ddp:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--seq', type=int, default=128)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--dtype', default='fp32')
    ap.add_argument('--variants', default='vanilla,gelu3,rand0.2,both')
    ap.add_argument('--variant', default=None)
    ap.add_argument('--json', default=None)
    ap.add_argument('--ddp', action='store_true', help='run under torchrun with DistributedDataParallel')
    args = ap.parse_args()
    if args.ddp and not args.variant:
        args.variant = 'both'
    if args.variant:
        return child(args)
    results = []
    for variant in args.variants.split(','):
        cmd = [sys.executable, __file__, '--variant', variant, '--batch', str(args.batch), '--seq',
               str(args.seq), '--steps', str(args.steps), '--dtype', args.dtype]
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ))
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ''
        try:
            results.append(json.loads(line))
        except json.JSONDecodeError:
            results.append({'variant': variant, 'error': (r.stderr or r.stdout)[-600:]})
    vanilla = next((r for r in results if r.get('variant') == 'vanilla' and 'peak_gib' in r), None)
    print(f'RoBERTa-base, batch {args.batch} x {args.seq} tokens, {args.dtype}, random init, synthetic ids')
    print(f'{"variant":10s} {"step ms":>9s} {"peak GiB":>9s} {"vs vanilla":>11s} {"reference":>10s}')
    for r in results:
        if 'error' in r:
            print(f'{r["variant"]:10s} ERROR {r["error"]}')
            continue
        delta = 100 * (r['peak_gib'] / vanilla['peak_gib'] - 1) if vanilla else float('nan')
        r['peak_vs_vanilla_pct'] = delta
        r['reference_published_pct'] = PUBLISHED.get(r['variant'])
        ref = f'{PUBLISHED[r["variant"]]:+.1f} %' if r['variant'] in PUBLISHED else ''
        print(f'{r["variant"]:10s} {r["step_ms"]:9.1f} {r["peak_gib"]:9.2f} {delta:+10.1f} % {ref:>10s}')
    if args.json:
        Path(args.json).write_text(json.dumps({'config': vars(args), 'results': results}, indent=1))


if __name__ == '__main__':
    main()
