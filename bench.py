#!/usr/bin/env python
"""bench.py -- FewBit hot path on B200: 1-bit mask pack + unpack on a 1 GiB bf16 tensor.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): per GPU one synthetic bf16 activation tensor of 2^29
elements (1 GiB) and a gradient of the same size.  One *step* = for each of ReLU,
LeakyReLU(0.01), Hardtanh(-1, 1): forward (y = f(x), 1-bit mask packed into 64 MiB) and
backward (gin = factor(mask) * gout) -- six kernel launches through the C ABI
(include/fewbit_b200.h).  Metric: algorithmic GB/s, bytes = n * (2 + 2 + 1/8) per pass
(SURVEY 8d), summed over all GPUs (weak scaling: every rank owns its own shard, no collective
on the data path).  Inputs are 8x larger than L2, so no flush is needed between iterations.

One JSON line on stdout (rank 0); see the contract in the task description for the keys.
Beyond the contract the line carries `rooflines` -- one entry per kernel of the path (the six
mask kernels from events inside the timed region; 3-bit GELU forward / backward in fp32 and bf16
on 128 x 128 x 3072 and the tcgen05 projection kernel from side runs, N = 1 only) -- and
`extra.roberta`: RoBERTa-base 128 x 128 step time and peak memory for {vanilla, 3-bit GELU,
RandomizedLinear 0.2, both} in fp32 and bf16 (child processes, benchmarks/roberta_step.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ELEMS = 1 << 29                      # 1 GiB of bf16
FUNCS = (('relu', 0.0, 0.0), ('leaky_relu', 0.01, 0.0), ('hardtanh', -1.0, 1.0))
BYTES_PER_PASS = N_ELEMS * 2 + N_ELEMS * 2 + N_ELEMS // 8          # read + write + mask
BYTES_PER_STEP = BYTES_PER_PASS * 2 * len(FUNCS)
METRIC = 'mask_pack_unpack_throughput'
UNIT = 'GB/s'
CONFIG = {
    'workload': '1-bit ReLU/LeakyReLU/Hardtanh mask pack+unpack, 1 GiB bf16 tensor per GPU '
                '(BASELINE.json configs[1])',
    'elements_per_gpu': N_ELEMS,
    'functions': [f[0] for f in FUNCS],
    'passes_per_step': 2 * len(FUNCS),
    'algorithmic_bytes_per_step_per_gpu': BYTES_PER_STEP,
    'l2_policy': 'inputs (2 GiB read per pass pair) are larger than the 126 MB L2; no flush needed',
    'parallelism': 'batch-sharded replicas, no collective on the data path',
    'reference_arm': 'the CPU path runs on rank 0 only, with every host core the process may use, on a '
                     'sample sized for ~90 s (a rate: GB/s of the same three functions, fwd+bwd); it does '
                     'not depend on --gpus',
}


def measured_peak():
    path = ROOT / 'MEASURED_PEAKS.json'
    if path.exists():
        try:
            return float(json.loads(path.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------- clocks ----

class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is busy (B200_PROFILING.md recipe)."""
    QUERY = ('clocks.sm,clocks.max.sm,utilization.gpu,power.draw,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                 '-lms', '50', '-i', str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, reasons, total = [], [], set(), 0
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for stamp, line in self.rows:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            total += 1
            try:
                busy = float(parts[2]) >= 50.0
            except ValueError:
                busy = True
            if not (t0 <= stamp <= t1 + 0.05) or not busy:
                continue
            try:
                sm.append(float(parts[0]))
                sm_max.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None,
                'sm_max_mhz': max(sm_max) if sm_max else None,
                'reasons': sorted(reasons), 'samples_under_load': len(sm), 'samples': total}


# ------------------------------------------------------------- CPU reference arm ----

def cpu_reference_step(x, g, use_ref_codec):
    """The reference's CPU implementation of one step on a sample: torch CPU ops for the value /
    predicate / multiply (what fewbit/cpu/gelu.cc does with ATen) + the reference's own
    single-threaded bit-stream codec fewbit::Deflate / Inflate (fewbit/cpu/codec.h) compiled
    into oracle/_ref (the CPU *op* cannot run 1-bit tables, App. C-4).  Without
    oracle/_ref the scalar C port (oracle/) stands in."""
    import torch
    import torch.nn.functional as F

    import oracle
    n = x.numel()
    for name, p0, p1 in FUNCS:
        if use_ref_codec:
            if name == 'relu':
                y, pred = F.relu(x), x > 0
            elif name == 'leaky_relu':
                y, pred = F.leaky_relu(x, p0), x < 0
            else:
                y, pred = F.hardtanh(x, p0, p1), (x > p0) & (x < p1)
            state = oracle.ref_deflate(pred.to(torch.int32).numpy(), 1)
            codes = torch.from_numpy(oracle.ref_inflate(state, n, 1))
            if name == 'leaky_relu':
                gin = torch.where(codes != 0, g * p0, g)
            else:
                gin = codes.to(g.dtype) * g
            del y, gin
        else:
            bits = x.view(torch.int16).numpy().view('uint16')
            gbits = g.view(torch.int16).numpy().view('uint16')
            _, state = oracle.piecewise_forward(name, bits, p0, p1)
            oracle.piecewise_backward(name, state, gbits, p0)


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference CPU path on the host cores, rank 0 only."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is one process that may use the
    # host exactly as it does at --gpus 1 (torch's default thread count), whatever --gpus says
    if world > 1:
        os.environ.pop('OMP_NUM_THREADS', None)
    import torch

    import oracle
    use_ref = oracle.ref_codec() is not None
    torch.manual_seed(0)
    probe = 1 << 20
    x = (torch.randn(probe) * 2).to(torch.bfloat16)
    g = torch.randn(probe).to(torch.bfloat16)
    cpu_reference_step(x, g, use_ref)
    t = time.perf_counter()
    cpu_reference_step(x, g, use_ref)
    per_elem = (time.perf_counter() - t) / probe
    budget = 90.0 / max(1, args.steps + args.warmup)          # whole run within ~1.5 minutes
    sample = int(min(1 << 26, max(1 << 20, budget / per_elem)))
    sample -= sample % 2048
    x = (torch.randn(sample) * 2).to(torch.bfloat16)
    g = torch.randn(sample).to(torch.bfloat16)
    for _ in range(args.warmup):
        cpu_reference_step(x, g, use_ref)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(x, g, use_ref)
    elapsed = time.perf_counter() - t0
    nbytes = (sample * 4 + sample // 8) * 2 * len(FUNCS)
    value = nbytes * args.steps / elapsed / 1e9
    cores = torch.get_num_threads()
    kind = 'reference' if use_ref else 'port'
    sample_txt = (f'{sample} bf16 elements per step ({sample * 2 / 2**20:.0f} MiB; the full workload '
                  f'is {N_ELEMS}), same three functions, fwd+bwd')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': elapsed / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
        'data': 'synthetic', 'config': dict(CONFIG, reference_sample_elements=sample, reference_threads=cores, n_gpus_independent=True),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': sample_txt},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def cpu_baseline_leg():
    """Bounded (~10-30 s) timing of the reference CPU path inside our arm (rank 0, N=1)."""
    import torch

    import oracle
    use_ref = oracle.ref_codec() is not None
    sample = 1 << 25
    torch.manual_seed(0)
    x = (torch.randn(sample) * 2).to(torch.bfloat16)
    g = torch.randn(sample).to(torch.bfloat16)
    cpu_reference_step(x, g, use_ref)
    reps, t0 = 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t0 < 10.0 and reps < 50):
        cpu_reference_step(x, g, use_ref)
        reps += 1
    elapsed = time.perf_counter() - t0
    nbytes = (sample * 4 + sample // 8) * 2 * len(FUNCS)
    return {'value': nbytes * reps / elapsed / 1e9, 'unit': UNIT, 'cores': torch.get_num_threads(),
            'kind': 'reference' if use_ref else 'port',
            'sample': f'{reps} steps over {sample} bf16 elements ({sample * 2 >> 20} MiB of the 1 GiB '
                      f'workload), {elapsed:.1f} s; torch CPU ops + fewbit::Deflate/Inflate '
                      f'(fewbit/cpu/codec.h, single-threaded by construction)'}


def load_traffic():
    """DRAM bytes per launch from the committed ncu capture (tools/ncu_table.py --traffic)."""
    path = ROOT / 'profiles' / 'ncu_traffic.json'
    try:
        return json.loads(path.read_text()) if path.exists() else {}
    except Exception:  # noqa: BLE001
        return {}


def traffic_name(kernel):
    # every 1-bit backward is the same kernel (MaskFactorOp) with other constants
    return 'relu_backward' if kernel.endswith('_backward') and kernel.split('_')[0] in ('relu', 'leaky', 'hardtanh') \
        else kernel


def roofline_entry(kernel, nbytes, ms, peak, traffic, where, bound='hbm', unit='GB/s', scale=1e9, **more):
    achieved = nbytes / (ms / 1e3) / scale
    entry = {'kernel': kernel, 'bound': bound, 'achieved': achieved, 'peak': peak, 'unit': unit,
             'frac': achieved / peak, 'traffic': traffic, 'avg_launch_ms': ms, 'measured': where}
    entry['algorithmic_bytes_per_launch' if bound == 'hbm' else 'flops_per_launch'] = nbytes
    entry.update(more)
    return entry


def cross_device_check(local_rank, world):
    """At N >= 2: run the operator on a tensor that lives on ANOTHER device than the current one
    and compare with torch's own bucket search -- the kernels must follow the tensor (device
    guard + that device's stream), which a 1-GPU box cannot show."""
    import torch

    import fewbit_b200 as fewbit
    try:
        other = torch.device('cuda', (local_rank + 1) % world)
        borders, levels = fewbit.functional.store.get('gelu', 3, other, torch.float32)
        bounds = borders[1:-1].contiguous()
        gen = torch.Generator(other).manual_seed(99 + local_rank)
        x = torch.randn(1 << 20, device=other, generator=gen) * 2
        leaf = x.clone().requires_grad_()
        y = torch.ops.fewbit.gelu(leaf * 1.0, bounds, levels)
        y.sum().backward()
        torch.cuda.synchronize(other)
        codes = torch.searchsorted(bounds, x, right=False)
        ok = (torch.equal(leaf.grad, levels[codes]) and y.device == other
              and torch.allclose(y, torch.nn.functional.gelu(x), atol=1e-6, rtol=1e-5)
              and torch.cuda.current_device() == local_rank)
        return 'ok' if ok else 'mismatch'
    except Exception as exc:  # noqa: BLE001
        return f'{type(exc).__name__}: {exc}'


def roberta_block():
    """RoBERTa-base (random init, synthetic 128 x 128 batch) step time and peak memory, the other half
    of BASELINE.json's metric: one child process per variant (benchmarks/roberta_step.py, the
    reference's measurement method benchmark/benchmark.py:165-188), fp32 and bf16."""
    out = {}
    for dtype in ('fp32', 'bf16'):
        dest = ROOT / 'gpurun_out' / f'bench_roberta_{dtype}.json'
        try:
            dest.parent.mkdir(exist_ok=True)
            subprocess.run([sys.executable, str(ROOT / 'benchmarks' / 'roberta_step.py'), '--dtype', dtype,
                            '--steps', '3', '--json', str(dest)], capture_output=True, text=True, timeout=240)
            rows = json.loads(dest.read_text())['results']
            out[dtype] = [{k: r.get(k) for k in ('variant', 'step_ms', 'peak_gib', 'peak_vs_vanilla_pct',
                                                  'reference_published_pct', 'error') if r.get(k) is not None}
                          for r in rows]
        except Exception as exc:  # noqa: BLE001
            out[dtype] = {'error': f'{type(exc).__name__}: {exc}'}
    out['note'] = ('batch 128 x 128 tokens, AdamW, median of 3 steps after 2 warm-ups; peak = max_memory_allocated '
                   'minus the allocation before the model is built; reference_published_pct = README.md:18-27')
    return out


# ------------------------------------------------------------------------ our arm ----

def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-GPU runs: keep this rank (and the pinned host buffers it is about to allocate and touch) on the
    CPUs that are local to its GPU's PCIe root, as a deployment would; the host-staged `e2e` pass moves
    12.9 GB per step and GPU.  Returns the CPU list used, or None where sysfs does not say."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        domain = torch.cuda.get_device_properties(local_rank).pci_domain_id
        device = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f'/sys/bus/pci/devices/{domain:04x}:{bus:02x}:{device:02x}.0/local_cpulist'
        cpus = set()
        for part in open(path).read().strip().split(','):
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f'{min(cpus)}-{max(cpus)} ({len(cpus)} cpus)'
    except (OSError, ValueError, AttributeError):
        pass
    return None


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from fewbit_b200 import native

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the product path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    host_cpus = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    torch.manual_seed(1234 + rank)
    x = torch.empty(N_ELEMS, dtype=torch.bfloat16, device=dev).normal_(0, 2)
    g = torch.empty(N_ELEMS, dtype=torch.bfloat16, device=dev).normal_()
    y, gin = torch.empty_like(x), torch.empty_like(g)
    state = native.new_state(x, 1)
    stream = torch.cuda.current_stream()

    def step(marks=None):
        for name, p0, p1 in FUNCS:
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append(e)
            native.piecewise_forward(name, x, y, state, p0, p1, stream)
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append(e)
            native.piecewise_backward(name, state, g, gin, p0, stream)
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(stream); marks.append(e)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_busy0 = time.perf_counter()
    for _ in range(max(args.warmup, 3)):
        step()
    fence()
    launches0 = native.launch_count()
    marks_all = []
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin.record(stream)
    for _ in range(args.steps):
        marks = []
        step(marks)
        marks_all.append(marks)
    end.record(stream)
    fence()
    launches = native.launch_count() - launches0
    elapsed_ms = begin.elapsed_time(end)
    t_busy1 = time.perf_counter()

    # per-kernel device times from the events recorded inside the timed region
    per_kernel = {}
    for marks in marks_all:
        for i, (name, _, _) in enumerate(FUNCS):
            per_kernel.setdefault(f'{name}_forward', []).append(marks[2 * i].elapsed_time(marks[2 * i + 1]))
            per_kernel.setdefault(f'{name}_backward', []).append(marks[2 * i + 1].elapsed_time(marks[2 * i + 2]))
    kernel_ms = {k: sum(v) / len(v) for k, v in per_kernel.items()}

    # ---- end to end through the host-buffer C ABI (pinned host memory, copies inside) ----
    e2e_steps = max(1, min(args.steps, 3))
    xh = torch.empty(N_ELEMS, dtype=torch.bfloat16).pin_memory()
    gh = torch.empty(N_ELEMS, dtype=torch.bfloat16).pin_memory()
    xh.copy_(x); gh.copy_(g)
    yh, ginh = torch.empty_like(xh).pin_memory(), torch.empty_like(gh).pin_memory()

    def e2e_step():
        for name, p0, p1 in FUNCS:
            native.piecewise_forward_host(name, xh, yh, state, p0, p1)
            native.piecewise_backward_host(name, state, gh, ginh, p0)

    e2e_step()
    fence()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_launches = None
    fence()
    t_busy2 = time.perf_counter()
    clocks = sampler.stop(t_busy0, t_busy2) if sampler else None

    # ---- max over ranks ----
    times = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = times.tolist()

    cross = cross_device_check(local_rank, world) if world > 1 else None
    if world > 1:
        flag = torch.tensor([1.0 if cross == 'ok' else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        cross = 'ok on every rank' if flag.item() == 1.0 else f'FAILED (rank {rank}: {cross})'

    extra, side_rooflines = {}, []
    if rank == 0 and world == 1:
        peak, _ = measured_peak()
        traffic_table = load_traffic()
        extra, side_rooflines = side_measurements(dev, peak, traffic_table)

    if rank == 0:
        peak, peak_src = measured_peak()
        traffic_table = load_traffic()
        value = world * BYTES_PER_STEP * args.steps / (elapsed_ms / 1e3) / 1e9
        e2e_value = world * BYTES_PER_STEP * e2e_steps / (e2e_ms / 1e3) / 1e9
        dominant = 'relu_forward'
        rooflines = [roofline_entry(k, BYTES_PER_PASS, v, peak, traffic_table.get(traffic_name(k)),
                                    where='CUDA events inside the timed region, 1 GiB bf16 tensor')
                     for k, v in kernel_ms.items()] + side_rooflines
        head = next(r for r in rooflines if r['kernel'] == dominant)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic', 'config': CONFIG,
            'roofline': {'bound': 'hbm', 'achieved': head['achieved'], 'peak': peak, 'unit': 'GB/s',
                         'frac': head['frac'], 'traffic': head['traffic'], 'kernel': dominant,
                         'peak_source': peak_src, 'algorithmic_bytes_per_launch': BYTES_PER_PASS,
                         'avg_launch_ms': kernel_ms[dominant]},
            'rooflines': rooflines,
            'kernels_GBps': {k: BYTES_PER_PASS / (v / 1e3) / 1e9 for k, v in kernel_ms.items()},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'steps': e2e_steps,
                    'h2d_bytes_per_step': 2 * len(FUNCS) * N_ELEMS * 2,
                    'd2h_bytes_per_step': 2 * len(FUNCS) * N_ELEMS * 2,
                    'api': 'fewbit_piecewise_{forward,backward}_host (pinned host buffers, chunked '
                           'H2D -> kernel -> D2H on 3 streams)'},
            'gpu_launches': launches,
            'clocks': clocks,
        }
        if cross is not None:
            line['cross_device_check'] = cross
            line['host_cpus_rank0'] = host_cpus
        if world == 1:
            line['cpu_baseline'] = cpu_baseline_leg()
        if extra:
            line['extra'] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def side_measurements(dev, peak, traffic_table):
    """Not part of the headline: the other kernels of the path, each timed alone (N = 1 only).

    3-bit GELU on 128 x 128 x 3072 (configs[0] / [2]), fp32 and bf16, forward and backward: median of
    20 launches; before every launch the L2 is flushed by READING a 1 GiB buffer, so the inputs
    come from HBM and the L2 holds clean lines (flushing by writing would leave 126 MB of dirty
    lines for the timed kernel to evict; back-to-back launches would leave its own ~50 MB of
    write-back in L2 -- the two differ by 5-10 % on a 100-200 MB tensor).
    The projection S X on tcgen05 (configs[4] shape: N = 16384 tokens, P = 3276 rows, D = 768): CUDA
    events around the C-ABI call with output and workspace preallocated, i.e. the projection kernel
    plus its split-K reduction and nothing else."""
    import torch

    from fewbit_b200 import native
    from fewbit_b200.functional import store
    out, rooflines = {}, []
    n = 128 * 128 * 3072
    flush = torch.ones(256 << 20, dtype=torch.float32, device=dev)     # 1 GiB: read before every timed launch
    sink = torch.zeros((), dtype=torch.float32, device=dev)

    GROUP = 6

    def timed(fn, reps=25, skip=5, group=GROUP):
        """Median over reps of (time of `group` launches, one per buffer set) / group.  Before every
        group the L2 is flushed by reading 1 GiB; the launches of a group touch different buffers,
        so every one of them finds its inputs in HBM.  The group is replayed from a CUDA graph:
        events around a single ~40 us launch add ~2 us of their own, and launches issued from
        Python arrive later than a 40 us kernel ends on a slow host."""
        for k in range(group):
            fn(k)
        torch.cuda.synchronize()
        side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for k in range(group):
                    fn(k)
        torch.cuda.synchronize()
        ts = []
        for it in range(reps):
            sink.copy_(flush.sum())
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            if it >= skip:
                ts.append(a.elapsed_time(b) / group)
        return statistics.median(ts)

    where = ('128x128x3072 elements; median of 20 groups of 6 launches on 6 distinct buffer sets (replayed from a '
             'CUDA graph), L2 flushed by reading 1 GiB before each group (inputs always come from HBM)')
    for tag, dtype, es in (('f32', torch.float32, 4), ('bf16', torch.bfloat16, 2)):
        borders, levels = store.get('gelu', 3, dev, dtype)
        bounds = borders[1:-1].contiguous()
        xs = [(torch.randn(n, device=dev) * 2).to(dtype) for _ in range(GROUP)]
        gs = [torch.randn(n, device=dev).to(dtype) for _ in range(GROUP)]
        ys, gins = [torch.empty_like(t) for t in xs], [torch.empty_like(t) for t in gs]
        states = [native.new_state(xs[0], 3) for _ in range(GROUP)]
        nbytes = n * (2 * es) + n * 3 // 8
        for label, key, fn in (('fwd', f'gelu3_{tag}_forward',
                                lambda k: native.stepwise_forward('gelu', xs[k], ys[k], states[k], 3, bounds)),
                               ('bwd', f'levels3_{tag}_backward',
                                lambda k: native.stepwise_backward(states[k], gs[k], gins[k], 3, levels))):
            ms = timed(fn)
            out[f'gelu3_{tag}_{label}_GBps'] = nbytes / (ms / 1e3) / 1e9
            rooflines.append(roofline_entry(f'gelu3_{tag}_{"forward" if label == "fwd" else "backward"}', nbytes, ms,
                                            peak, traffic_table.get(key), where))
        del xs, gs, ys, gins, states
    out['note'] = where
    tokens, rows, features = 16384, 3276, 768
    x = torch.randn(tokens, features, device=dev).to(torch.bfloat16)
    result = torch.empty(rows, features, dtype=torch.float32, device=dev)
    workspace = native.sketch_workspace(x, rows)
    tensor_peak = None
    peaks = ROOT / 'MEASURED_PEAKS.json'
    if peaks.exists():
        tensor_peak = json.loads(peaks.read_text()).get('bf16_tflops')
    tensor_peak = tensor_peak or 1590.0
    for kind in ('gaussian', 'rademacher'):
        calls = [0]

        def run(_):
            calls[0] += 1
            native.sketch_forward(x, rows, 1, 4 * calls[0], kind, 1.0 / rows, out=result, workspace=workspace)

        ms = timed(run)
        flops = 2.0 * rows * tokens * features
        out[f'sketch_{kind}_D768_TFLOPs'] = flops / (ms / 1e3) / 1e12
        rooflines.append(roofline_entry(
            f'sketch_{kind}_D768', flops, ms, tensor_peak, traffic_table.get('sketch_kernel'),
            'N=16384 P=3276 D=768 bf16, median of 20 groups of 6 calls of fewbit_sketch_forward with preallocated output '
            'and workspace (projection kernel + split-K reduction), L2 flushed by reading 1 GiB before each group',
            bound='tensor', unit='TFLOP/s', scale=1e12, peak_source='MEASURED_PEAKS.json bf16_tflops (burst)'))
    out['sketch_peak_bf16_TFLOPs'] = tensor_peak
    del flush
    torch.cuda.empty_cache()
    out['roberta'] = roberta_block()
    return out, rooflines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
