"""How much of a forward kernel's time is the last, partly filled round of its persistent grid?
Times one (function, dtype, bits) cell at the RoBERTa activation size (128 x 128 x 3072 = 49152 tiles of 1024
elements over 148 SMs x 4 CTAs x 8 warps = 4736 warps: 10.38 tiles per warp) and at sizes that divide evenly
(10 and 11 tiles per warp).  12 back-to-back launches on four buffer sets replayed from a CUDA graph, median of 5."""
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from fewbit_b200 import native  # noqa: E402
from fewbit_b200.functional import store  # noqa: E402

dev = torch.device('cuda:0')
peak = 6548.5
cells = sys.argv[1].split(',') if len(sys.argv) > 1 else ['gelu:bf16:7', 'gelu:bf16:3', 'gelu:bf16:8', 'hardswish:bf16:7']
warps = torch.cuda.get_device_properties(0).multi_processor_count * 4 * 8
sizes = [('128x128x3072', 128 * 128 * 3072), ('10 tiles/warp', warps * 10 * 1024), ('11 tiles/warp', warps * 11 * 1024),
         ('10.5 tiles/warp', warps * 10 * 1024 + warps * 512)]
for cell in cells:
    name, tag, bits = cell.split(':')
    bits = int(bits)
    dtype = torch.bfloat16 if tag == 'bf16' else torch.float32
    es = 2 if tag == 'bf16' else 4
    borders, _ = store.get(name, bits, dev, dtype)
    bounds = borders[1:-1].contiguous()
    line = []
    for label, n in sizes:
        xs = [(torch.randn(n, device=dev) * 2).to(dtype) for _ in range(4)]
        ys = [torch.empty_like(t) for t in xs]
        states = [native.new_state(xs[0], bits) for _ in range(4)]
        k = [0]

        def fwd():
            i = k[0] % 4
            k[0] += 1
            native.stepwise_forward(name, xs[i], ys[i], states[i], bits, bounds)
        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fwd()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(12):
                    fwd()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 12)
        ms = statistics.median(ts)
        gbs = n * (2 * es + bits / 8) / (ms / 1e3) / 1e9
        line.append(f'{label}: {ms * 1e3:6.1f} us {gbs:5.0f} GB/s ({100 * gbs / peak:.0f} %)')
        del xs, ys, states
    print(f'{cell:20s} ' + '   '.join(line), flush=True)
