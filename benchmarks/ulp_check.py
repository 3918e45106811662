"""Worst error in ulps of the fp32 ELU family, logsigmoid and softplus against float64 (and of ATen's
own fp32 kernels, for scale) on ~6 M points incl. special values."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch.nn.functional as F
import fewbit_b200 as fb
torch.manual_seed(0)
dev = 'cuda'
x = torch.cat([torch.randn(1 << 22, device=dev) * 2, torch.randn(1 << 20, device=dev) * 20, torch.linspace(-100, 100, 1 << 20, device=dev),
               torch.tensor([0.0, -0.0, 1e-30, -1e-30, 1e-8, -1e-8, float('inf'), -float('inf'), float('nan'), 88.0, -88.0, -87.3, -103.0, -104.0, 20.0, 20.000002, 80.0, -80.0, -79.9, -80.1, -1e4, 1e4], device=dev)])
names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['elu', 'celu', 'selu', 'logsigmoid', 'softplus', 'sigmoid', 'silu', 'mish']
for name in names:
    y = getattr(fb.functional, name)(x.clone())
    ref64 = getattr(F, name)(x.double())
    ref32 = getattr(F, name)(x)
    tiny = torch.isfinite(ref64) & (ref64.abs() < 1e-30)         # results near the bottom of the exponent range
    fin = torch.isfinite(ref64) & ~tiny
    sp = (torch.nextafter(ref64.float().abs(), torch.tensor(float('inf'), device=dev)) - ref64.float().abs()).double()
    ours = ((y.double() - ref64).abs() / sp)[fin]
    aten = ((ref32.double() - ref64).abs() / sp)[fin]
    same_nan = bool((torch.isnan(y) == torch.isnan(ref32)).all())
    below = (y.double() - ref64).abs()[tiny].max().item() if tiny.any() else 0.0
    print(f'{name:11s} ours max {ours.max().item():.2f} ulp  (ATen fp32 max {aten.max().item():.2f} ulp), vs ATen max {(((y - ref32).abs().double() / sp)[fin]).max().item():.2f} ulp, nan agree {same_nan}, inf agree {bool((torch.isinf(y) == torch.isinf(ref32)).all())}, abs error where |f| < 1e-30: {below:.1e}')
if 'softplus' not in names:
    sys.exit(0)
y = fb.functional.softplus(x.clone(), beta=2.5, threshold=7.0)
ref64 = F.softplus(x.double(), beta=2.5, threshold=7.0)
fin = torch.isfinite(ref64)
sp = (torch.nextafter(ref64.float().abs(), torch.tensor(float('inf'), device=dev)) - ref64.float().abs()).double()
print('softplus beta 2.5: max', (((y.double() - ref64).abs() / sp)[fin]).max().item(), 'ulp')
