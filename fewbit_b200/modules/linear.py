"""``RandomizedLinear`` / ``LinearGRP`` (reference ``fewbit/modules/linear.py:39-146``)."""
from __future__ import annotations

from typing import Literal, Optional

import torch as T

from ..functional.linear import linear_crs, linear_grp

__all__ = ('LinearCRS', 'LinearGRP', 'RandomizedLinear')

MatMulType = Literal['gaussian', 'rademacher', 'dct', 'dft']


class LinearCRS(T.nn.Linear):
    """``torch.nn.Linear`` whose weight gradient is estimated by column-row sampling and which keeps
    only the sampled input columns for backward (reference ``fewbit/modules/linear.py:16-36``).

    ``proj_dim`` is the number of draws (default ``out_features // 2``).  Two slips of the
    reference are not reproduced: it passes ``proj_dim`` to ``torch.nn.Linear`` in the place of
    ``bias`` (so ``bias=False`` is ignored whenever ``proj_dim`` is given), and its ``extra_repr``
    reads an attribute that does not exist.
    """

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None,
                 proj_dim: Optional[int] = None) -> None:
        super().__init__(in_features, out_features, bias, device, dtype)
        self.proj_dim: int = proj_dim or out_features // 2

    def forward(self, input: T.Tensor) -> T.Tensor:
        return linear_crs(input, self.weight, self.bias, self.proj_dim)

    def extra_repr(self) -> str:
        return f'{super().extra_repr()}, proj_dim={self.proj_dim}'


class LinearGRP(T.nn.Linear):
    r"""``torch.nn.Linear`` that keeps a random projection of its input for backward.

    Approximates the weight gradient with Gaussian (or Rademacher) random projections along
    the batch/token axis, as in `Memory-Efficient Backpropagation through Large Linear Layers
    <https://arxiv.org/abs/2201.13195>`_; the forward pass :math:`y = xA^T + b` is exact.

    Args:
        in_features, out_features, bias, device, dtype: as :class:`torch.nn.Linear`.
        proj_dim_ratio: projection size as a fraction of the number of input rows.
        proj_dim: exact projection size (takes precedence over the ratio).
        proj_dim_min, proj_dim_max: clamp of the projection size.
        matmul: ``'gaussian'`` (default) or ``'rademacher'``.
        generator: random generator; default is the device's global generator.
        share_sketch: (not in the reference) on CUDA, layers that are called one after the other
            on the very same tensor -- query / key / value of an attention block -- share one
            sketch ``S X`` instead of taking three: the same ``S`` then serves the three weight
            gradients (each still unbiased, their noise correlated).  Default ``False``.

    Either ``proj_dim_ratio`` or ``proj_dim`` must be given.

    Example::

        >>> m = fewbit.RandomizedLinear(20, 30, proj_dim_ratio=0.5)
        >>> m(torch.randn(128, 20)).size()
        torch.Size([128, 30])
    """

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None,
                 dtype=None, proj_dim_ratio: Optional[float] = None,
                 proj_dim: Optional[int] = None, proj_dim_min: Optional[int] = None,
                 proj_dim_max: Optional[int] = None, matmul: MatMulType = 'gaussian',
                 generator: Optional[T.Generator] = None, share_sketch: bool = False) -> None:
        super().__init__(in_features, out_features, bias, device, dtype)
        self.generator = generator
        self.share_sketch = share_sketch
        self.matmul = matmul
        self.proj_dim_ratio = proj_dim_ratio
        self.proj_dim = proj_dim
        self.proj_dim_max = proj_dim_max
        self.proj_dim_min = proj_dim_min

    def forward(self, input: T.Tensor) -> T.Tensor:
        return linear_grp(input, self.weight, self.bias, self.proj_dim_ratio, self.proj_dim,
                          self.proj_dim_max, self.proj_dim_min, self.matmul, self.generator,
                          self.share_sketch)

    def extra_repr(self) -> str:
        return ', '.join([
            super().extra_repr(), f'matmul={self.matmul}', f'proj_dim={self.proj_dim}',
            f'proj_dim_ratio={self.proj_dim_ratio}', f'proj_dim_max={self.proj_dim_max}',
            f'proj_dim_min={self.proj_dim_min}'
        ])


RandomizedLinear = LinearGRP
