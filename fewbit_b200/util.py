"""Module-tree utilities (reference ``fewbit/util.py``): ``map_module`` / ``convert_linear``
with identical traversal semantics, plus the saved-tensor memory accounting helpers."""
from __future__ import annotations

import re
from contextlib import contextmanager
from dataclasses import dataclass
from functools import partial
from typing import Any, Callable, Optional

import torch as T

__all__ = ('HookedMemoryUsage', 'convert_linear', 'estimate_memory_usage', 'map_module',
           'memory_usage_hooks', 'teniter', 'traverse')


def map_module(root: T.nn.Module, func: Callable[[T.nn.Module, str], T.nn.Module],
               patt: Optional[str] = None) -> T.nn.Module:
    """Apply ``func(module, path)`` to every module of the tree, children first.

    Semantics of reference ``util.py:147-187``: post-order walk over ``named_children``;
    paths look like ``/encoder/layer/0/output/dense`` and the root is ``/``; a module is
    visited iff ``re.match(patt, path)`` (default: everything); a child is replaced in its
    parent when ``func`` returns a different object; the (possibly replaced) root is returned.
    """
    regex = re.compile(patt or r'.*')

    def visit(node: T.nn.Module, path: str) -> T.nn.Module:
        for name, child in node.named_children():
            mapped = visit(child, f'{path}/{name}')
            if mapped is not child:
                setattr(node, name, mapped)
        if regex.match(path or '/'):
            node = func(node, path or '/')
            if not isinstance(node, T.nn.Module):
                raise ValueError('Mapped result should be toch.nn.Module type.')
        return node

    return visit(root, '')


def convert_linear(module: T.nn.Module, ctor, **kwargs) -> T.nn.Module:
    """``nn.Linear`` -> ``ctor(in_features, out_features, bias, device, dtype, **kwargs)``
    sharing the original parameters; any other module is returned untouched
    (reference ``util.py:190-208``)."""
    if not isinstance(module, T.nn.Linear):
        return module
    layer = ctor(in_features=module.in_features, out_features=module.out_features,
                 bias=module.bias is not None, device=module.weight.device,
                 dtype=module.weight.dtype, **kwargs)
    layer.weight = T.nn.Parameter(module.weight)
    if layer.bias is not None:
        layer.bias = T.nn.Parameter(module.bias)
    return layer


# ------------------------------------------------------- saved-tensor accounting ----
# Diagnostics of the reference (util.py:20-144): how many bytes does the autograd graph keep
# alive?  Used by the benchmarks to show what the packed state saves.

def traverse(variable: T.Tensor, callback: Callable[[Any, T.Tensor, bool], Any]):
    """Walk the backward graph from ``variable`` and report every reachable tensor as
    ``callback(node, tensor, is_saved)``."""
    seen = set()
    stack = [variable.grad_fn]
    while stack:
        node = stack.pop()
        if node is None or node in seen:
            continue
        seen.add(node)
        if hasattr(node, 'saved_tensors'):  # Python autograd.Function
            for ten in node.saved_tensors:
                callback(node, ten, True)
        if hasattr(node, 'variable'):  # AccumulateGrad: the leaf itself
            callback(node, node.variable.data, False)
        for attr in dir(node):  # built-in nodes expose _saved_<name>
            if attr.startswith('_saved_'):
                try:
                    val = getattr(node, attr)
                except RuntimeError:
                    continue
                if T.is_tensor(val):
                    callback(node, val, True)
                elif isinstance(val, (tuple, list)):
                    for ten in val:
                        if T.is_tensor(ten):
                            callback(node, ten, True)
        stack.extend(child for child, _ in getattr(node, 'next_functions', ()))


def teniter(variable: T.Tensor, include_ordinary=True, include_saved=False):
    """Unique tensors reachable from ``variable``: leaves ("ordinary") and/or saved ones."""
    state = {}

    def note(_, ten, saved):
        _, ordinary, was_saved = state.get(id(ten), (ten, False, False))
        state[id(ten)] = (ten, ordinary or not saved, was_saved or saved)

    traverse(variable, note)
    for ten, ordinary, saved in state.values():
        if (include_ordinary and ordinary) or (include_saved and saved):
            yield ten


def estimate_memory_usage(variable: T.Tensor, saved_only=False) -> int:
    flags = (False, True) if saved_only else (True, False)
    return sum(ten.numel() * ten.element_size() for ten in teniter(variable, *flags))


@dataclass
class HookedMemoryUsage:
    forward: Optional[int] = None
    backward: Optional[int] = None

    @property
    def value(self) -> Optional[int]:
        return self.backward or self.forward


@contextmanager
def memory_usage_hooks():
    """Count bytes packed for / unpacked in backward via ``saved_tensors_hooks``."""
    usage = HookedMemoryUsage()

    def pack(ten):
        usage.forward = (usage.forward or 0) + ten.numel() * ten.element_size()
        return ten

    def unpack(ten):
        usage.backward = (usage.backward or 0) + ten.numel() * ten.element_size()
        return ten

    with T.autograd.graph.saved_tensors_hooks(pack, unpack):
        yield usage
