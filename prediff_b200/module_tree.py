"""Builds an nn.Module tree whose state_dict() has exactly the reference's key names, from a flat spec.

The CUDA models keep their parameters in such a tree so that `load_state_dict(torch.load("pretrained_*.pt"))`,
`.parameters()`, `.eval()` and `.to(device)` behave as they do on the reference modules
(surfaces 2 and 3); the tensors are pushed to the C++ side and repacked on first use.
"""
from typing import Dict, Iterable, Tuple

import torch
from torch import nn


class _Node(nn.Module):
    """Pure container (never called)."""


def build_param_tree(root: nn.Module, spec: Iterable[Tuple[str, Tuple[int, ...]]],
                     int_buffers: Dict[str, torch.Tensor] = None):
    """Registers zero-initialised fp32 Parameters under dotted names on `root` (creating container nodes), and
    optional int64 buffers (e.g. the derived `relative_position_index`)."""

    def descend(path):
        node = root
        for part in path:
            child = node._modules.get(part)
            if child is None:
                child = _Node()
                node.add_module(part, child)
            node = child
        return node

    for name, shape in spec:
        parts = name.split(".")
        descend(parts[:-1]).register_parameter(parts[-1], nn.Parameter(torch.zeros(*shape), requires_grad=False))
    for name, buf in (int_buffers or {}).items():
        parts = name.split(".")
        descend(parts[:-1]).register_buffer(parts[-1], buf.clone())
    return root
