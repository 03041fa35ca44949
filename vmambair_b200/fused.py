"""Fused inference pipeline of one OSS block (hand-written kernels through the C-ABI).
`available()` gates on what has been built; there is no CPU path."""
from __future__ import annotations

import torch


def available(block, x: torch.Tensor) -> bool:
    return False


def block_forward(block, x: torch.Tensor) -> torch.Tensor:
    raise RuntimeError("vmambair_b200.fused: fused OSS block pipeline not built")
