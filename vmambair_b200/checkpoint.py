"""Checkpoint compatibility with the reference's BasicSR files.

The reference stores networks as `{param_key: state_dict}` with `param_key` in {"params", "params_ema"}, keys possibly
prefixed with DataParallel's "module." (save_network / load_network, Deraining/basicsr/models/base_model.py:213-244,
281-309; the SR / RealSR trees use pip basicsr's identical functions).  The drop-in nets of `vmambair_b200.archs` keep the
reference's parameter names and shapes, so a published `.pth` loads with strict=True; these helpers restate the
file-level conventions (which key to read, the "params_ema -> params" fall-back, "module." stripping, size-mismatch
handling for strict=False) without depending on basicsr.
"""
from __future__ import annotations

import logging
from typing import Dict, Optional

import torch

log = logging.getLogger("vmambair_b200")


def _bare(net: torch.nn.Module) -> torch.nn.Module:
    """the wrapped module of DataParallel / DistributedDataParallel (get_bare_model, base_model.py:86-91)"""
    return net.module if isinstance(net, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else net


def strip_module_prefix(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}


def select_params(ckpt: dict, param_key: Optional[str] = "params") -> Dict[str, torch.Tensor]:
    """the state_dict inside a loaded checkpoint file: ckpt[param_key], falling back from a missing "params_ema" to
    "params" as the reference does (base_model.py:296-300); param_key=None means the file is a bare state_dict"""
    if param_key is None:
        return ckpt
    if param_key not in ckpt and "params" in ckpt:
        log.info("checkpoint has no %r, using 'params'", param_key)
        param_key = "params"
    if param_key not in ckpt:
        raise KeyError(f"checkpoint has no {param_key!r} entry (keys: {sorted(ckpt)[:8]})")
    return ckpt[param_key]


def load_network(net: torch.nn.Module, path: str, strict: bool = True, param_key: Optional[str] = "params") -> torch.nn.Module:
    """Load a reference `.pth` into `net` (values land in the parameters' current dtype / device)."""
    net = _bare(net)
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    state = strip_module_prefix(select_params(ckpt, param_key))
    if not strict:
        # same key, different size: reported and skipped, as the reference does by renaming the entry to "<key>.ignore"
        own = net.state_dict()
        for k in [k for k in state if k in own and own[k].shape != state[k].shape]:
            log.warning("size differs, ignored [%s]: net %s, file %s", k, tuple(own[k].shape), tuple(state[k].shape))
            state[k + ".ignore"] = state.pop(k)
    missing = set(net.state_dict()) - set(state)
    extra = set(state) - set(net.state_dict())
    if missing or extra:
        log.warning("keys only in the net: %s; only in the file: %s", sorted(missing)[:8], sorted(extra)[:8])
    net.load_state_dict(state, strict=strict)
    return net


def save_network(net: torch.nn.Module, path: str, param_key: str = "params", ema: Optional[torch.nn.Module] = None) -> None:
    """Write `{param_key: state_dict}` (CPU tensors, no "module." prefix); with `ema` also "params_ema", which is the
    layout of the reference's released files (save_network([net_g, net_g_ema], ..., param_key=["params", "params_ema"]))."""
    out = {param_key: {k: v.detach().cpu() for k, v in strip_module_prefix(_bare(net).state_dict()).items()}}
    if ema is not None:
        out["params_ema"] = {k: v.detach().cpu() for k, v in strip_module_prefix(_bare(ema).state_dict()).items()}
    torch.save(out, path)
