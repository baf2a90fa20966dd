"""Checkpoint compatibility with the reference's Lightning runs.

``ModelCheckpoint`` (/root/reference/main_mist.py:174-180) writes ``torch.save`` dictionaries whose ``"state_dict"`` entry
is the LightningModule's ``state_dict()``; the drop-in modules keep the reference's parameter names and shapes
(tests/test_modules_cpu.py), so such a file loads directly, and a file written here loads into the reference model.
"""
from __future__ import annotations

from typing import Any, Dict

import torch

from . import _abi


def load_reference_checkpoint(model: torch.nn.Module, path: str, strict: bool = True) -> Dict[str, Any]:
    """Load a Lightning ``.ckpt`` (or a bare ``state_dict`` file) of the reference model into the drop-in ``model``.
    Returns the rest of the checkpoint dictionary (epoch, optimizer states, ...)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(ckpt, dict):
        raise _abi.CavitError(f"{path}: not a checkpoint dictionary")
    sd = ckpt.get("state_dict", ckpt)
    model.load_state_dict(sd, strict=strict)
    return {k: v for k, v in ckpt.items() if k != "state_dict"} if "state_dict" in ckpt else {}


def save_reference_checkpoint(model: torch.nn.Module, path: str, **extra) -> None:
    """Write ``{"state_dict": ..., **extra}`` with CPU tensors — the layout ``LightningModule.load_from_checkpoint`` and
    ``load_reference_checkpoint`` read."""
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    torch.save({"state_dict": sd, **extra}, path)
