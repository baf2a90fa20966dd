"""Checkpoint compatibility with the reference's Lightning runs.

``ModelCheckpoint`` (/root/reference/main_mist.py:174-180) writes ``torch.save`` dictionaries whose ``"state_dict"`` entry
is the LightningModule's ``state_dict()``; the drop-in modules keep the reference's parameter names and shapes
(tests/test_modules_cpu.py), so such a file loads directly, and a file written here loads into the reference model.
"""
from __future__ import annotations

from typing import Any, Dict

import torch

from . import _abi


def load_reference_checkpoint(model: torch.nn.Module, path: str, strict: bool = True, trusted: bool = False) -> Dict[str, Any]:
    """Load a Lightning ``.ckpt`` (or a bare ``state_dict`` file) of the reference model into the drop-in ``model``.
    Returns the rest of the checkpoint dictionary (epoch, optimizer states, ...).

    The file is read with ``weights_only=True`` (tensors and plain containers only). Lightning checkpoints that pickle
    arbitrary objects (hyper-parameter namespaces, callbacks) need ``trusted=True``, which unpickles — i.e. can execute —
    whatever the file contains: only for files you wrote yourself."""
    try:
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as exc:
        if not trusted:
            raise _abi.CavitError(f"{path}: cannot be read with weights_only=True ({type(exc).__name__}); pass trusted=True "
                                  "to unpickle it if you trust its origin") from exc
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(ckpt, dict):
        raise _abi.CavitError(f"{path}: not a checkpoint dictionary")
    sd = ckpt.get("state_dict", ckpt)
    model.load_state_dict(sd, strict=strict)
    return {k: v for k, v in ckpt.items() if k != "state_dict"} if "state_dict" in ckpt else {}


def save_reference_checkpoint(model: torch.nn.Module, path: str, **extra) -> None:
    """Write ``{"state_dict": ..., "epoch", "global_step", "pytorch-lightning_version", **extra}`` with CPU tensors: the
    layout of a ``ModelCheckpoint`` file as far as the weights go — ``model.load_state_dict(torch.load(p)["state_dict"])`` on
    the reference model and ``load_reference_checkpoint`` read it. (Whether ``LightningModule.load_from_checkpoint`` accepts it
    unchanged could not be checked here: Lightning is not installed; the version / epoch / step keys its loader looks for are
    written.)"""
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    meta = {"epoch": 0, "global_step": 0, "pytorch-lightning_version": "2.0.0"}
    meta.update(extra)
    torch.save({"state_dict": sd, **meta}, path)
