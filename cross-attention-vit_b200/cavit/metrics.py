"""Epoch metrics without per-step host synchronisation.

The reference's ``log_stats`` (/root/reference/model_cross.py:243-255) runs after every training and validation step:
``compute_metrics`` (/root/reference/utils.py:18-62) builds six torchmetrics objects and reads six results back with
``.item()``, ``torchmetrics.functional.auroc`` adds a sort, and ``self.log(..., on_epoch=True, sync_dist=True)`` makes
Lightning average the per-batch values over the epoch (weighted by batch size) and then over the ranks. ``EpochMetrics``
keeps that definition — the epoch value of each metric is the batch-size-weighted mean of its PER-BATCH values, averaged
over ranks — but a step costs one single-block kernel launch on the step's own stream (``cavit_batch_metrics``) and an
epoch one 8-double all-reduce and one device-to-host copy.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _abi, ops

NAMES = ("acc", "prec", "rec", "spec", "f1", "npv", "auc_roc", "loss")   # suffixes of the reference's log keys


class EpochMetrics:
    def __init__(self, device, prefix: str = "train"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.CavitError("EpochMetrics needs a CUDA device (there is no CPU path)")
        _abi.require_device(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.prefix = prefix
        self.accum = torch.zeros(16, dtype=torch.float64, device=self.device)   # 10 values + scratch (include/cavit.h)

    def reset(self) -> None:
        self.accum.zero_()

    def update(self, logits: torch.Tensor, labels: torch.Tensor, loss: Optional[torch.Tensor] = None) -> None:
        """Add one batch. No host synchronisation; safe inside the training loop right after ``model(img, labels)``."""
        if logits.dim() != 2 or logits.dtype != torch.float32 or labels.dtype != torch.int64 \
                or labels.shape != logits.shape[:1]:
            raise _abi.CavitError("EpochMetrics.update: logits fp32 [B, 2] and labels int64 [B] expected")
        if loss is not None:
            loss = loss.detach().reshape(1).to(torch.float32)
        with torch.cuda.device(self.device):
            ops.batch_metrics(logits.detach().contiguous(), labels.contiguous(), loss, self.accum, B=logits.shape[0],
                              classes=logits.shape[1])

    def compute(self, process_group=None) -> Dict[str, float]:
        """Epoch values under the reference's log keys (``train_acc`` ... ``train_auc_roc``, ``train_loss``): per-rank
        weighted mean over the batches, then the mean over ranks (Lightning's ``sync_dist=True`` reduction). One packed
        all-reduce when torch.distributed is initialised, one device-to-host copy."""
        vals = epoch_means(self.accum, process_group).cpu().tolist()
        return {f"{self.prefix}_{n}": v for n, v in zip(NAMES, vals)}


def epoch_means(accum: torch.Tensor, process_group=None) -> torch.Tensor:
    """The accumulator of ``cavit_batch_metrics`` (first 10 doubles) -> 8 epoch values: this rank's batch-size-weighted means, then
    the mean over the ranks of ``process_group`` (one all-reduce of 8 doubles; skipped without torch.distributed)."""
    import torch.distributed as dist
    means = torch.where(accum[8] > 0, accum[:8] / accum[8].clamp_min(1.0), torch.zeros_like(accum[:8]))
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(means, op=dist.ReduceOp.SUM, group=process_group)
        means = means / dist.get_world_size(group=process_group)
    return means


class TestOutputs:
    """``test_step`` / ``on_test_epoch_end`` of the reference (/root/reference/model_cross.py:294-308) copy every batch's
    logits and labels to the host as they are produced (``logits.cpu()`` blocks the host once per batch). Here the batches
    stay on the device and ``finish()`` moves them with one concatenation and one device-to-host copy each."""
    __test__ = False   # not a pytest class

    def __init__(self):
        self._logits, self._targets = [], []

    def append(self, logits: torch.Tensor, labels: torch.Tensor) -> None:
        if logits.device.type != "cuda":
            raise _abi.CavitError("TestOutputs.append: device tensors expected (there is no CPU path)")
        self._logits.append(logits.detach().clone())     # the engine reuses its logits buffer every step
        self._targets.append(labels.detach().clone())

    def finish(self):
        """-> (test_logits [N, classes], test_targets [N]) on the host, as ``on_test_epoch_end`` leaves them."""
        if not self._logits:
            raise _abi.CavitError("TestOutputs.finish: no batches")
        out = torch.cat(self._logits).cpu(), torch.cat(self._targets).cpu()
        self._logits, self._targets = [], []
        return out
