"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — restatement of the reference's per-batch metrics.

``log_stats`` (/root/reference/model_cross.py:243-255) = ``compute_metrics(argmax(logits, 1), labels)``
(/root/reference/utils.py:18-62) + ``torchmetrics.functional.auroc(softmax(logits, 1)[:, 1], labels, task="binary")``.
torchmetrics is absent from the image and unpinned by the reference (PARITY UNPINNED against it); the definitions below
are its documented binary ones (0 / 0 = 0 through ``_safe_divide``; exact-mode AUROC = trapezoid over the distinct
thresholds, 0 when a class is missing) and are pinned against scikit-learn 1.9 (present) in tests/test_metrics_cpu.py.
Lightning's ``on_epoch=True`` reduction: batch-size-weighted mean of the per-batch values, then the mean over ranks.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch

NAMES = ("acc", "prec", "rec", "spec", "f1", "npv", "auc_roc", "loss")


def _safe(a: float, b: float) -> float:
    return float(np.float32(a) / np.float32(b)) if b != 0 else 0.0


def batch_metrics(logits: torch.Tensor, labels: torch.Tensor) -> Dict[str, float]:
    pred = torch.argmax(logits, dim=1)
    y = (labels != 0).long()
    tp = int(((pred == 1) & (y == 1)).sum())
    tn = int(((pred == 0) & (y == 0)).sum())
    fp = int(((pred == 1) & (y == 0)).sum())
    fn = int(((pred == 0) & (y == 1)).sum())
    out = {"acc": _safe(tp + tn, tp + tn + fp + fn), "prec": _safe(tp, tp + fp), "rec": _safe(tp, tp + fn),
           "spec": _safe(tn, tn + fp), "f1": _safe(2 * tp, 2 * tp + fn + fp), "npv": _safe(tn, tn + fn)}
    prob = torch.softmax(logits.float(), dim=1)[:, 1].numpy()
    out["auc_roc"] = auroc(prob, y.numpy())
    return out


def auroc(prob: np.ndarray, y: np.ndarray) -> float:
    """Exact binary ROC AUC: trapezoid over the ROC points at the distinct score thresholds."""
    npos, nneg = int(y.sum()), int((1 - y).sum())
    if npos == 0 or nneg == 0:
        return 0.0
    order = np.argsort(-prob, kind="stable")
    p, t = prob[order], y[order]
    last = np.r_[np.nonzero(np.diff(p))[0], len(p) - 1]        # last index of every run of equal scores
    tps = np.r_[0, np.cumsum(t)[last]].astype(np.float64)
    fps = np.r_[0, np.cumsum(1 - t)[last]].astype(np.float64)
    return float(np.trapezoid(tps / npos, fps / nneg))


def epoch_metrics(batches: Sequence[Tuple[torch.Tensor, torch.Tensor, float]], prefix: str = "train") -> Dict[str, float]:
    """``batches``: (logits, labels, loss) of one rank's epoch -> batch-size-weighted means under the reference's keys."""
    tot = {n: 0.0 for n in NAMES}
    w = 0.0
    for logits, labels, loss in batches:
        m = batch_metrics(logits, labels)
        m["loss"] = float(loss)
        for n in NAMES:
            tot[n] += logits.shape[0] * m[n]
        w += logits.shape[0]
    return {f"{prefix}_{n}": tot[n] / w for n in NAMES}
