"""Cost of the per-step metrics launch (tools; GPU): cavit_batch_metrics against the reference's formulation of log_stats
(six confusion-matrix quotients with .item() each + a sort-based AUROC) written with plain torch ops on the same device."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))
import torch  # noqa: E402

from cavit.metrics import EpochMetrics  # noqa: E402


def eager_log_stats(logits, labels):
    """The reference's per-step work in spirit: argmax, counts, seven host read-backs (torchmetrics is not installed)."""
    pred = logits.argmax(1)
    tp = ((pred == 1) & (labels == 1)).sum().float()
    tn = ((pred == 0) & (labels == 0)).sum().float()
    fp = ((pred == 1) & (labels == 0)).sum().float()
    fn = ((pred == 0) & (labels == 1)).sum().float()
    vals = [((tp + tn) / (tp + tn + fp + fn)).item(), (tp / (tp + fp).clamp_min(1)).item(), (tp / (tp + fn).clamp_min(1)).item(),
            (tn / (tn + fp).clamp_min(1)).item(), (2 * tp / (2 * tp + fp + fn).clamp_min(1)).item(),
            (tn / (tn + fn).clamp_min(1)).item()]
    prob = torch.softmax(logits, 1)[:, 1]
    order = torch.argsort(prob, descending=True)
    y = labels[order].float()
    tps, fps = torch.cumsum(y, 0), torch.cumsum(1 - y, 0)
    vals.append(torch.trapz(tps / tps[-1].clamp_min(1), fps / fps[-1].clamp_min(1)).item())
    return vals


for B in (8, 256, 1024, 8192):
    g = torch.Generator().manual_seed(B)
    logits = torch.randn(B, 2, generator=g).cuda()
    labels = torch.randint(0, 2, (B,), generator=g).cuda()
    loss = torch.tensor(0.5, device="cuda")
    em = EpochMetrics("cuda:0")
    for _ in range(3):
        em.update(logits, labels, loss)
        eager_log_stats(logits, labels)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        em.update(logits, labels, loss)
    e1.record()
    torch.cuda.synchronize()
    dev_us = e0.elapsed_time(e1) / 50 * 1e3
    t0 = time.perf_counter()
    for _ in range(50):
        em.update(logits, labels, loss)
    host_us = (time.perf_counter() - t0) / 50 * 1e6
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        eager_log_stats(logits, labels)
    eager_us = (time.perf_counter() - t0) / 20 * 1e6
    print(f"B = {B}: cavit_batch_metrics {dev_us:.1f} us on the device, {host_us:.1f} us of host time per update (no sync); "
          f"eager torch formulation with 7 .item() read-backs {eager_us:.0f} us of host time per step, each a pipeline drain")
