"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — golden vectors at the BASELINE.json shapes.

Runs the UNMODIFIED reference (``ModelCross`` / ``ModelVIT`` imported from /root/reference through
oracle/ref_loader.py) in fp64 (and fp32, as shipped) on the cases of oracle/full_cases.py and freezes into
``tests/golden/full_<case>.pt``:

    state_checksum, img_checksum   fingerprints of the seeded weights / inputs (detect generator drift)
    logits64 / loss64              reference run as ``.double()``
    logits32 / loss32              reference run in fp32 as shipped (context for the fp32-mode tolerance)
    grad_norm[k]                   ||dL/dp_k|| of the fp64 reference gradient, every parameter
    grad_sample[k]                 fp64 reference gradient at oracle.full_cases.sample_index(numel, k) (fp32 storage)

The batch is fed sample by sample (the reference has no cross-sample op: LayerNorm, per-sample attention, mean
CE), logits concatenated, loss and gradients averaged — identical in real arithmetic to one batched call and it
bounds the memory of the materialised [B, H, N, N] score tensors (cfg5: 1.07 GB per layer-stream and sample).

Usage (build container only; minutes per case on 8 cores):  python -m oracle.gen_golden_full [case ...]
"""
from __future__ import annotations

import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from oracle.full_cases import FULL_CASES, build_full_case, sample_index  # noqa: E402
from oracle.functional import make_config  # noqa: E402
from oracle.weights import state_checksum  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _ref_classes():
    mc, mv = ref_loader.load("model_cross"), ref_loader.load("modelv3")
    return (lambda cfg: mc.ModelCross(ref_loader.to_config_dict(cfg)),
            lambda cfg: mv.ModelVIT(ref_loader.to_config_dict(cfg)))


def run(model, img, labels, dtype):
    model = model.to(dtype).train()   # dropout = 0.0: identity in train mode
    for p in model.parameters():
        p.grad = None
    B = img.shape[0]
    logits, loss = [], 0.0
    for b in range(B):
        lg, ls = model(img[b:b + 1].to(dtype), labels[b:b + 1])
        (ls / B).backward()
        logits.append(lg.detach())
        loss = loss + ls.detach() / B
    return torch.cat(logits), loss, {k: p.grad.detach() for k, p in model.named_parameters()}


def main(names):
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cross, vit = _ref_classes()
    for name in names:
        t0 = time.time()
        kind, cfg, model, img, labels = build_full_case(name, cross, vit, make_config)
        state = {k: v.detach().clone() for k, v in model.state_dict().items()}
        rec = {"case": name, "torch": torch.__version__, "state_checksum": state_checksum(state),
               "img_checksum": float(img.double().sum()), "labels": labels.clone()}
        l32, s32, _ = run(model, img, labels, torch.float32)
        rec["logits32"], rec["loss32"] = l32.clone(), s32.clone()
        model.load_state_dict(state)
        l64, s64, g64 = run(model, img, labels, torch.float64)
        rec["logits64"], rec["loss64"] = l64.clone(), s64.clone()
        rec["grad_norm"] = {k: float(g.norm()) for k, g in g64.items()}
        rec["grad_sample"] = {k: g.flatten()[sample_index(g.numel(), i)].to(torch.float32).clone()
                              for i, (k, g) in enumerate(g64.items())}
        path = os.path.join(OUT, "full_" + name + ".pt")
        torch.save(rec, path)
        print(f"{name}: loss64={float(s64):.12f} loss32={float(s32):.8f} |logits|={float(l64.norm()):.6f} "
              f"fp32-vs-fp64 logits rel {float((l32.double() - l64).norm() / l64.norm()):.2e} -> {path} "
              f"({os.path.getsize(path)} B, {time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(FULL_CASES))
