"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — golden-vector generator.

Runs the UNMODIFIED reference modules (imported from /root/reference through
oracle/ref_loader.py) on the cases of oracle/cases.py and freezes what they return into
``tests/golden/<case>.pt``:

    state_checksum   fingerprint of the seeded weights (detects generator drift)
    logits64/loss64  reference run as ``.double()``         (tight oracle pin)
    logits32/loss32  reference run in fp32 as shipped        (tolerance context)
    grad_probes      per-parameter [norm, <g,p1>, <g,p2>] of the fp64 reference gradient
    grad_full        full fp64 gradients of the small shared tensors (cls, pos, heads)

Usage (build container only):  python -m oracle.gen_golden
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_loader  # noqa: E402
from oracle.cases import CASES, build_case, grad_probes  # noqa: E402
from oracle.weights import state_checksum  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FULL = ("cls_token", "pos_embedding", "patch_to_embedding.bias")


def run_reference(kind, cfg, state, img, labels, dtype):
    mod = ref_loader.load("model_cross" if kind == "cross" else "modelv3")
    cls = mod.ModelCross if kind == "cross" else mod.ModelVIT
    model = cls(ref_loader.to_config_dict(cfg))
    missing = model.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(dtype).train()  # dropout=0.0 => identity in train mode
    logits, loss = model(img.to(dtype), labels)
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in model.named_parameters()}
    return logits.detach(), loss.detach(), grads


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name in CASES:
        kind, cfg, state, img, labels = build_case(name)
        l64, s64, g64 = run_reference(kind, cfg, state, img, labels, torch.float64)
        l32, s32, _ = run_reference(kind, cfg, state, img, labels, torch.float32)
        rec = {
            "case": name,
            "torch": torch.__version__,
            "state_checksum": state_checksum(state),
            "img_checksum": float(img.double().sum()),
            "labels": labels.clone(),
            "logits64": l64, "loss64": s64, "logits32": l32, "loss32": s32,
            "grad_probes": {k: grad_probes(k, g64[k], i) for i, k in enumerate(state.keys())},
            "grad_full": {k: g64[k].clone() for k in state.keys()
                          if k in FULL or k.startswith("mlp_head.") or k.startswith("norm.")},
        }
        path = os.path.join(OUT, name + ".pt")
        torch.save(rec, path)
        print(f"{name}: loss64={float(s64):.12f} loss32={float(s32):.8f} "
              f"|logits|={float(l64.norm()):.6f} -> {path} ({os.path.getsize(path)} B)")


def build_reference_encoder(name):
    """Construct the reference's ViT / ViT3D for an encoder case -> (model, state) with the seeded 'test' state loaded."""
    from oracle import encoders as E
    kind, _, ctor, B, M, sseed, _ = E.ENC_CASES[name]
    cfg = E.enc_config(name)
    rc = ref_loader.ConfigDict()
    for k, v in vars(cfg).items():
        setattr(rc, k, dict(vars(v)) if hasattr(v, "__dict__") else v)
    if kind == "cnnvit":
        model = ref_loader.load("model").ViT(rc)
    else:
        model = ref_loader.load("modelv2").ViT3D({}, 1e-4, 0.0, M, rc, **ctor)
    schema = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}
    state = E.make_state_generic(schema, sseed)
    model.load_state_dict(state, strict=True)
    return model, state


def main_encoders():
    """Golden vectors of the CNN-stem models (model.py ViT, modelv2.py ViT3D), train mode (BatchNorm batch statistics)."""
    from oracle import encoders as E
    for name, (kind, *_rest) in E.ENC_CASES.items():
        x, labels = E.enc_inputs(name)
        rec = {"case": name, "torch": torch.__version__, "img_checksum": float(x.double().sum())}
        for dt, tag in ((torch.float64, "64"), (torch.float32, "32")):
            model, state = build_reference_encoder(name)
            model = model.to(dt).train()
            logits, loss = model(x.to(dt), labels.to(dt) if kind == "cnnvit" else labels)
            loss.backward()
            rec["logits" + tag], rec["loss" + tag] = logits.detach(), loss.detach()
            if dt == torch.float64:
                rec["state_checksum"] = state_checksum({k: v for k, v in state.items() if v.is_floating_point()})
                rec["schema"] = {k: (tuple(v.shape), v.dtype) for k, v in state.items()}
                rec["grad_probes"] = {k: grad_probes(k, p.grad, i) for i, (k, p) in enumerate(model.named_parameters())}
        path = os.path.join(OUT, name + ".pt")
        torch.save(rec, path)
        print(f"{name}: loss64={float(rec['loss64']):.12f} loss32={float(rec['loss32']):.8f} -> {path} "
              f"({os.path.getsize(path)} B)")


if __name__ == "__main__":
    if "--encoders-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        main_encoders()
        sys.exit(0)
    main()
    main_encoders()
