"""Noise floor of the bf16-mode logits at BASELINE.json configs[1] (cfg2_b2), and the embedding's own share of it.

The four reference logits of tests/golden/full_cfg2_b2.pt have ||logits||_2 = 0.088, so the relative logits error of the bf16
mode is the network's absolute rounding noise divided by a small number. This test measures that noise directly: the same
case is run with the input scaled by 1 + eps for a few eps ~ 1e-6 — far below bf16 resolution, but enough to flip rounding
decisions throughout the network — with the fused TF32 embedding (csrc/embed.cu) and with patchify + bf16 GEMM. The spread of
the relative error over those runs is the noise floor the tolerance of tests/test_gpu_baseline_shapes.py has to sit above;
the absolute error stays below 5e-3 in every run. It also checks the embedding output itself against fp64: the TF32 operands
(fp32 words read with 10 mantissa bits) are closer than the bf16 copies.

The report goes to gpurun_out/logit_noise_floor.txt (committed copy: profiles/logit_noise_floor_r02.txt)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.full_cases import build_full_case          # noqa: E402
from oracle.functional import patchify                 # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "logit_noise_floor.txt")
NAME = "cfg2_b2"


def test_bf16_mode_logit_noise_floor_and_embedding_share(monkeypatch):
    from cavit import _abi, ops
    from cavit._abi import EPI_EMBED
    from cavit.config import make_config
    from cavit.modules import ModelCross, ModelVIT
    _abi.require_device(0)
    path = os.path.join(ROOT, "tests", "golden", f"full_{NAME}.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    rec = torch.load(path, weights_only=False)
    ref = rec["logits64"]
    lines = [f"{NAME}: ||ref logits||_2 = {float(ref.norm()):.4f}"]

    # ---- the embedding output of both paths against fp64
    kind, cfg, model, img, labels = build_full_case(NAME, ModelCross, ModelVIT, make_config)
    sd = model.state_dict()
    W, bias = sd["patch_to_embedding.weight"].cuda().contiguous(), sd["patch_to_embedding.bias"].cuda()
    B, M = img.shape[0], img.shape[1]
    patch, C, P = tuple(cfg.patch_size), W.shape[0], W.shape[1]
    pat = torch.stack([patchify(img[:, m], patch) for m in range(M)], 1).double()
    Np = pat.shape[2]
    N = Np + 1
    pos = sd["pos_embedding"][0, :N].cuda().contiguous()
    proj = torch.einsum("bmtp,cp->mbtc", pat, W.double().cpu())
    want = proj + bias.double().cpu() + pos[1:].double().cpu()
    a = torch.zeros(M, B * N, C, device="cuda")
    b = torch.zeros_like(a)
    imgc = img.cuda().contiguous()
    ops.embed_fused_fwd(imgc, W, bias, pos, a, patch_size=patch, C_=C)
    patches = torch.empty(M * B * Np, P, device="cuda", dtype=torch.bfloat16)
    ops.patchify(imgc, patches, patch_size=patch)
    ops.gemm(patches, W.bfloat16(), b, M=M * B * Np, N=C, K=P, lda=P, ldb=P, ldo=C, epi=EPI_EMBED, bias=bias, resid=pos, ldr=C,
             embed_np=Np)
    err = {}
    for tag, t in (("fused, TF32 operands", a), ("patchify + GEMM, bf16 operands", b)):
        g = t.view(M, B, N, C)[:, :, 1:].double().cpu()
        err[tag] = float((g - want).norm() / proj.norm())
        lines.append(f"embedding output vs fp64, {tag}: ||err|| / ||x W^T|| = {err[tag]:.2e}")
    assert err["fused, TF32 operands"] < 0.6 * err["patchify + GEMM, bf16 operands"], err

    # ---- logits under tiny input perturbations, both embedding paths
    rels = []
    for fused in ("1", "0"):
        monkeypatch.setenv("CAVIT_EMBED_FUSED", fused)
        for eps in (0.0, 1e-6, -1e-6, 3e-6, -3e-6):
            kind, cfg, model, img, labels = build_full_case(NAME, ModelCross, ModelVIT, make_config)
            model = model.cuda().train()
            logits, _ = model((img * (1 + eps)).cuda(), labels.cuda())
            d = logits.detach().double().cpu()
            rel, ab = float((d - ref).norm() / ref.norm()), float((d - ref).norm())
            rels.append(rel)
            lines.append(f"bf16 mode, CAVIT_EMBED_FUSED={fused}, input * (1 {eps:+.0e}): logits rel {rel:.4f}  abs {ab:.2e}")
            assert ab < 5e-3, lines[-1]
            del model
    lines.append(f"spread of the relative logits error over the {len(rels)} runs: {min(rels):.4f} ... {max(rels):.4f}")
    print("\n".join(lines))
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "w") as f:
            f.write("\n".join(lines) + "\n")
    except OSError:
        pass
    assert _abi.device_status() == 0
