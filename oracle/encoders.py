"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — CPU restatement of the reference's two CNN-stem models.

``vit_forward``   : ``ViT.forward``   (/root/reference/model.py:253-286) incl. ``Embeddings`` (:79-105),
                    ``CNNEncoder`` (:23-75), ``Block`` / ``MultiHeadAttention`` / ``Mlp`` (:107-193), ``Encoder`` (:196-206).
``vit3d_forward`` : ``ViT3D.forward`` (/root/reference/modelv2.py:187-241) incl. ``CNN3DEncoder`` (:14-58) and the
                    post-norm ``nn.TransformerEncoderLayer`` stack (:61-87; arithmetic defined by torch:
                    x = LN1(x + MHA(x)); x = LN2(x + W2 relu(W1 x + b1) + b2), packed biased in-proj, eps 1e-5).

Plain torch ops on a ``state_dict``-keyed mapping; runs in the tensors' dtype (fp64 for the tight pin) and is
differentiable, like oracle/functional.py. Dropout rates are 0 in every parity case.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Mapping

import torch
import torch.nn.functional as F

from .functional import cross_entropy, gelu_erf, layer_norm, linear

Tensor = torch.Tensor
Params = Mapping[str, Tensor]


# ----------------------------------------------------------------------------------------- shared
def mha(x: Tensor, w_in: Tensor, b_in: Tensor, w_out: Tensor, b_out: Tensor, heads: int) -> Tensor:
    """softmax(q k^T / sqrt(d)) v with biased projections (model.py:153-176; torch MultiheadAttention)."""
    B, N, C = x.shape
    d = C // heads
    q, k, v = linear(x, w_in, b_in).split(C, dim=-1)

    def hf(t):
        return t.reshape(B, N, heads, d).permute(0, 2, 1, 3)

    p = torch.softmax((hf(q) @ hf(k).transpose(-1, -2)) / math.sqrt(d), dim=-1)
    ctx = (p @ hf(v)).permute(0, 2, 1, 3).reshape(B, N, C)
    return linear(ctx, w_out, b_out)


# ----------------------------------------------------------------------------------------- ViT (model.py)
def make_vit_config(**kw) -> SimpleNamespace:
    """The attribute bag `ViT(config)` reads (model.py:79-90,107-146,196-224); the reference's own config.py lacks
    these fields, so a hand-made one is the only way to construct it."""
    base = dict(hidden_size=128, transformer={"num_heads": 2, "num_layers": 2, "mlp_dim": 256, "dropout_rate": 0.0,
                                              "attention_dropout_rate": 0.0},
                encoder_channels=(4, 8, 16), down_factor=2, patches=SimpleNamespace(grid=(2, 2, 2)), img_size=(16, 16, 16))
    base.update(kw)
    return SimpleNamespace(**base)


def _double_conv(p: Params, pre: str, x: Tensor) -> Tensor:
    x = F.relu(F.conv3d(x, p[pre + "0.weight"], p[pre + "0.bias"], padding=1))
    return F.relu(F.conv3d(x, p[pre + "2.weight"], p[pre + "2.bias"], padding=1))


def cnn_encoder(p: Params, pre: str, x: Tensor) -> Tensor:
    """CNNEncoder.forward (model.py:66-74): DoubleConv, then two (MaxPool3d(2), DoubleConv)."""
    x = _double_conv(p, pre + "inc.double_conv.", x)
    x = _double_conv(p, pre + "down1.maxpool_conv.1.double_conv.", F.max_pool3d(x, 2))
    return _double_conv(p, pre + "down2.maxpool_conv.1.double_conv.", F.max_pool3d(x, 2))


def vit_embeddings(p: Params, x: Tensor, cfg) -> Tensor:
    """Embeddings.forward (model.py:92-105) for one modality: stem, strided Conv3d patch embedding, flatten(-3),
    transpose, CLS, positional add."""
    grid = tuple(cfg.patches.grid)
    f = cnn_encoder(p, "embeddings.cnn_encoder.", x)
    t = F.conv3d(f, p["embeddings.patch_embed.weight"], p["embeddings.patch_embed.bias"], stride=grid)
    t = t.flatten(-3).transpose(-2, -1)
    cls = p["embeddings.class_token"].expand(t.shape[0], -1, -1)
    return torch.cat((cls, t), dim=1) + p["embeddings.positional_embedding"]


def vit_forward(p: Params, x: Tensor, label, cfg):
    """ViT.forward (model.py:253-286): x [B, M, 1, A, B', C']; only modality 0 keeps its CLS row."""
    M = x.shape[1]
    H = cfg.transformer["num_heads"]
    toks = [vit_embeddings(p, x.select(1, 0), cfg)] + [vit_embeddings(p, x.select(1, i), cfg)[:, 1:] for i in range(1, M)]
    h = torch.cat(toks, dim=1)
    for l in range(cfg.transformer["num_layers"]):
        pre = f"encoder.layers.{l}."
        xn = layer_norm(h, p[pre + "attention_norm.weight"], p[pre + "attention_norm.bias"], 1e-6)
        w_in = torch.cat([p[pre + f"multi_head.{n}.weight"] for n in ("query", "key", "value")], dim=0)
        b_in = torch.cat([p[pre + f"multi_head.{n}.bias"] for n in ("query", "key", "value")], dim=0)
        h = h + mha(xn, w_in, b_in, p[pre + "multi_head.out.weight"], p[pre + "multi_head.out.bias"], H)
        xn = layer_norm(h, p[pre + "ffn_norm.weight"], p[pre + "ffn_norm.bias"], 1e-6)
        u = gelu_erf(linear(xn, p[pre + "ffn.fc1.weight"], p[pre + "ffn.fc1.bias"]))
        h = h + linear(u, p[pre + "ffn.fc2.weight"], p[pre + "ffn.fc2.bias"])
    h = layer_norm(h, p["encoder.encoder_norm.weight"], p["encoder.encoder_norm.bias"], 1e-6)
    z = linear(h[:, 0], p["final.weight"], p["final.bias"]).squeeze(-1)
    if label is None:
        return z
    y = label.to(z.dtype)
    loss = (torch.clamp(z, min=0) - z * y + torch.log1p(torch.exp(-z.abs()))).mean()   # BCEWithLogitsLoss
    return z, loss


# ----------------------------------------------------------------------------------------- ViT3D (modelv2.py)
def make_vit3d_config(**kw) -> SimpleNamespace:
    base = dict(hidden_dim=128, transformer=SimpleNamespace(num_heads=2, num_layers=2), img_size=(32, 32, 32))
    base.update(kw)
    return SimpleNamespace(**base)


def cnn3d_encoder(p: Params, pre: str, x: Tensor, training: bool) -> Tensor:
    """CNN3DEncoder.forward (modelv2.py:41-58): conv+BN+ReLU x4, max-pool after the first two, stride 2 in the last two.
    BatchNorm3d uses batch statistics when `training` (running buffers are not updated here)."""
    for i, stride in enumerate((1, 1, 2, 2), start=1):
        x = F.conv3d(x, p[f"{pre}conv{i}.weight"], p[f"{pre}conv{i}.bias"], stride=stride, padding=1)
        x = F.batch_norm(x, None if training else p[f"{pre}bn{i}.running_mean"],
                         None if training else p[f"{pre}bn{i}.running_var"], p[f"{pre}bn{i}.weight"], p[f"{pre}bn{i}.bias"],
                         training, 0.1, 1e-5)
        x = F.relu(x)
        if i <= 2:
            x = F.max_pool3d(x, 2, 2)
    return x


def vit3d_tokens(p: Params, x: Tensor, training: bool) -> Tensor:
    """modelv2.py:203-212: per-modality stem, flatten, concatenate on the token axis -> [B, C, M*S]."""
    return torch.cat([cnn3d_encoder(p, "encoder_3d.", x.select(1, m), training).flatten(start_dim=2)
                      for m in range(x.shape[1])], dim=2)


def vit3d_core(p: Params, feat: Tensor, labels: Tensor, heads: int, layers: int, label_smoothing: float = 0.0,
               relu_masks=None, mask_report=None):
    """modelv2.py:214-241 from the concatenated stem features [B, C, N-1] on.

    relu_masks (optional, one bool [B, N, 4C] tensor per layer): replay the ReLU activation pattern of another
    implementation instead of the oracle's own `u > 0`. ReLU's derivative is discontinuous, so an implementation with
    bf16 operands flips the units whose pre-activation is within rounding error of zero, and each flipped unit changes
    its gradient entry by 100 %: with the pattern replayed the comparison measures arithmetic error only (the same
    device as the dropout-mask replay of oracle/functional.py). mask_report collects, per layer,
    (fraction of units that differ from the oracle's own pattern, max |u| / std(u) over those units)."""
    B = feat.shape[0]
    h = feat.transpose(1, 2)
    has_cls = "cls_token" in p      # ViT3D(add_cls_token=False) registers no cls_token (modelv2.py:140-143)
    if has_cls:
        h = torch.cat((p["cls_token"].expand(B, -1, -1), h), dim=1)
    h = h + p["pos_embed"]
    for l in range(layers):
        pre = f"transformer.transformer.layers.{l}."
        a = mha(h, p[pre + "self_attn.in_proj_weight"], p[pre + "self_attn.in_proj_bias"],
                p[pre + "self_attn.out_proj.weight"], p[pre + "self_attn.out_proj.bias"], heads)
        h = layer_norm(h + a, p[pre + "norm1.weight"], p[pre + "norm1.bias"], 1e-5)
        u = linear(h, p[pre + "linear1.weight"], p[pre + "linear1.bias"])
        if relu_masks is None:
            act = torch.relu(u)
        else:
            mk = relu_masks[l].to(u.device)
            if mask_report is not None:
                diff = mk != (u.detach() > 0)
                worst = float((u.detach().abs()[diff]).max() / u.detach().std()) if bool(diff.any()) else 0.0
                mask_report.append((float(diff.float().mean()), worst))
            act = u * mk.to(u.dtype)
        f = linear(act, p[pre + "linear2.weight"], p[pre + "linear2.bias"])
        h = layer_norm(h + f, p[pre + "norm2.weight"], p[pre + "norm2.bias"], 1e-5)
    pooled = h[:, 0] if has_cls else h.mean(dim=1)     # modelv2.py:229-235
    c = layer_norm(pooled, p["mlp_head.0.weight"], p["mlp_head.0.bias"], 1e-5)
    logits = linear(linear(c, p["mlp_head.1.weight"], p["mlp_head.1.bias"]), p["mlp_head.2.weight"], p["mlp_head.2.bias"])
    return logits, cross_entropy(logits, labels, label_smoothing)


def vit3d_forward(p: Params, x: Tensor, labels: Tensor, cfg, label_smoothing: float = 0.0, training: bool = True,
                  relu_masks=None, mask_report=None):
    feat = vit3d_tokens(p, x, training)
    return vit3d_core(p, feat, labels, cfg.transformer.num_heads, cfg.transformer.num_layers, label_smoothing,
                      relu_masks, mask_report)


# ----------------------------------------------------------------------------------------- deterministic states / cases
def make_state_generic(schema: "Mapping[str, tuple]", seed: int) -> "OrderedDict[str, Tensor]":
    """Seeded fill of an arbitrary state_dict schema {name: (shape, dtype)} with non-trivial values for every tensor
    (non-zero biases / affine parameters / BatchNorm buffers) so that each code path matters in the parity tests."""
    g = torch.Generator().manual_seed(seed)
    out: "OrderedDict[str, Tensor]" = OrderedDict()
    for name, (shp, dt) in schema.items():
        shp = tuple(shp)
        if name.endswith("num_batches_tracked"):
            t = torch.zeros(shp, dtype=torch.int64)
        elif name.endswith("running_var"):
            t = 1.0 + 0.2 * torch.rand(shp, generator=g, dtype=torch.float64)
        elif name.endswith("running_mean"):
            t = 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)
        elif any(s in name for s in ("pos_embed", "positional_embedding", "cls_token", "class_token")):
            t = 0.02 * torch.randn(shp, generator=g, dtype=torch.float64)
        elif len(shp) >= 2:
            rf = 1
            for s in shp[2:]:
                rf *= s
            a = math.sqrt(6.0 / ((shp[0] + shp[1]) * rf))
            t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * a
        elif name.endswith("weight"):      # LayerNorm / BatchNorm scale
            t = 1.0 + 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)
        else:                               # every bias
            t = 0.05 * torch.randn(shp, generator=g, dtype=torch.float64)
        out[name] = t if t.dtype == torch.int64 else t.to(torch.float32)
    return out


ENC_CASES = {
    # name: (kind, config kwargs, ctor kwargs, batch, modalities, state seed, input seed)
    "cnnvit_small": ("cnnvit", dict(), dict(), 3, 2, 21, 31),
    "vit3d_small": ("vit3d", dict(), dict(num_classes=2, label_smoothing=0.1), 2, 3, 22, 33),
    "vit3d_meanpool": ("vit3d", dict(), dict(num_classes=3, add_cls_token=False), 2, 2, 23, 34),
}


def enc_config(name: str):
    kind, kw, *_ = ENC_CASES[name]
    return make_vit_config(**kw) if kind == "cnnvit" else make_vit3d_config(**kw)


def enc_inputs(name: str):
    kind, _, ctor, B, M, _, iseed = ENC_CASES[name]
    cfg = enc_config(name)
    g = torch.Generator().manual_seed(iseed)
    x = torch.randn((B, M, 1) + tuple(cfg.img_size), generator=g, dtype=torch.float32)
    if kind == "cnnvit":
        labels = torch.randint(0, 2, (B,), generator=g).to(torch.float32)
    else:
        labels = torch.randint(0, ctor.get("num_classes", 2), (B,), generator=g)
    return x, labels


def enc_forward_backward(name: str, state: Params, x: Tensor, labels: Tensor, dtype=torch.float64, relu_masks=None,
                         mask_report=None):
    """fwd+bwd of the restatement -> (logits, loss, grads dict)."""
    kind, _, ctor, *_ = ENC_CASES[name]
    cfg = enc_config(name)
    lp: Dict[str, Tensor] = {}
    for k, v in state.items():
        if not v.is_floating_point():
            lp[k] = v
        elif k.endswith(("running_mean", "running_var")):     # BatchNorm buffers: not parameters
            lp[k] = v.detach().to(dtype)
        else:
            lp[k] = v.detach().to(dtype).clone().requires_grad_(True)
    xin = x.to(dtype)
    if kind == "cnnvit":
        logits, loss = vit_forward(lp, xin, labels, cfg)
    else:
        logits, loss = vit3d_forward(lp, xin, labels, cfg, ctor.get("label_smoothing", 0.0), training=True,
                                     relu_masks=relu_masks, mask_report=mask_report)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in lp.items()
             if v.is_floating_point() and v.requires_grad}
    return logits.detach(), loss.detach(), grads
