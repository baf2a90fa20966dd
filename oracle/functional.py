"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — functional CPU restatement.

An independent restatement, in plain torch ops, of the reference's hot path. It takes a
``state_dict``-keyed mapping of tensors (schema: SURVEY.md §A.3) and a config attribute
bag, runs in whatever dtype the tensors have (fp32 for the stated tolerances, fp64 for the
tight reference) and is differentiable through autograd, so ``loss.backward()`` on leaf
copies of the weights gives the reference gradients.

Each function cites the reference lines it restates. Dropout
(/root/reference/model_cross.py:25,27,47,84,86,170,180,182) is the identity unless a ``dm``
mapping of multiplier tensors (0 or 1/(1-p), one per nn.Dropout call site) is passed: ATen's
Philox stream cannot be reproduced by another implementation, so p > 0 parity is checked by
replaying the masks the CUDA path used (tests/test_gpu_dropout.py).

Dropout site keys (all multipliers shaped like the tensor they scale):
  ("embed", m)                      x = dropout(x + pos)              model_cross.py:198 / modelv3.py:141
  ("out" | "gelu" | "fc2", l, m)    to_out / FeedForward dropouts of self block l (0-based over the
                                    whole depth) of stream m          model_cross.py:47,25,27
  ("f_attn" | "f_proj" | "f_gelu" | "f_fc2", mb, k)   fusion k of multi-scale block mb   :97,101,25,27
  ("head_gelu", m), ("head_logits", m)                                :180,182 / modelv3.py:115,117
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Mapping, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Mapping[str, Tensor]


# --------------------------------------------------------------------------- config
def make_config(**kw) -> SimpleNamespace:
    """Attribute bag with the fields the reference constructors read
    (/root/reference/model_cross.py:153-183, config2.py:3-28)."""
    base = dict(
        hidden_dim=1024, mlp_dim=4096, num_heads=16, num_multi_blocks=2, num_self_blocks=2,
        num_layers=4, patch_size=(16, 16, 8), img_size=(128, 128, 64), num_classes=2,
        dropout=0.0, lr=1e-4, weight_decay=5e-4, optim_params={"T_max": 250, "eta_min": 1e-6},
        label_smoothing=0.0, num_modalities=4,
        attn_order={"0": "1", "1": "2", "2": "3", "3": "0"},
    )
    base.update(kw)
    return SimpleNamespace(**base)


def num_patches(cfg) -> int:
    D, H, W = cfg.img_size
    dp, hp, wp = cfg.patch_size
    return (D // dp) * (H // hp) * (W // wp)


# --------------------------------------------------------------------------- patch map
def patchify(vol: Tensor, patch_size: Sequence[int]) -> Tensor:
    """``rearrange(x, 'b c (d p1) (h p2) (w p3) -> b (h w d) (p1 p2 p3 c)')``
    (/root/reference/model_cross.py:193, modelv3.py:129) without einops.
    vol: [B, c, D, H, W] -> [B, Hn*Wn*Dn, dp*hp*wp*c]; token (h w d) with d fastest,
    feature (p1 p2 p3 c)."""
    B, c, D, H, W = vol.shape
    dp, hp, wp = patch_size
    Dn, Hn, Wn = D // dp, H // hp, W // wp
    x = vol.reshape(B, c, Dn, dp, Hn, hp, Wn, wp)
    #            0  1   2   3   4   5   6   7
    x = x.permute(0, 4, 6, 2, 3, 5, 7, 1)  # b h w d p1 p2 p3 c
    return x.reshape(B, Hn * Wn * Dn, dp * hp * wp * c)


def patch_source_index(img_size, patch_size):
    """Closed-form index map of SURVEY.md §A.1: returns int64 [Np, P] flat offsets into a
    [D, H, W] volume such that patches[b, t, f] = vol[b, 0].flatten()[idx[t, f]]."""
    D, H, W = img_size
    dp, hp, wp = patch_size
    Dn, Hn, Wn = D // dp, H // hp, W // wp
    t = torch.arange(Hn * Wn * Dn)
    di = t % Dn
    wi = (t // Dn) % Wn
    hi = t // (Dn * Wn)
    f = torch.arange(dp * hp * wp)
    c = f % wp
    b = (f // wp) % hp
    a = f // (wp * hp)
    z0 = di[:, None] * dp + a[None, :]
    z1 = hi[:, None] * hp + b[None, :]
    z2 = wi[:, None] * wp + c[None, :]
    return (z0 * H + z1) * W + z2


# --------------------------------------------------------------------------- blocks
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm(hidden_dim): biased variance, affine (/root/reference/model_cross.py:14)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (/root/reference/model_cross.py:24)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _drop(dm, key, x: Tensor) -> Tensor:
    """nn.Dropout with an externally supplied multiplier tensor (identity when dm is None)."""
    if dm is None:
        return x
    return x * dm[key].to(x.dtype).reshape(x.shape)


def linear(x: Tensor, w: Tensor, b: Tensor | None = None) -> Tensor:
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def feed_forward(p: Params, pre: str, x: Tensor, dm=None, k_gelu=None, k_out=None) -> Tensor:
    """FeedForward.net: Linear, GELU, Dropout, Linear, Dropout (/root/reference/model_cross.py:19-31)."""
    h = _drop(dm, k_gelu, gelu_erf(linear(x, p[pre + "net.0.weight"], p[pre + "net.0.bias"])))
    return _drop(dm, k_out, linear(h, p[pre + "net.3.weight"], p[pre + "net.3.bias"]))


def self_attention(p: Params, pre: str, x: Tensor, heads: int, dm=None, k_out=None) -> Tensor:
    """Attention.forward (/root/reference/model_cross.py:50-61): bias-free packed QKV,
    'b n (h d) -> b h n d', softmax(q k^T * d^-0.5) v, merge heads, to_out (+bias)."""
    B, N, C = x.shape
    d = C // heads
    qkv = linear(x, p[pre + "to_qkv.weight"])
    q, k, v = qkv.split(C, dim=-1)

    def heads_first(t):
        return t.reshape(B, N, heads, d).permute(0, 2, 1, 3)

    q, k, v = heads_first(q), heads_first(k), heads_first(v)
    dots = (q @ k.transpose(-1, -2)) * (d ** -0.5)
    attn = torch.softmax(dots, dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(B, N, C)
    if heads == 1:  # project_out is False when heads == 1 and dim_head == hidden_dim
        return out  # (/root/reference/model_cross.py:37,44-48: to_out = nn.Identity())
    return _drop(dm, k_out, linear(out, p[pre + "to_out.0.weight"], p[pre + "to_out.0.bias"]))


def self_attention_block(p: Params, pre: str, x: Tensor, heads: int, dm=None, l: int = 0, m: int = 0) -> Tensor:
    """SelfAttentionBlock.forward (/root/reference/model_cross.py:69-72) with PreNorm (:11-17)."""
    xn = layer_norm(x, p[pre + "attn.norm.weight"], p[pre + "attn.norm.bias"])
    x = self_attention(p, pre + "attn.fn.", xn, heads, dm, ("out", l, m)) + x
    xn = layer_norm(x, p[pre + "ffn.norm.weight"], p[pre + "ffn.norm.bias"])
    x = feed_forward(p, pre + "ffn.fn.", xn, dm, ("gelu", l, m), ("fc2", l, m)) + x
    return x


def cross_attention(p: Params, pre: str, x: Tensor, heads: int, dm=None, mb: int = 0, kf: int = 0) -> Tensor:
    """CrossAttention.forward (/root/reference/model_cross.py:88-102): query from token 0
    only, keys/values from all N tokens (token 0 included), biased projections."""
    B, N, C = x.shape
    d = C // heads
    q = linear(x[:, 0:1], p[pre + "wq.weight"], p[pre + "wq.bias"]).reshape(B, 1, heads, d).permute(0, 2, 1, 3)
    k = linear(x, p[pre + "wk.weight"], p[pre + "wk.bias"]).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    v = linear(x, p[pre + "wv.weight"], p[pre + "wv.bias"]).reshape(B, N, heads, d).permute(0, 2, 1, 3)
    attn = _drop(dm, ("f_attn", mb, kf), torch.softmax((q @ k.transpose(-2, -1)) * (d ** -0.5), dim=-1))
    y = (attn @ v).transpose(1, 2).reshape(B, 1, C)
    return _drop(dm, ("f_proj", mb, kf), linear(y, p[pre + "proj.weight"], p[pre + "proj.bias"]))


def cross_attention_block(p: Params, pre: str, x: Tensor, heads: int, dm=None, mb: int = 0, kf: int = 0) -> Tensor:
    """CrossAttentionBlock.forward (/root/reference/model_cross.py:111-114): residual uses
    the UN-normalised CLS row; FFN runs on that single token."""
    xn = layer_norm(x, p[pre + "attn.norm.weight"], p[pre + "attn.norm.bias"])
    y = cross_attention(p, pre + "attn.fn.", xn, heads, dm, mb, kf) + x[:, 0:1]
    yn = layer_norm(y, p[pre + "ffn.norm.weight"], p[pre + "ffn.norm.bias"])
    return feed_forward(p, pre + "ffn.fn.", yn, dm, ("f_gelu", mb, kf), ("f_fc2", mb, kf)) + y


def multi_scale_block(p: Params, pre: str, xs: List[Tensor], cfg, dm=None, mb: int = 0) -> List[Tensor]:
    """MultiScaleBlock.forward (/root/reference/model_cross.py:128-148). All fusions read
    the post-self-attention streams of THIS block; fusion modules are indexed by a running
    count over ascending i among the keys present."""
    M = len(xs)
    attn = []
    for m in range(M):
        x = xs[m]
        for sb in range(cfg.num_self_blocks):
            x = self_attention_block(p, f"{pre}blocks.{m}.{sb}.", x, cfg.num_heads, dm, mb * cfg.num_self_blocks + sb, m)
        attn.append(x)
    outs = []
    k = 0
    for i in range(M):
        if str(i) in cfg.attn_order:
            j = int(cfg.attn_order[str(i)])
            tmp = torch.cat((attn[i][:, 0:1], attn[j][:, 1:]), dim=1)
            tmp = cross_attention_block(p, f"{pre}fusion.{k}.", tmp, cfg.num_heads, dm, mb, k)
            outs.append(torch.cat((tmp, attn[i][:, 1:]), dim=1))
            k += 1
        else:
            outs.append(attn[i])
    return outs


def embed_stream(p: Params, vol: Tensor, cfg) -> Tensor:
    """Per-stream tokenisation (/root/reference/model_cross.py:193-198): unfold, shared
    Linear(P->C), prepend shared CLS, add shared positional embedding."""
    x = patchify(vol, cfg.patch_size)
    x = linear(x, p["patch_to_embedding.weight"], p["patch_to_embedding.bias"])
    cls = p["cls_token"].expand(vol.shape[0], -1, -1)
    x = torch.cat((cls, x), dim=1)
    return x + p["pos_embedding"]


def cross_entropy(logits: Tensor, labels: Tensor, label_smoothing: float = 0.0) -> Tensor:
    """F.cross_entropy(mean reduction, label_smoothing) in closed form
    (/root/reference/model_cross.py:211)."""
    logp = torch.log_softmax(logits, dim=-1)
    K = logits.shape[-1]
    nll = -logp.gather(1, labels[:, None]).squeeze(1)
    smooth = -logp.sum(dim=-1) / K
    return ((1.0 - label_smoothing) * nll + label_smoothing * smooth).mean()


def model_cross_forward(p: Params, img: Tensor, labels: Tensor, cfg, return_tokens: bool = False, dm=None):
    """ModelCross.forward (/root/reference/model_cross.py:186-212).
    img [B, M, 1, D, H, W]; labels int64 [B]. Returns (logits [B, classes], loss)."""
    M = img.shape[1]
    xs = [_drop(dm, ("embed", m), embed_stream(p, img[:, m], cfg)) for m in range(M)]
    for mb in range(cfg.num_multi_blocks):
        xs = multi_scale_block(p, f"transformer.{mb}.", xs, cfg, dm, mb)
    tokens = xs
    heads = []
    for m in range(M):
        xn = layer_norm(xs[m], p[f"norm.{m}.weight"], p[f"norm.{m}.bias"])[:, 0]
        h = _drop(dm, ("head_gelu", m), gelu_erf(linear(xn, p[f"mlp_head.{m}.0.weight"], p[f"mlp_head.{m}.0.bias"])))
        heads.append(_drop(dm, ("head_logits", m), linear(h, p[f"mlp_head.{m}.3.weight"], p[f"mlp_head.{m}.3.bias"])))
    logits = torch.stack(heads).mean(dim=0)
    loss = cross_entropy(logits, labels, cfg.label_smoothing)
    if return_tokens:
        return logits, loss, tokens
    return logits, loss


def model_vit_forward(p: Params, img: Tensor, labels: Tensor, cfg, dm=None):
    """ModelVIT.forward (/root/reference/modelv3.py:123-147): streams concatenated on the
    token axis BEFORE the single CLS / positional embedding; `num_layers` pre-norm blocks;
    head = LN, Linear, GELU, Linear on the CLS row; plain CE."""
    M = img.shape[1]
    toks = [linear(patchify(img[:, m], cfg.patch_size), p["patch_to_embedding.weight"],
                   p["patch_to_embedding.bias"]) for m in range(M)]
    x = torch.cat(toks, dim=1)
    x = torch.cat((p["cls_token"].expand(img.shape[0], -1, -1), x), dim=1) + p["pos_embedding"]
    x = _drop(dm, ("embed", 0), x)
    for l in range(cfg.num_layers):
        pre = f"transformer.layers.{l}."
        xn = layer_norm(x, p[pre + "0.norm.weight"], p[pre + "0.norm.bias"])
        x = self_attention(p, pre + "0.fn.", xn, cfg.num_heads, dm, ("out", l, 0)) + x
        xn = layer_norm(x, p[pre + "2.norm.weight"], p[pre + "2.norm.bias"])
        x = feed_forward(p, pre + "2.fn.", xn, dm, ("gelu", l, 0), ("fc2", l, 0)) + x
    c = layer_norm(x[:, 0], p["mlp_head.0.weight"], p["mlp_head.0.bias"])
    h = _drop(dm, ("head_gelu", 0), gelu_erf(linear(c, p["mlp_head.1.weight"], p["mlp_head.1.bias"])))
    logits = _drop(dm, ("head_logits", 0), linear(h, p["mlp_head.4.weight"], p["mlp_head.4.bias"]))
    return logits, cross_entropy(logits, labels, 0.0)


# --------------------------------------------------------------------------- helpers
def leaf_params(p: Params, dtype=torch.float64) -> Dict[str, Tensor]:
    """Detached leaf copies (requires_grad) of a state dict in `dtype`."""
    return {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items()}


def forward_backward(p: Params, img: Tensor, labels: Tensor, cfg, kind: str = "cross",
                     dtype=torch.float64, dm=None):
    """Run fwd+bwd of the restatement; returns (logits, loss, grads dict)."""
    lp = leaf_params(p, dtype)
    fwd = model_cross_forward if kind == "cross" else model_vit_forward
    logits, loss = fwd(lp, img.to(dtype), labels, cfg, dm=dm)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in lp.items()}
    return logits.detach(), loss.detach(), grads
