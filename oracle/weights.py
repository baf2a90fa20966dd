"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — deterministic weights and inputs.

`state_schema_*` enumerate the reference ``state_dict`` (name -> shape) in registration
order (/root/reference/model_cross.py:153-183, modelv3.py:91-121).
`make_state` fills it from a seeded CPU generator:

* ``init="reference"``  the reference initialiser's distributions (Xavier-uniform Linear
  weights, zero biases, LayerNorm 1/0, pos/cls N(0, 0.02^2); model_cross.py:214-241);
* ``init="test"``       same weights but random non-zero biases and LayerNorm affine
  parameters, so every bias / affine code path is exercised by the parity tests.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .functional import num_patches


def _ffn(schema, pre, C, F):
    schema[pre + "net.0.weight"] = (F, C)
    schema[pre + "net.0.bias"] = (F,)
    schema[pre + "net.3.weight"] = (C, F)
    schema[pre + "net.3.bias"] = (C,)


def _ln(schema, pre, C):
    schema[pre + "weight"] = (C,)
    schema[pre + "bias"] = (C,)


def state_schema_cross(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    C, F, M = cfg.hidden_dim, cfg.mlp_dim, cfg.num_modalities
    Np = num_patches(cfg)
    P = cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["pos_embedding"] = (1, Np + 1, C)
    s["cls_token"] = (1, 1, C)
    s["patch_to_embedding.weight"] = (C, P)
    s["patch_to_embedding.bias"] = (C,)
    for mb in range(cfg.num_multi_blocks):
        for m in range(M):
            for sb in range(cfg.num_self_blocks):
                pre = f"transformer.{mb}.blocks.{m}.{sb}."
                _ln(s, pre + "attn.norm.", C)
                s[pre + "attn.fn.to_qkv.weight"] = (3 * C, C)
                if cfg.num_heads != 1:  # project_out quirk, model_cross.py:37,44-48
                    s[pre + "attn.fn.to_out.0.weight"] = (C, C)
                    s[pre + "attn.fn.to_out.0.bias"] = (C,)
                _ln(s, pre + "ffn.norm.", C)
                _ffn(s, pre + "ffn.fn.", C, F)
        for k in range(len(cfg.attn_order)):
            pre = f"transformer.{mb}.fusion.{k}."
            _ln(s, pre + "attn.norm.", C)
            for nm in ("wq", "wk", "wv", "proj"):
                s[pre + f"attn.fn.{nm}.weight"] = (C, C)
                s[pre + f"attn.fn.{nm}.bias"] = (C,)
            _ln(s, pre + "ffn.norm.", C)
            _ffn(s, pre + "ffn.fn.", C, F)
    for m in range(M):
        _ln(s, f"norm.{m}.", C)
    for m in range(M):
        s[f"mlp_head.{m}.0.weight"] = (F, C)
        s[f"mlp_head.{m}.0.bias"] = (F,)
        s[f"mlp_head.{m}.3.weight"] = (cfg.num_classes, F)
        s[f"mlp_head.{m}.3.bias"] = (cfg.num_classes,)
    return s


def state_schema_vit(cfg) -> "OrderedDict[str, Tuple[int, ...]]":
    C, F, M = cfg.hidden_dim, cfg.mlp_dim, cfg.num_modalities
    Np = num_patches(cfg) * M
    P = cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    s["pos_embedding"] = (1, Np + 1, C)
    s["cls_token"] = (1, 1, C)
    s["patch_to_embedding.weight"] = (C, P)
    s["patch_to_embedding.bias"] = (C,)
    for l in range(cfg.num_layers):
        pre = f"transformer.layers.{l}."
        _ln(s, pre + "0.norm.", C)
        s[pre + "0.fn.to_qkv.weight"] = (3 * C, C)
        if cfg.num_heads != 1:  # project_out quirk, modelv3.py:44,51-55
            s[pre + "0.fn.to_out.0.weight"] = (C, C)
            s[pre + "0.fn.to_out.0.bias"] = (C,)
        _ln(s, pre + "2.norm.", C)
        _ffn(s, pre + "2.fn.", C, F)
    _ln(s, "mlp_head.0.", C)
    s["mlp_head.1.weight"] = (F, C)
    s["mlp_head.1.bias"] = (F,)
    s["mlp_head.4.weight"] = (cfg.num_classes, F)
    s["mlp_head.4.bias"] = (cfg.num_classes,)
    return s


def _is_ln(name: str) -> bool:
    return (".norm." in name or name.startswith("norm.") or name.startswith("mlp_head.0.")) and \
        not name.endswith(".fn.weight")


def make_state(schema, seed: int = 0, init: str = "test", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = OrderedDict()
    ln_names = set()
    # LayerNorm tensors are the 1-D ".weight"/".bias" pairs whose parent has no 2-D weight.
    twod_parents = {k.rsplit(".", 1)[0] for k, shp in schema.items() if len(shp) == 2}
    for name, shp in schema.items():
        parent = name.rsplit(".", 1)[0]
        if len(shp) == 1 and parent not in twod_parents:
            ln_names.add(name)
    for name, shp in schema.items():
        if name in ("pos_embedding", "cls_token"):
            t = torch.randn(shp, generator=g, dtype=torch.float64) * 0.02
        elif len(shp) == 2:
            fan_out, fan_in = shp
            a = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * a
        elif name in ln_names:
            if name.endswith("weight"):
                t = torch.ones(shp, dtype=torch.float64)
                if init == "test":
                    t = t + 0.1 * torch.randn(shp, generator=g, dtype=torch.float64)
            else:
                t = torch.zeros(shp, dtype=torch.float64)
                if init == "test":
                    t = 0.05 * torch.randn(shp, generator=g, dtype=torch.float64)
        else:  # Linear bias
            t = torch.zeros(shp, dtype=torch.float64)
            if init == "test":
                t = 0.05 * torch.randn(shp, generator=g, dtype=torch.float64)
        out[name] = t.to(dtype)
    return out


def make_inputs(cfg, batch: int, seed: int = 1234, dtype=torch.float32, mri_like: bool = False):
    """Synthetic volumes [B, M, 1, D, H, W] N(0,1) (or raw-MRI-like intensities) + labels."""
    g = torch.Generator().manual_seed(seed)
    D, H, W = cfg.img_size
    img = torch.randn((batch, cfg.num_modalities, 1, D, H, W), generator=g, dtype=torch.float32)
    if mri_like:
        img = (img * 1000.0 + 2000.0).clamp_min(0.0)
    labels = torch.randint(0, cfg.num_classes, (batch,), generator=g)
    return img.to(dtype), labels


def state_checksum(state) -> float:
    """Order-dependent scalar fingerprint of a state dict (detects RNG drift)."""
    acc = 0.0
    for i, (k, v) in enumerate(state.items()):
        acc += float(v.double().sum()) * (1.0 + 1e-3 * i) + float(v.double().abs().sum())
    return acc
