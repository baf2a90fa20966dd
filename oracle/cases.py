"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — the named parity cases.

Every case has head_dim 64 (the only head_dim in BASELINE.json's configs and the one the
CUDA kernels specialise), ragged "tile + 1" sequence lengths, and small enough shapes that
the fp64 oracle finishes in seconds. The same table drives tests/golden generation
(oracle/gen_golden.py), the CPU oracle tests and the GPU parity tests.
"""
from __future__ import annotations

import torch

from .functional import make_config

RING4 = {"0": "1", "1": "2", "2": "3", "3": "0"}

CASES = {
    # name: (kind, config kwargs, batch, state seed, input seed)
    "cross_ring4": ("cross", dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_multi_blocks=2,
                                  num_self_blocks=2, patch_size=(16, 16, 8), img_size=(32, 32, 16),
                                  num_modalities=4, attn_order=RING4, label_smoothing=0.1), 2, 0, 1234),
    "cross_chain3": ("cross", dict(hidden_dim=128, mlp_dim=192, num_heads=2, num_multi_blocks=1,
                                   num_self_blocks=1, patch_size=(8, 16, 8), img_size=(16, 32, 16),
                                   num_modalities=3, attn_order={"0": "1", "1": "2"},
                                   label_smoothing=0.0), 3, 1, 77),
    "cross_heads3": ("cross", dict(hidden_dim=192, mlp_dim=256, num_heads=3, num_multi_blocks=1,
                                   num_self_blocks=1, patch_size=(16, 16, 1), img_size=(48, 32, 1),
                                   num_modalities=2, attn_order={"0": "1", "1": "0"},
                                   label_smoothing=0.0), 2, 2, 5),
    # heads == 1: the reference drops the attention out-projection (to_out = Identity)
    "cross_noattn_h1": ("cross", dict(hidden_dim=64, mlp_dim=128, num_heads=1, num_multi_blocks=1,
                                   num_self_blocks=2, patch_size=(8, 8, 8), img_size=(16, 16, 16),
                                   num_modalities=2, attn_order={}, label_smoothing=0.0), 2, 3, 9),
    "vit_small": ("vit", dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_layers=2,
                              patch_size=(8, 16, 8), img_size=(16, 32, 16), num_modalities=2,
                              attn_order={}, label_smoothing=0.0), 2, 4, 11),
}


def build_case(name: str):
    """-> (kind, cfg, state(fp32, 'test' init), img fp32, labels)."""
    from .weights import make_inputs, make_state, state_schema_cross, state_schema_vit
    kind, kw, batch, sseed, iseed = CASES[name]
    cfg = make_config(**kw)
    schema = state_schema_cross(cfg) if kind == "cross" else state_schema_vit(cfg)
    state = make_state(schema, seed=sseed, init="test")
    img, labels = make_inputs(cfg, batch, seed=iseed)
    return kind, cfg, state, img, labels


def grad_probes(name: str, grad: torch.Tensor, idx: int):
    """Two deterministic linear functionals + the norm of a gradient tensor."""
    g = grad.detach().double().flatten()
    i = torch.arange(g.numel(), dtype=torch.float64)
    p1 = torch.cos(0.37 * i + 0.11 * idx)
    p2 = torch.sin(0.013 * i * (1 + idx % 7) + 0.5)
    return torch.stack([g.norm(), (g * p1).sum(), (g * p2).sum()])
