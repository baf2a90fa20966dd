"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — parity cases at the BASELINE.json shapes.

The toy cases of oracle/cases.py finish in seconds in fp64; these are the EXACT model shapes of
BASELINE.json's configs at a reduced batch, which the fp64 reference needs minutes
for. They are therefore run ONCE in the build container through the unmodified reference
(oracle/gen_golden_full.py) and frozen into ``tests/golden/full_<case>.pt``; the GPU parity tests
rebuild the same seeded weights / inputs and compare against the frozen reference outputs.

Weights: the reference's own initialiser under ``torch.manual_seed(state_seed)`` (the drop-in's
constructors are registration-order identical, tests/test_modules_cpu.py) plus — so that bias and
LayerNorm-affine code paths carry signal at these sizes too — a seeded N(0, 0.05^2) perturbation of
every 1-D parameter (`perturb_1d`). Inputs: N(0,1) volumes, `randint` labels (oracle.weights.make_inputs).
"""
from __future__ import annotations

import torch

RING4 = {"0": "1", "1": "2", "2": "3", "3": "0"}

FULL_CASES = {
    # name: (kind, config kwargs, batch, state seed, input seed)
    # BASELINE.json configs[1]: 2-D slices 224x224, patch 16, dim 384, 6 layers (3 x 2), N = 197
    "cfg2_b2": ("cross", dict(hidden_dim=384, mlp_dim=1536, num_heads=6, num_multi_blocks=3, num_self_blocks=2,
                              patch_size=(16, 16, 1), img_size=(224, 224, 1), num_modalities=4, attn_order=RING4,
                              num_classes=2, dropout=0.0, label_smoothing=0.0), 2, 0, 1234),
    # BASELINE.json configs[0] shape (config2.py defaults): dim 1024, 2 x 2 blocks, N = 513
    "cfg1_b2": ("cross", dict(hidden_dim=1024, mlp_dim=4096, num_heads=16, num_multi_blocks=2, num_self_blocks=2,
                              patch_size=(16, 16, 8), img_size=(128, 128, 64), num_modalities=4, attn_order=RING4,
                              num_classes=2, dropout=0.0, label_smoothing=0.0), 2, 0, 1234),
    # BASELINE.json configs[2]: 240x240x160 volumes, 16^3 patches, dim 768, 12 layers (6 x 2), N = 2251
    "cfg3_b1": ("cross", dict(hidden_dim=768, mlp_dim=3072, num_heads=12, num_multi_blocks=6, num_self_blocks=2,
                              patch_size=(16, 16, 16), img_size=(240, 240, 160), num_modalities=4, attn_order=RING4,
                              num_classes=2, dropout=0.0, label_smoothing=0.0), 1, 0, 1234),
    # BASELINE.json configs[4]: 128^3 volumes, 8^3 patches, dim 512, 2 x 2 blocks, N = 4097
    "cfg5_b2": ("cross", dict(hidden_dim=512, mlp_dim=2048, num_heads=8, num_multi_blocks=2, num_self_blocks=2,
                              patch_size=(8, 8, 8), img_size=(128, 128, 128), num_modalities=4, attn_order=RING4,
                              num_classes=2, dropout=0.0, label_smoothing=0.0), 2, 0, 1234),
    # ModelVIT (modelv3.py) on the cfg2 slices: the four streams concatenated, N = 4 * 196 + 1 = 785
    "vit785_b2": ("vit", dict(hidden_dim=384, mlp_dim=1536, num_heads=6, num_layers=4, patch_size=(16, 16, 1),
                              img_size=(224, 224, 1), num_modalities=4, attn_order={}, num_classes=2, dropout=0.0,
                              label_smoothing=0.0), 2, 0, 1234),
}

GRAD_SAMPLES = 1024


def perturb_1d(named_params, seed: int):
    """In place: every 1-D parameter (biases, LayerNorm affine) += N(0, 0.05^2), seeded, in registration order."""
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for _, p in named_params:
            if p.ndim == 1:
                p.add_((0.05 * torch.randn(p.shape, generator=g, dtype=torch.float32)).to(p.dtype))


def sample_index(numel: int, idx: int) -> torch.Tensor:
    """Deterministic sample of flat positions of parameter number `idx` (all positions when the tensor is small)."""
    if numel <= GRAD_SAMPLES:
        return torch.arange(numel)
    g = torch.Generator().manual_seed(7919 * (idx + 1))
    return torch.randint(0, numel, (GRAD_SAMPLES,), generator=g)


def build_full_case(name: str, model_cls_cross, model_cls_vit, make_config):
    """-> (kind, cfg, model (fp32, CPU, seeded), img fp32, labels). `model_cls_*` are either the reference's classes
    (golden generation) or the drop-in's (GPU tests): same seed => same weights."""
    from .weights import make_inputs
    kind, kw, batch, sseed, iseed = FULL_CASES[name]
    cfg = make_config(**kw)
    torch.manual_seed(sseed)
    model = (model_cls_cross if kind == "cross" else model_cls_vit)(cfg)
    perturb_1d(model.named_parameters(), sseed)
    img, labels = make_inputs(cfg, batch, seed=iseed)
    return kind, cfg, model, img, labels
