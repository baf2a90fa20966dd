"""CPU, build container only: the restatement against the imported, unmodified reference
modules on fresh seeds (skipped where /root/reference is absent, e.g. on the GPU box)."""
import pytest
import torch

from oracle import functional as OF
from oracle import ref_loader
from oracle.functional import make_config
from oracle.weights import make_inputs

pytestmark = pytest.mark.skipif(ref_loader.reference_dir() is None, reason="reference tree not present")


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


@pytest.mark.parametrize("kind,kw", [
    ("cross", dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_multi_blocks=2, num_self_blocks=1,
                   patch_size=(8, 8, 8), img_size=(16, 16, 16), num_modalities=3,
                   attn_order={"0": "1", "1": "2", "2": "0"}, label_smoothing=0.05)),
    ("cross", dict(hidden_dim=64, mlp_dim=96, num_heads=1, num_multi_blocks=1, num_self_blocks=1,
                   patch_size=(8, 8, 8), img_size=(16, 16, 8), num_modalities=2,
                   attn_order={"1": "0"}, label_smoothing=0.0)),
    ("vit", dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_layers=3, patch_size=(8, 8, 8),
                 img_size=(16, 16, 16), num_modalities=2, attn_order={})),
])
def test_restatement_equals_reference_module(kind, kw):
    cfg = make_config(**kw)
    mod = ref_loader.load("model_cross" if kind == "cross" else "modelv3")
    torch.manual_seed(3)
    model = (mod.ModelCross if kind == "cross" else mod.ModelVIT)(ref_loader.to_config_dict(cfg)).double().train()
    # make biases / LN affine non-trivial
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.ndim == 1:
                p.add_(0.05 * torch.randn_like(p))
    img, labels = make_inputs(cfg, 3, seed=21, dtype=torch.float64)
    logits_r, loss_r = model(img, labels)
    loss_r.backward()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    logits, loss, grads = OF.forward_backward(state, img, labels, cfg, kind, torch.float64)
    assert rel(logits, logits_r.detach()) < 1e-12
    assert abs(float(loss) - float(loss_r)) < 1e-12
    for n, p in model.named_parameters():
        assert rel(grads[n], p.grad) < 1e-9 or float(p.grad.norm()) < 1e-13, n


def test_state_schema_matches_reference_state_dict():
    from oracle.weights import state_schema_cross, state_schema_vit
    for kind, kw in [("cross", dict(hidden_dim=128, mlp_dim=256, num_heads=2, patch_size=(8, 8, 8),
                                    img_size=(16, 16, 16), num_modalities=3,
                                    attn_order={"0": "1", "2": "0"})),
                     ("vit", dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_layers=2,
                                  patch_size=(8, 8, 8), img_size=(16, 16, 16), num_modalities=2))]:
        cfg = make_config(**kw)
        mod = ref_loader.load("model_cross" if kind == "cross" else "modelv3")
        model = (mod.ModelCross if kind == "cross" else mod.ModelVIT)(ref_loader.to_config_dict(cfg))
        schema = state_schema_cross(cfg) if kind == "cross" else state_schema_vit(cfg)
        sd = model.state_dict()
        assert list(sd.keys()) == list(schema.keys())
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(schema[k]), k
