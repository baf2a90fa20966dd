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


def _site_key(kind, name, cfg, calls):
    """Reference nn.Dropout module name (+ how often it has been called) -> oracle dropout site key."""
    p = name.split(".")
    if name == "dropout":                       # model_cross.py:198 (once per stream) / modelv3.py:141
        return ("embed", calls)
    if kind == "vit":
        if p[0] == "mlp_head":
            return ("head_gelu", 0) if p[1] == "3" else ("head_logits", 0)
        l = int(p[2])                            # transformer.layers.{l}.{0|2}.fn...
        if p[3] == "0":
            return ("out", l, 0)
        return ("gelu", l, 0) if p[-1] == "2" else ("fc2", l, 0)
    if p[0] == "mlp_head":                      # mlp_head.{m}.{2|4}
        return ("head_gelu", int(p[1])) if p[2] == "2" else ("head_logits", int(p[1]))
    mb = int(p[1])
    if p[2] == "blocks":                        # transformer.{mb}.blocks.{m}.{sb}.(attn.fn.to_out.1 | ffn.fn.net.{2|4})
        m, sb = int(p[3]), int(p[4])
        l = mb * cfg.num_self_blocks + sb
        if p[5] == "attn":
            return ("out", l, m)
        return ("gelu", l, m) if p[-1] == "2" else ("fc2", l, m)
    k = int(p[3])                               # transformer.{mb}.fusion.{k}.(attn.fn.{attn_drop|proj_drop} | ffn.fn.net.{2|4})
    if p[4] == "attn":
        return ("f_attn", mb, k) if p[-1] == "attn_drop" else ("f_proj", mb, k)
    return ("f_gelu", mb, k) if p[-1] == "2" else ("f_fc2", mb, k)


@pytest.mark.parametrize("kind,kw", [
    ("cross", dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_multi_blocks=2, num_self_blocks=2,
                   patch_size=(8, 8, 8), img_size=(16, 16, 16), num_modalities=3,
                   attn_order={"0": "1", "1": "2", "2": "0"}, label_smoothing=0.05, dropout=0.25)),
    ("vit", dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_layers=3, patch_size=(8, 8, 8),
                 img_size=(16, 16, 16), num_modalities=2, attn_order={}, dropout=0.25)),
])
def test_dropout_sites_match_reference(kind, kw):
    """Pins WHERE the oracle applies dropout: the multipliers the reference's nn.Dropout modules actually
    used (recovered from each module's input/output by a forward hook) are replayed in the restatement."""
    cfg = make_config(**kw)
    mod = ref_loader.load("model_cross" if kind == "cross" else "modelv3")
    torch.manual_seed(5)
    model = (mod.ModelCross if kind == "cross" else mod.ModelVIT)(ref_loader.to_config_dict(cfg)).double().train()
    dm, ncalls = {}, {}

    def hook_for(name):
        def hook(module, inp, out):
            x = inp[0].detach()
            mult = torch.where(out.detach() != 0, torch.full_like(x, 1.0 / (1.0 - module.p)), torch.zeros_like(x))
            mult = torch.where(x == 0, torch.full_like(x, 1.0 / (1.0 - module.p)), mult)   # 0 * m is 0 either way
            key = _site_key(kind, name, cfg, ncalls.get(name, 0))
            ncalls[name] = ncalls.get(name, 0) + 1
            assert key not in dm, key
            dm[key] = mult
        return hook

    for name, m in model.named_modules():
        if isinstance(m, torch.nn.Dropout):
            m.register_forward_hook(hook_for(name))
    img, labels = make_inputs(cfg, 3, seed=22, dtype=torch.float64)
    logits_r, loss_r = model(img, labels)
    loss_r.backward()
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    logits, loss, grads = OF.forward_backward(state, img, labels, cfg, kind, torch.float64, dm=dm)
    assert rel(logits, logits_r.detach()) < 1e-12
    assert abs(float(loss) - float(loss_r.detach())) < 1e-12
    for n, p in model.named_parameters():
        assert rel(grads[n], p.grad) < 1e-9 or float(p.grad.norm()) < 1e-13, n
    # dropout really was active
    assert any(float((v == 0).double().mean()) > 0.1 for v in dm.values())
