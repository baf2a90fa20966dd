"""GPU: every sub-module of the drop-in is individually callable — its own `forward` runs the same C-ABI
kernels as the fused engine — and matches its reference twin (/root/reference/model_cross.py:11-148, modelv3.py:69-88,
model.py:107-214) as restated by the fp64 oracle: outputs and gradients w.r.t. inputs and parameters, bf16-mode tolerance."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import functional as OF          # noqa: E402
from oracle.cases import build_case          # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(name):
    from cavit.modules import ModelCross, ModelVIT
    kind, cfg, state, img, labels = build_case(name)
    m = (ModelCross if kind == "cross" else ModelVIT)(cfg)
    m.load_state_dict(state)
    return kind, cfg, state, m.cuda().train()


def _leaf(state, prefix):
    return {k: v.double().clone().requires_grad_(True) for k, v in state.items() if k.startswith(prefix)}


def _compare(outs, refs, module, prefix, p64, xs, xs64, tol=2e-2):
    g = torch.Generator().manual_seed(5)
    total, total_ref = 0.0, 0.0
    for o, r in zip(outs, refs):
        assert o.shape == r.shape
        assert rel(o, r) < tol
        w = torch.randn(r.shape, generator=g, dtype=torch.float64)
        total = total + (o.double() * w.to(o.device)).sum()
        total_ref = total_ref + (r * w).sum()
    total.backward()
    total_ref.backward()
    for x, x64 in zip(xs, xs64):
        assert rel(x.grad, x64.grad) < 3e-2
    gmax = max(float(v.grad.norm()) for v in p64.values() if v.grad is not None)
    for k, p in module.named_parameters():
        ref = p64[prefix + k].grad
        if ref is None or float(ref.norm()) < 1e-3 * gmax:
            continue
        assert rel(p.grad, ref) < 5e-2, k


def test_multi_scale_block_and_its_parts_match_the_oracle():
    kind, cfg, state, model = _model("cross_ring4")
    msb = model.transformer[0]
    B, N, C, M = 2, model.pos_embedding.shape[1], cfg.hidden_dim, cfg.num_modalities
    g = torch.Generator().manual_seed(1)
    xs64 = [torch.randn(B, N, C, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(M)]
    xs = [x.detach().float().cuda().requires_grad_(True) for x in xs64]
    p64 = _leaf(state, "transformer.0.")
    outs = msb(xs)
    refs = OF.multi_scale_block(p64, "transformer.0.", xs64, cfg)
    _compare(outs, refs, msb, "transformer.0.", p64, xs, xs64)


@pytest.mark.parametrize("which", ["self_block", "ffn", "attention", "cross_block"])
def test_leaf_modules_match_the_oracle(which):
    kind, cfg, state, model = _model("cross_heads3")
    msb = model.transformer[0]
    B, N, C = 3, 7, cfg.hidden_dim
    g = torch.Generator().manual_seed(2)
    x64 = torch.randn(B, N, C, generator=g, dtype=torch.float64).requires_grad_(True)
    x = x64.detach().float().cuda().requires_grad_(True)
    if which == "self_block":
        mod, pre = msb.blocks[1][0], "transformer.0.blocks.1.0."
        p64 = _leaf(state, pre)
        ref = OF.self_attention_block(p64, pre, x64, cfg.num_heads)
    elif which == "ffn":
        mod, pre = msb.blocks[0][0].ffn.fn, "transformer.0.blocks.0.0.ffn.fn."
        p64 = _leaf(state, pre)
        ref = OF.feed_forward(p64, pre, x64)
    elif which == "attention":
        mod, pre = msb.blocks[0][0].attn.fn, "transformer.0.blocks.0.0.attn.fn."
        p64 = _leaf(state, pre)
        ref = OF.self_attention(p64, pre, x64, cfg.num_heads)
    else:
        mod, pre = msb.fusion[1], "transformer.0.fusion.1."
        p64 = _leaf(state, pre)
        ref = OF.cross_attention_block(p64, pre, x64, cfg.num_heads)
    out = mod(x)
    _compare([out], [ref], mod, pre, p64, [x], [x64])


def test_vit_transformer_stack_matches_the_oracle():
    kind, cfg, state, model = _model("vit_small")
    B, N, C = 2, 11, cfg.hidden_dim
    g = torch.Generator().manual_seed(3)
    x64 = torch.randn(B, N, C, generator=g, dtype=torch.float64).requires_grad_(True)
    x = x64.detach().float().cuda().requires_grad_(True)
    p64 = _leaf(state, "transformer.")
    ref = x64
    for l in range(cfg.num_layers):
        pre = f"transformer.layers.{l}."
        xn = OF.layer_norm(ref, p64[pre + "0.norm.weight"], p64[pre + "0.norm.bias"])
        ref = OF.self_attention(p64, pre + "0.fn.", xn, cfg.num_heads) + ref
        xn = OF.layer_norm(ref, p64[pre + "2.norm.weight"], p64[pre + "2.norm.bias"])
        ref = OF.feed_forward(p64, pre + "2.fn.", xn) + ref
    out = model.transformer(x)
    _compare([out], [ref], model.transformer, "transformer.", p64, [x], [x64])


def test_cnn_stem_vit_block_matches_plain_torch():
    """`Block` of the reference's model.py (biased q / k / v, scores / sqrt(d), LayerNorm eps 1e-6) against the same math in
    fp64 torch ops."""
    from types import SimpleNamespace
    from cavit.encoders import Block
    cfg = SimpleNamespace(hidden_size=128, transformer={"num_heads": 2, "mlp_dim": 256, "dropout_rate": 0.0,
                                                        "attention_dropout_rate": 0.0, "num_layers": 1})
    torch.manual_seed(7)
    blk = Block(cfg)
    with torch.no_grad():
        for p in blk.parameters():
            if p.ndim == 1:
                p.add_(0.1 * torch.randn_like(p))
    ref_blk = Block(cfg).double()
    ref_blk.load_state_dict({k: v.double() for k, v in blk.state_dict().items()})
    blk = blk.cuda().train()
    x64 = torch.randn(2, 9, 128, dtype=torch.float64).requires_grad_(True)
    x = x64.detach().float().cuda().requires_grad_(True)

    def ref_forward(m, t):
        F = torch.nn.functional
        h = F.layer_norm(t, (128,), m.attention_norm.weight, m.attention_norm.bias, 1e-6)
        mh = m.multi_head
        q, k, v = (lin(h).view(2, 9, 2, 64).permute(0, 2, 1, 3) for lin in (mh.query, mh.key, mh.value))
        a = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1) @ v
        t = t + mh.out(a.permute(0, 2, 1, 3).reshape(2, 9, 128))
        h = F.layer_norm(t, (128,), m.ffn_norm.weight, m.ffn_norm.bias, 1e-6)
        return t + m.ffn.fc2(F.gelu(m.ffn.fc1(h)))

    out, ref = blk(x), ref_forward(ref_blk, x64)
    assert rel(out, ref) < 2e-2
    w = torch.randn(ref.shape, dtype=torch.float64)
    (out.double() * w.cuda()).sum().backward()
    (ref * w).sum().backward()
    assert rel(x.grad, x64.grad) < 3e-2
    gmax = max(float(q.grad.norm()) for q in ref_blk.parameters())
    for (k, p), (_, q) in zip(blk.named_parameters(), ref_blk.named_parameters()):
        if float(q.grad.norm()) < 1e-3 * gmax:      # key.bias: analytically zero (softmax is shift invariant)
            assert float(p.grad.norm()) < 1e-2 * gmax, k
        else:
            assert rel(p.grad, q.grad) < 5e-2, k


def test_submodule_forward_rejects_active_dropout_and_cpu_tensors():
    from cavit import _abi
    from cavit.modules import ModelCross
    kind, cfg, state, img, labels = build_case("cross_chain3")
    cfg.dropout = 0.2
    m = ModelCross(cfg).cuda().train()
    ffn = m.transformer[0].blocks[0][0].ffn
    x = torch.randn(2, 5, cfg.hidden_dim, device="cuda")
    with pytest.raises(_abi.CavitError):
        ffn(x)
    m.eval()
    assert ffn(x).shape == x.shape
    with pytest.raises(_abi.CavitError):
        ffn(x.cpu())
