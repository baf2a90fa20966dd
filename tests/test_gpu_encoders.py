"""GPU parity of the CNN-stem models' transformer cores (`ViT`, /root/reference/model.py; `ViT3D`,
/root/reference/modelv2.py) and of the kernels added for them, through the C ABI, against the fp64 CPU oracle
(oracle/encoders.py) and the golden vectors frozen from the real reference.

Tolerances (north_star, bf16 mode): logits <= 2e-2 relative L2, gradients <= 3e-2 on the concatenated vector and
<= 6e-2 per tensor of non-negligible norm; index maps (token order, patch feature order) bit-exact."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import encoders as E       # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
BF = torch.bfloat16


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _module(name):
    from cavit.encoders import ViT, ViT3D
    kind, _, ctor, B, M, *_ = E.ENC_CASES[name]
    cfg = E.enc_config(name)
    return ViT(cfg) if kind == "cnnvit" else ViT3D({}, 1e-4, 0.0, M, cfg, **ctor)


# ------------------------------------------------------------------------------------------------ kernels
def test_tokens_from_channels_roundtrip_bit_exact():
    from cavit import ops
    B, C, S = 3, 128, 37
    feat = torch.randn(B, C, S, device="cuda")
    cls, pos = torch.randn(C, device="cuda"), torch.randn(S + 1, C, device="cuda")
    X = torch.empty(B, S + 1, C, device="cuda")
    ops.tokens_from_channels(feat, cls, pos, X, B=B, C_=C, S=S)
    ref = torch.cat(((cls + pos[0]).expand(B, 1, C), feat.transpose(1, 2) + pos[1:]), dim=1)
    assert torch.equal(X, ref)
    back = torch.empty_like(feat)
    dX = torch.randn(B, S + 1, C, device="cuda")
    ops.tokens_to_channels(dX, back, B=B, C_=C, S=S)
    assert torch.equal(back, dX[:, 1:].transpose(1, 2).contiguous())


def test_conv_patch_rows_index_map_bit_exact():
    from cavit import ops
    M, B, Cin, dims, grid = 2, 3, 4, (4, 6, 4), (2, 3, 2)
    feat = torch.randn(M * B, Cin, *dims, device="cuda")
    An, Bn, Cn = (dims[i] // grid[i] for i in range(3))
    P = Cin * grid[0] * grid[1] * grid[2]
    rows = torch.empty(B * M * An * Bn * Cn, P, dtype=BF, device="cuda")
    ops.conv_patch_rows(feat, rows, M=M, B=B, Cin=Cin, dims=dims, grid=grid)
    # reference: exactly what Conv3d(kernel = stride = grid) contracts against weight.flatten(1), in conv output order
    v = feat.view(M, B, Cin, An, grid[0], Bn, grid[1], Cn, grid[2]).permute(1, 0, 3, 5, 7, 2, 4, 6, 8)
    ref = v.reshape(B * M * An * Bn * Cn, P).to(BF)
    assert torch.equal(rows, ref)
    w = torch.randn(5, Cin, *grid, device="cuda")
    y = F.conv3d(feat.to(BF).float(), w, stride=grid)             # [M*B, 5, An, Bn, Cn]
    y = y.flatten(2).transpose(1, 2).reshape(M, B, -1, 5).transpose(0, 1).reshape(-1, 5)
    assert torch.allclose(rows.float() @ w.flatten(1).t(), y, atol=1e-4, rtol=1e-4)
    back = torch.full_like(feat, 7.0)
    ops.conv_patch_rows_bwd(rows, back, M=M, B=B, Cin=Cin, dims=dims, grid=grid)
    assert torch.equal(back, feat.to(BF).float())


def test_bce_head_matches_torch():
    from cavit import ops
    B, C = 37, 128
    x = torch.randn(B, C, device="cuda", requires_grad=True)
    w = torch.randn(1, C, device="cuda", requires_grad=True)
    b0 = torch.randn(1, device="cuda", requires_grad=True)
    y = torch.randint(0, 2, (B,), device="cuda").float()
    z = F.linear(x, w, b0).squeeze(-1)
    loss = F.binary_cross_entropy_with_logits(z, y)
    (3.0 * loss).backward()
    logits, l = torch.empty(B, device="cuda"), torch.empty(1, device="cuda")
    ops.bce_head_fwd(x.detach(), w.detach().view(-1), b0.detach(), y, logits, l, B=B, C_=C)
    assert torch.allclose(logits, z.detach(), atol=1e-4, rtol=1e-5) and abs(float(l) - float(loss)) < 1e-5
    dx, dw, db = torch.empty(B, C, device="cuda"), torch.empty(C, device="cuda"), torch.empty(1, device="cuda")
    ops.bce_head_bwd(x.detach(), w.detach().view(-1), y, logits, dx, dw, db, B=B, C_=C, loss_scale=1.5,
                     loss_scale_dev=torch.full((1,), 2.0, device="cuda"))
    assert rel(dx, x.grad) < 1e-5 and rel(dw, w.grad.view(-1)) < 1e-5 and rel(db, b0.grad) < 1e-5


def test_layernorm_fp32_output_and_fp32_gradient_input():
    from cavit import ops
    T, C = 300, 192
    x = torch.randn(1, T, C, device="cuda", requires_grad=True)
    g = (1 + 0.1 * torch.randn(1, C, device="cuda")).requires_grad_(True)
    b = (0.1 * torch.randn(1, C, device="cuda")).requires_grad_(True)
    y = F.layer_norm(x, (C,), g[0], b[0], 1e-6)
    dy = torch.randn_like(y)
    y.backward(dy)
    y32, yb = torch.empty(1, T, C, device="cuda"), torch.empty(1, T, C, dtype=BF, device="cuda")
    mean, rstd = torch.empty(1, T, device="cuda"), torch.empty(1, T, device="cuda")
    ops.ln_fwd(x.detach(), g.detach(), b.detach(), yb, mean, rstd, rows_per_group=T, groups=1, C=C, eps=1e-6, y_f32=y32)
    assert torch.allclose(y32, y.detach(), atol=2e-5, rtol=1e-5) and torch.equal(yb, y32.to(BF))
    ws = ops.ln_bwd_workspace(1, C, "cuda")
    dx, dxb = torch.empty(1, T, C, device="cuda"), torch.empty(1, T, C, dtype=BF, device="cuda")
    dg, db, dcol = torch.empty(1, C, device="cuda"), torch.empty(1, C, device="cuda"), torch.empty(1, C, device="cuda")
    ops.ln_bwd(dy, x.detach(), mean, rstd, g.detach(), dx, dg, db, ws, rows_per_group=T, groups=1, C=C, dx_bf16=dxb, dcol=dcol)
    assert rel(dx, x.grad) < 1e-5 and rel(dg, g.grad) < 1e-5 and rel(db, b.grad) < 1e-5
    assert rel(dcol, x.grad.sum(dim=1)) < 1e-4 and torch.equal(dxb, dx.to(BF))


def test_gemm_relu_epilogues():
    from cavit import ops
    from cavit._abi import EPI_BIAS_RELU, EPI_RELU_BWD
    G, T, K, N = 1, 517, 128, 512
    x = torch.randn(G, T, K, device="cuda").to(BF)
    w = (torch.randn(G, N, K, device="cuda") / K ** 0.5).to(BF)
    bias = torch.randn(G, N, device="cuda")
    h = torch.empty(G, T, N, dtype=BF, device="cuda")
    ops.linear_fwd(x, w, h, epi=EPI_BIAS_RELU, bias=bias)
    ref = torch.relu(x.float() @ w.float().transpose(1, 2) + bias[:, None])
    assert rel(h, ref) < 5e-3 and bool(((h > 0) == (ref.to(BF) > 0)).float().mean() > 0.999)
    dy = torch.randn(G, T, K, device="cuda").to(BF)
    du = torch.empty(G, T, N, dtype=BF, device="cuda")
    w2 = (torch.randn(G, K, N, device="cuda") / N ** 0.5).to(BF)        # Linear(N -> K): dU = dY W2 * [h > 0]
    ops.linear_dgrad(dy, w2, du, epi=EPI_RELU_BWD, aux=h)
    ref = (dy.float() @ w2.float()) * (h > 0)
    assert rel(du, ref) < 5e-3
    from cavit import _abi
    assert _abi.device_status() == 0


def test_fused_adam_matches_torch_adam():
    from cavit import ops
    n = 4096 + 64
    p0 = torch.randn(n)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=3e-3, weight_decay=5e-4)   # the reference's optimiser (model_cross.py:277)
    p = p0.cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    pb = torch.empty(n, dtype=BF, device="cuda")
    gen = torch.Generator().manual_seed(3)
    for step in range(1, 6):
        g = torch.randn(n, generator=gen)
        p_ref.grad = g.clone()
        opt.step()
        ops.adam_step(p, g.cuda(), m, v, pb, lr=3e-3, weight_decay=5e-4, step=step)
    assert float((p.cpu() - p_ref.detach()).abs().max()) < 1e-6
    assert torch.equal(pb, p.to(BF))


# ------------------------------------------------------------------------------------------------ models
@pytest.mark.parametrize("name", list(E.ENC_CASES))
def test_encoder_model_matches_oracle_and_golden(name):
    from cavit import _abi
    kind = E.ENC_CASES[name][0]
    rec = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    state = E.make_state_generic(rec["schema"], E.ENC_CASES[name][5])
    x, labels = E.enc_inputs(name)
    ref_logits, ref_loss, ref_grads = E.enc_forward_backward(name, state, x, labels, torch.float64)
    model = _module(name)
    model.load_state_dict(state, strict=True)
    model = model.cuda().train()
    outs = []
    for it in range(4):       # two eager steps, graph capture, graph replay: all must agree
        model.load_state_dict(state, strict=True)     # BatchNorm running buffers back to the start
        model.zero_grad(set_to_none=True)
        n0 = _abi.launch_count()
        logits, loss = model(x.cuda(), labels.cuda())
        loss.backward()
        torch.cuda.synchronize()
        assert _abi.device_status() == 0
        outs.append((logits.detach().clone(), loss.detach().clone(),
                     {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
        if it == 0:
            assert _abi.launch_count() - n0 > 20
    logits, loss, grads = outs[0]
    if kind == "vit3d":
        # ReLU FFN: replay the activation pattern the CUDA path used (units whose pre-activation is within bf16 rounding of
        # zero flip, and a flipped unit changes its gradient entry by 100 % — see oracle/encoders.py: vit3d_core)
        eng = model._engine_obj
        B, N = x.shape[0], eng.N
        masks = [(eng.a["h"][l][0] > 0).view(B, N, -1).cpu() for l in range(eng.L)]
        report = []
        ref_logits, ref_loss, ref_grads = E.enc_forward_backward(name, state, x, labels, torch.float64, masks, report)
        for frac, worst in report:
            assert frac < 0.03 and worst < 0.1, report    # few units flip, and only ones with |u| << std(u)
    for l2, s2, g2 in outs[1:]:
        assert rel(l2, logits) < 1e-3 and abs(float(s2) - float(loss)) < 1e-3
        for k in grads:
            assert float((g2[k] - grads[k]).norm()) <= 2e-2 * float(grads[k].norm()) + 1e-6, k
    # ViT's single logit is a 128-term dot product that cancels to ~0.04 in this case: its error is measured against the
    # scale of the terms (max(1, |z|)), not against the cancelled value
    scale = max(1.0, float(ref_logits.abs().max()))
    assert float((logits.double().cpu() - ref_logits).abs().max()) < 2e-2 * scale
    assert float((logits.double().cpu() - rec["logits64"]).abs().max()) < 2e-2 * scale   # golden = the real reference's output
    if kind == "vit3d":
        assert rel(logits, ref_logits) < 2e-2
    assert abs(float(loss) - float(rec["loss64"])) < 2e-2 * max(1.0, abs(float(rec["loss64"])))
    tot_err, tot_ref = 0.0, 0.0
    gmax = max(float(g.norm()) for g in ref_grads.values())
    for k, g in ref_grads.items():
        d = grads[k].double().cpu() - g
        tot_err += float(d.norm()) ** 2
        tot_ref += float(g.norm()) ** 2
        if float(g.norm()) > 1e-2 * gmax:
            assert float(d.norm()) / float(g.norm()) < 6e-2, (k, float(d.norm()) / float(g.norm()))
    assert (tot_err / tot_ref) ** 0.5 < 3e-2, (tot_err / tot_ref) ** 0.5


def test_encoder_inference_paths():
    name = "cnnvit_small"
    rec = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    state = E.make_state_generic(rec["schema"], E.ENC_CASES[name][5])
    x, labels = E.enc_inputs(name)
    model = _module(name)
    model.load_state_dict(state)
    model = model.cuda().eval()
    with torch.no_grad():
        z = model(x.cuda())                       # label=None: logits only (model.py:281-282)
        z2, loss = model(x.cuda(), labels.cuda())
    assert z.shape == (x.shape[0],) and torch.equal(z, z2)
    assert float((z.double().cpu() - rec["logits64"]).abs().max()) < 2e-2
    name = "vit3d_small"
    rec = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    state = E.make_state_generic(rec["schema"], E.ENC_CASES[name][5])
    x, labels = E.enc_inputs(name)
    model = _module(name)
    model.load_state_dict(state)
    model = model.cuda().eval()               # BatchNorm running statistics
    with torch.no_grad():
        logits, loss = model(x.cuda(), labels.cuda())
    p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in state.items()}
    ref_logits, ref_loss = E.vit3d_forward(p64, x.double(), labels, E.enc_config(name), 0.1, training=False)
    assert rel(logits, ref_logits) < 2e-2 and abs(float(loss) - float(ref_loss)) < 2e-2


def test_vit3d_long_token_axis_uses_generic_attention_kernels():
    """ViT3D with 4 modalities x 64 tokens + CLS = 257 tokens (> 256: the tiled attention kernels with the three-warp MMA
    issue of the backward), hidden 256 / 4 heads, against the fp64 oracle with the ReLU pattern replayed."""
    from types import SimpleNamespace
    from cavit import _abi
    from cavit.encoders import ViT3D
    cfg = SimpleNamespace(hidden_dim=256, transformer=SimpleNamespace(num_heads=4, num_layers=2), img_size=(64, 64, 64))
    M, B = 4, 2
    torch.manual_seed(5)
    model = ViT3D({}, 1e-4, 0.0, M, cfg, num_classes=2, label_smoothing=0.0)
    schema = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}
    state = E.make_state_generic(schema, 41)
    model.load_state_dict(state)
    model = model.cuda().train()
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, M, 1, 64, 64, 64, generator=g)
    labels = torch.tensor([0, 1])
    logits, loss = model(x.cuda(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert _abi.device_status() == 0
    eng = model._engine_obj
    assert eng.N == 257
    masks = [(eng.a["h"][l][0] > 0).view(B, eng.N, -1).cpu() for l in range(eng.L)]
    lp = {}
    for k, v in state.items():
        if not v.is_floating_point():
            lp[k] = v
        elif k.endswith(("running_mean", "running_var")):
            lp[k] = v.double()
        else:
            lp[k] = v.double().clone().requires_grad_(True)
    report = []
    ref_logits, ref_loss = E.vit3d_forward(lp, x.double(), labels, cfg, 0.0, training=True, relu_masks=masks, mask_report=report)
    ref_loss.backward()
    assert all(frac < 0.03 for frac, _ in report), report
    assert rel(logits, ref_logits) < 2e-2 and abs(float(loss) - float(ref_loss)) < 2e-2
    tot_err = tot_ref = 0.0
    for k, p in model.named_parameters():
        gref = lp[k].grad
        tot_err += float((p.grad.double().cpu() - gref).norm()) ** 2
        tot_ref += float(gref.norm()) ** 2
    assert (tot_err / tot_ref) ** 0.5 < 4e-2, (tot_err / tot_ref) ** 0.5
