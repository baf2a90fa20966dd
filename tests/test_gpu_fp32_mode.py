"""GPU parity of the fp32-tolerance mode (`precision="fp32"`): the split-operand tensor-core GEMM, the fp32 attention
kernels, the hi / lo producing kernels, and the whole models against the fp64 oracle and the golden vectors of the
unmodified reference. north_star: ~1e-3 relative on logits and attention outputs "in fp32" — the kernel-level bars here
are far tighter (they are what leaves room for a 12-layer model): GEMM <= 5e-5, attention <= 2e-5."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module", autouse=True)
def _device():
    from cavit import _abi
    _abi.require_device(0)
    yield
    assert _abi.device_status() == 0


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def split(x):
    from cavit import ops
    t = ops.split_pair(x.shape, x.device)
    ops.cast_split(x.contiguous(), t)
    return t


def test_cast_split_planes_reconstruct_16_bits():
    from cavit import ops
    torch.manual_seed(0)
    x = torch.randn(3, 1000, 77, device=DEV) * torch.logspace(-3, 4, 77, device=DEV)
    t = split(x)
    rec = t.float() + ops.lo_of(t).float()
    assert float(((rec - x).abs() / x.abs().clamp_min(1e-30)).max()) < 2.0 ** -15
    assert torch.equal(t, x.to(torch.bfloat16))


@pytest.mark.parametrize("G,T,N,K", [(1, 128, 128, 64), (2, 200, 384, 192), (3, 591, 1152, 384), (1, 77, 136, 72),
                                     (4, 1026, 1536, 384), (1, 2, 256, 128)])
def test_split_gemm_forward_dgrad_wgrad(G, T, N, K):
    from cavit import ops
    torch.manual_seed(1)
    x = torch.randn(G, T, K, device=DEV)
    w = torch.randn(G, N, K, device=DEV) / math.sqrt(K)
    dy = torch.randn(G, T, N, device=DEV)
    xs, ws_, dys = split(x), split(w), split(dy)
    out = torch.full((G, T, N), float("nan"), device=DEV)
    ops.linear_fwd(xs, ws_, out)
    assert rel(out, torch.einsum("gtk,gnk->gtn", x.double(), w.double())) < 5e-5
    if K % 8 == 0 and N % 8 == 0:
        dx = torch.full((G, T, K), float("nan"), device=DEV)
        ops.linear_dgrad(dys, ws_, dx)
        assert rel(dx, torch.einsum("gtn,gnk->gtk", dy.double(), w.double())) < 5e-5
        dw = torch.full((G, N, K), float("nan"), device=DEV)
        ops.linear_wgrad(dys, xs, dw, split_k=3 if T >= 512 else 1)
        assert rel(dw, torch.einsum("gtn,gtk->gnk", dy.double(), x.double())) < 5e-5


def test_split_gemm_needs_both_operands_split():
    from cavit import _abi, ops
    x = torch.randn(1, 128, 64, device=DEV)
    w = torch.randn(1, 128, 64, device=DEV)
    out = torch.empty(1, 128, 128, device=DEV)
    with pytest.raises(_abi.CavitError):
        ops.linear_fwd(split(x), w.to(torch.bfloat16), out)


def test_gelu_split_forward_backward():
    from cavit import ops
    torch.manual_seed(2)
    u = (torch.randn(4, 333, 256, device=DEV) * 3).contiguous()
    dh = torch.randn_like(u)
    h = ops.split_pair(u.shape, DEV)
    h32 = torch.empty_like(u)
    ops.gelu_split(u, h=h, h32=h32)
    ud = u.double().requires_grad_(True)
    ref = torch.nn.functional.gelu(ud)
    ref.backward(dh.double())
    assert rel(h32, ref) < 1e-6
    assert rel(h.float() + ops.lo_of(h).float(), ref) < 2e-5
    du = ops.split_pair(u.shape, DEV)
    ops.gelu_bwd_split(dh, u, du)
    assert rel(du.float() + ops.lo_of(du).float(), ud.grad) < 2e-5


@pytest.mark.parametrize("C", [128, 384, 768, 1024])
def test_layernorm_split_forward_backward(C):
    from cavit import ops
    torch.manual_seed(3)
    G, T = 2, 301
    x = (torch.randn(G, T, C, device=DEV) * 2 + 0.5).contiguous()
    gamma, beta = torch.randn(G, C, device=DEV) * 0.2 + 1, torch.randn(G, C, device=DEV) * 0.1
    y = ops.split_pair((G, T, C), DEV)
    mean, rstd = torch.empty(G, T, device=DEV), torch.empty(G, T, device=DEV)
    ops.ln_fwd_split(x, gamma, beta, y, mean, rstd, rows_per_group=T, groups=G, C=C)
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    mu = xd.mean(-1, keepdim=True)
    var = ((xd - mu) ** 2).mean(-1, keepdim=True)
    ref = (xd - mu) * torch.rsqrt(var + 1e-5) * gd[:, None] + bd[:, None]
    assert rel(y.float() + ops.lo_of(y).float(), ref) < 2e-5
    dy = torch.randn(G, T, C, device=DEV)
    dres = torch.randn(G, T, C, device=DEV)
    ref.backward(dy.double())
    dx = torch.empty(G, T, C, device=DEV)
    dxs = ops.split_pair((G, T, C), DEV)
    dg, db, dcol = torch.empty(G, C, device=DEV), torch.empty(G, C, device=DEV), torch.empty(G, C, device=DEV)
    ws = ops.ln_bwd_workspace(G, C, DEV)
    ops.ln_bwd_split(dy, x, mean, rstd, gamma, dx, dg, db, ws, rows_per_group=T, groups=G, C=C, dresid=dres, dx_split=dxs,
                     dcol=dcol)
    want = xd.grad + dres.double()
    assert rel(dx, want) < 1e-5 and rel(dxs.float() + ops.lo_of(dxs).float(), want) < 2e-5
    assert rel(dg, gd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5 and rel(dcol, want.sum(1)) < 1e-5


def _attn_ref(qkv, B, N, H):
    G = qkv.shape[0]
    C = H * 64
    x = qkv.double().view(G, B, N, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]
    s = torch.einsum("gbqhd,gbkhd->gbhqk", q, k) * 64 ** -0.5
    lse = torch.logsumexp(s, dim=-1)
    o = torch.einsum("gbhqk,gbkhd->gbqhd", torch.softmax(s, dim=-1), v).reshape(G, B * N, C)
    return x, o, lse


@pytest.mark.parametrize("G,B,N,H", [(1, 1, 9, 1), (1, 2, 64, 2), (2, 2, 197, 3), (1, 1, 513, 2), (1, 3, 65, 2),
                                     (1, 1, 785, 2), (1, 1, 2251, 1)])
def test_attention_f32_forward_backward(G, B, N, H):
    from cavit import ops
    torch.manual_seed(4)
    C = H * 64
    qkv = torch.randn(G, B * N, 3 * C, device=DEV)
    out = torch.full((G, B * N, C), float("nan"), device=DEV)
    lse = torch.full((G, B, H, N), float("nan"), device=DEV)
    ops.attn_fwd_f32(qkv, out, lse, G=G, B=B, N=N, H=H, scale=64 ** -0.5)
    x, o, l = _attn_ref(qkv, B, N, H)
    assert rel(out, o) < 2e-5 and rel(lse, l) < 1e-5
    dout = torch.randn(G, B * N, C, device=DEV)
    o.backward(dout.double())
    dqkv = torch.full((G, B * N, 3 * C), float("nan"), device=DEV)
    delta = torch.empty(G, B, H, N, device=DEV)
    ops.attn_bwd_f32(qkv, out, dout, lse, dqkv, delta, G=G, B=B, N=N, H=H, scale=64 ** -0.5)
    want = x.grad.reshape(G, B * N, 3 * C)
    for i, nm in enumerate("qkv"):
        assert rel(dqkv.view(G, B * N, 3, C)[:, :, i], want.view(G, B * N, 3, C)[:, :, i]) < 2e-5, nm


def test_colsum_and_patchify_split():
    from cavit import ops
    from oracle.functional import patchify as ref_patchify
    torch.manual_seed(5)
    x = torch.randn(3, 1000, 384, device=DEV)
    out = torch.empty(3, 384, device=DEV)
    ops.colsum_bf16(split(x), out, rows=1000, C_=384, groups=3)
    assert rel(out, x.double().sum(1)) < 2e-5
    img = (torch.randn(2, 3, 1, 16, 32, 16, device=DEV) * 1000 + 2000).clamp_min(0)
    P, Np = 8 * 16 * 8, 2 * 2 * 2
    patches = ops.split_pair((3 * 2 * Np, P), DEV)
    ops.patchify(img, patches, patch_size=(8, 16, 8))
    want = torch.stack([ref_patchify(img[:, m].cpu(), (8, 16, 8)) for m in range(3)]).reshape(3 * 2 * Np, P)
    got = (patches.float() + ops.lo_of(patches).float()).cpu()
    assert torch.equal(patches.cpu(), want.to(torch.bfloat16))                 # hi plane: exactly the bf16-mode patches
    assert float(((got - want).abs() / want.abs().clamp_min(1.0)).max()) < 2.0 ** -15


def test_head_loss_f32():
    from cavit import ops
    torch.manual_seed(6)
    M, B, F, classes = 3, 5, 256, 2
    h = torch.randn(M, B, F, device=DEV)
    W2, b2 = torch.randn(M, classes, F, device=DEV) * 0.1, torch.randn(M, classes, device=DEV) * 0.1
    labels = torch.randint(0, classes, (B,), device=DEV)
    logits, loss = torch.empty(B, classes, device=DEV), torch.empty(1, device=DEV)
    ops.head_loss_fwd_f32(h, W2, b2, labels, logits, loss, M=M, B=B, F=F, classes=classes, smoothing=0.1)
    hd, Wd, bd = h.double().requires_grad_(True), W2.double().requires_grad_(True), b2.double().requires_grad_(True)
    ref = (torch.einsum("mbf,mkf->mbk", hd, Wd) + bd[:, None]).mean(0)
    ref_loss = torch.nn.functional.cross_entropy(ref, labels, label_smoothing=0.1)
    ref_loss.backward()
    assert rel(logits, ref) < 1e-5 and abs(float(loss) - float(ref_loss)) < 1e-6
    dh, dW, db = torch.empty_like(h), torch.empty_like(W2), torch.empty_like(b2)
    ops.head_loss_bwd_f32(h, W2, labels, logits, dh, dW, db, M=M, B=B, F=F, classes=classes, smoothing=0.1)
    assert rel(dh, hd.grad) < 1e-5 and rel(dW, Wd.grad) < 1e-5 and rel(db, bd.grad) < 1e-5


def _run_model(name, mri_like=False):
    from cavit import _abi
    from cavit.modules import ModelCross, ModelVIT
    from oracle import functional as OF
    from oracle.cases import CASES, build_case
    from oracle.weights import make_inputs
    kind, cfg, state, img, labels = build_case(name)
    if mri_like:
        img, labels = make_inputs(cfg, CASES[name][2], seed=CASES[name][4], mri_like=True)
    model = (ModelCross if kind == "cross" else ModelVIT)(cfg).set_precision("fp32")
    model.load_state_dict(state, strict=True)
    model = model.cuda().train()
    outs = []
    for _ in range(4):      # eager, eager, graph capture, replay: all must agree
        for p in model.parameters():
            p.grad = None
        logits, loss = model(img.cuda(), labels.cuda())
        loss.backward()
        outs.append((logits.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert _abi.device_status() == 0
    ref_logits, ref_loss, ref_grads = OF.forward_backward(state, img, labels, cfg, kind, torch.float64)
    return kind, cfg, outs, loss, ref_logits, ref_loss, ref_grads


@pytest.mark.parametrize("name", ["cross_ring4", "cross_chain3", "cross_heads3", "cross_noattn_h1", "vit_small"])
def test_fp32_mode_model_matches_oracle_and_golden(name):
    kind, cfg, outs, loss, ref_logits, ref_loss, ref_grads = _run_model(name)
    rec = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
    gmax = max(float(g.norm()) for g in ref_grads.values())
    for logits, grads in (outs[0], outs[-1]):       # first eager run and the graph replay
        assert rel(logits, ref_logits) < 1e-3 and rel(logits, rec["logits64"]) < 1e-3
        num = den = 0.0
        for k, g in ref_grads.items():
            d = grads[k].double().cpu() - g
            num += float(d.norm()) ** 2
            den += float(g.norm()) ** 2
            if float(g.norm()) > 1e-3 * gmax:
                assert float(d.norm()) / float(g.norm()) < 5e-3, k
            else:
                assert float(d.norm()) < 1e-4 * gmax + 1e-7, k
        assert (num / den) ** 0.5 < 2e-3
    assert abs(float(loss) - float(ref_loss)) < 1e-4


def test_fp32_mode_raw_mri_intensities():
    """Raw MRI intensities (dataset_ucsf.py:81-89: no normalisation), where the bf16 mode needs 3.5e-2 / 4e-2
    (tests/test_gpu_model.py): the fp32 mode holds the fp32 bar."""
    kind, cfg, outs, loss, ref_logits, ref_loss, ref_grads = _run_model("cross_ring4", mri_like=True)
    logits, grads = outs[-1]
    assert rel(logits, ref_logits) < 1e-3, rel(logits, ref_logits)
    num = sum(float((grads[k].double().cpu() - g).norm()) ** 2 for k, g in ref_grads.items())
    den = sum(float(g.norm()) ** 2 for g in ref_grads.values())
    assert (num / den) ** 0.5 < 3e-3, (num / den) ** 0.5


def test_fp32_mode_rejects_dropout_and_encoder_kinds():
    from cavit import _abi
    from cavit.modules import ModelCross
    from oracle.cases import build_case
    kind, cfg, state, img, labels = build_case("cross_chain3")
    cfg.dropout = 0.1
    m = ModelCross(cfg).set_precision("fp32").cuda().train()
    with pytest.raises(_abi.CavitError):
        m(img.cuda(), labels.cuda())
    m.eval()                                     # dropout inactive: fine
    with torch.no_grad():
        m(img.cuda(), labels.cuda())
