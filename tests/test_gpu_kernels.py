"""GPU parity tests of the individual CUDA kernels, called through the C ABI (cavit.ops), against
plain PyTorch fp32/fp64 references of the same op on identical (bf16-rounded) inputs.

Tolerances: GEMM / attention outputs in bf16 mode: <= 2e-2 relative (north_star); measured values
are ~3e-3 (bf16 output rounding). fp32 outputs: <= 1e-4 relative. Index work is bit-exact."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _device():
    from cavit import _abi
    _abi.require_device(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    assert _abi.device_status() == 0


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(t):
    return t.to(torch.bfloat16)


def status_ok():
    from cavit import _abi
    st = _abi.device_status()
    assert st == 0, f"kernel-side status {st}"


GEMM_SHAPES = [
    # G, T(M), N, K
    (1, 128, 128, 64),
    (1, 256, 256, 128),
    (2, 200, 384, 192),     # ragged M, K multiple of 64
    (3, 197 * 3, 1152, 384),
    (1, 77, 136, 72),       # ragged everything (K % 64 != 0, N % 32 != 0)
    (4, 1026, 1536, 384),
    (1, 2, 256, 128),       # skinny
]


@pytest.mark.parametrize("G,T,N,K", GEMM_SHAPES)
def test_gemm_forward_layout(G, T, N, K):
    from cavit import ops
    torch.manual_seed(0)
    x = bf(torch.randn(G, T, K, device=DEV))
    w = bf(torch.randn(G, N, K, device=DEV) / math.sqrt(K))
    want = torch.einsum("gtk,gnk->gtn", x.float(), w.float())
    out = torch.full((G, T, N), float("nan"), device=DEV, dtype=torch.float32)
    ops.linear_fwd(x, w, out)
    status_ok()
    assert rel(out, want) < 1e-5
    outb = torch.empty((G, T, N), device=DEV, dtype=torch.bfloat16)
    ops.linear_fwd(x, w, outb)
    status_ok()
    assert rel(outb, want) < 5e-3


@pytest.mark.parametrize("G,T,N,K", GEMM_SHAPES)
def test_gemm_dgrad_layout(G, T, N, K):
    from cavit import ops
    torch.manual_seed(1)
    dy = bf(torch.randn(G, T, N, device=DEV))
    w = bf(torch.randn(G, N, K, device=DEV) / math.sqrt(N))
    want = torch.einsum("gtn,gnk->gtk", dy.float(), w.float())
    out = torch.full((G, T, K), float("nan"), device=DEV, dtype=torch.float32)
    ops.linear_dgrad(dy, w, out)
    status_ok()
    assert rel(out, want) < 1e-5


@pytest.mark.parametrize("G,T,N,K", GEMM_SHAPES)
def test_gemm_wgrad_layout(G, T, N, K):
    from cavit import ops
    torch.manual_seed(2)
    dy = bf(torch.randn(G, T, N, device=DEV))
    x = bf(torch.randn(G, T, K, device=DEV) / math.sqrt(T))
    want = torch.einsum("gtn,gtk->gnk", dy.float(), x.float())
    out = torch.full((G, N, K), float("nan"), device=DEV, dtype=torch.float32)
    ops.linear_wgrad(dy, x, out)
    status_ok()
    assert rel(out, want) < 1e-5
    ops.linear_wgrad(dy, x, out, accumulate=True)
    status_ok()
    assert rel(out, 2 * want) < 1e-5
    for sk in (2, 5, 64):
        out.fill_(float("nan"))
        ops.linear_wgrad(dy, x, out, split_k=sk)
        status_ok()
        assert rel(out, want) < 1e-5, sk


def test_gemm_epilogues():
    from cavit import ops
    from cavit._abi import EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID, EPI_GELU_BWD
    torch.manual_seed(3)
    G, T, N, K = 2, 300, 384, 128
    x = bf(torch.randn(G, T, K, device=DEV))
    w = bf(torch.randn(G, N, K, device=DEV) / math.sqrt(K))
    bias = torch.randn(G, N, device=DEV)
    resid = torch.randn(G, T, N, device=DEV)
    base = torch.einsum("gtk,gnk->gtn", x.float(), w.float())
    out = torch.empty(G, T, N, device=DEV)
    ops.linear_fwd(x, w, out, epi=EPI_BIAS, bias=bias)
    assert rel(out, base + bias[:, None]) < 1e-5
    ops.linear_fwd(x, w, out, epi=EPI_BIAS_RESID, bias=bias, resid=resid)
    assert rel(out, base + bias[:, None] + resid) < 1e-5
    # in-place residual (out aliases resid) is what the engine does
    r2 = resid.clone()
    ops.linear_fwd(x, w, r2, epi=EPI_BIAS_RESID, bias=bias, resid=r2)
    assert rel(r2, base + bias[:, None] + resid) < 1e-5
    h = torch.empty(G, T, N, device=DEV, dtype=torch.bfloat16)
    u = torch.empty(G, T, N, device=DEV, dtype=torch.bfloat16)
    ops.linear_fwd(x, w, h, epi=EPI_BIAS_GELU, bias=bias, aux=u)
    u_want = base + bias[:, None]
    assert rel(u, u_want) < 5e-3
    assert rel(h, torch.nn.functional.gelu(u.float())) < 5e-3
    # GELU backward epilogue on the dgrad layout: dU = (dY W) * gelu'(u)
    dy = bf(torch.randn(G, T, K, device=DEV))   # pretend fc2: dY [T, K] -> dH [T, N] with W2 [K, N]
    w2 = bf(torch.randn(G, K, N, device=DEV) / math.sqrt(K))
    du = torch.empty(G, T, N, device=DEV, dtype=torch.bfloat16)
    ops.linear_dgrad(dy, w2, du, epi=EPI_GELU_BWD, aux=u)
    uu = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).backward(torch.einsum("gtk,gkn->gtn", dy.float(), w2.float()))
    assert rel(du, uu.grad) < 5e-3
    status_ok()


@pytest.mark.parametrize("C", [128, 192, 384, 768, 1024])
def test_layernorm_fwd_bwd(C):
    from cavit import ops
    torch.manual_seed(4)
    G, R = 3, 173
    x = torch.randn(G, R, C, device=DEV) * 3 + 1
    gamma = 1 + 0.1 * torch.randn(G, C, device=DEV)
    beta = 0.1 * torch.randn(G, C, device=DEV)
    y = torch.empty(G, R, C, device=DEV, dtype=torch.bfloat16)
    mean = torch.empty(G, R, device=DEV)
    rstd = torch.empty(G, R, device=DEV)
    ops.ln_fwd(x, gamma, beta, y, mean, rstd, rows_per_group=R, groups=G, C=C)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = beta.double().requires_grad_(True)
    mu = xd.mean(-1, keepdim=True)
    var = ((xd - mu) ** 2).mean(-1, keepdim=True)
    want = (xd - mu) / torch.sqrt(var + 1e-5) * gd[:, None] + bd[:, None]
    assert rel(y, want) < 4e-3
    assert rel(mean, mu.squeeze(-1)) < 1e-5
    assert rel(rstd, 1 / torch.sqrt(var.squeeze(-1) + 1e-5)) < 1e-5
    dy = bf(torch.randn(G, R, C, device=DEV))
    dres = torch.randn(G, R, C, device=DEV)
    want.backward(dy.double())
    dx = torch.empty_like(x)
    dxb = torch.empty(G, R, C, device=DEV, dtype=torch.bfloat16)
    dg = torch.empty(G, C, device=DEV)
    db = torch.empty(G, C, device=DEV)
    ws = ops.ln_bwd_workspace(G, C, DEV)
    ops.ln_bwd(dy, x, mean, rstd, gamma, dx, dg, db, ws, rows_per_group=R, groups=G, C=C, dresid=dres, dx_bf16=dxb)
    assert rel(dx, xd.grad + dres.double()) < 1e-4
    assert rel(dxb, xd.grad + dres.double()) < 4e-3
    assert rel(dg, gd.grad) < 1e-4
    assert rel(db, bd.grad) < 1e-4
    # fused column sums of the output gradient (bias gradient of the producing Linear), repeated launches on the
    # same workspace (ticket counters return to zero), more rows than one block per group
    R2 = 5000
    x2 = torch.randn(G, R2, C, device=DEV)
    mean2, rstd2 = torch.empty(G, R2, device=DEV), torch.empty(G, R2, device=DEV)
    y2 = torch.empty(G, R2, C, device=DEV, dtype=torch.bfloat16)
    ops.ln_fwd(x2, gamma, beta, y2, mean2, rstd2, rows_per_group=R2, groups=G, C=C)
    dy2 = bf(torch.randn(G, R2, C, device=DEV))
    dres2 = torch.randn(G, R2, C, device=DEV)
    dx2 = torch.empty_like(x2)
    dcol = torch.full((G, C), float("nan"), device=DEV)
    for _ in range(3):
        ops.ln_bwd(dy2, x2, mean2, rstd2, gamma, dx2, dg, db, ws, rows_per_group=R2, groups=G, C=C, dresid=dres2, dcol=dcol)
        assert rel(dcol, dx2.double().sum(1)) < 1e-5
    x2d = x2.double().requires_grad_(True)
    torch.nn.functional.layer_norm(x2d, (C,)).mul(gamma.double()[:, None]).backward(dy2.double())
    assert rel(dx2, x2d.grad + dres2.double()) < 1e-4
    assert rel(dg, (dy2.double() * torch.nn.functional.layer_norm(x2.double(), (C,))).sum(1)) < 1e-4
    status_ok()


def test_layernorm_fusion_gather_and_scatter():
    from cavit import ops
    torch.manual_seed(5)
    M, B, N, C = 3, 2, 9, 128
    cls_src, tok_src = [0, 1, 2], [1, 2, 1]   # stream 1 donates patches to two fusions
    K = 3
    streams = torch.randn(M, B * N, C, device=DEV)
    gamma = 1 + 0.1 * torch.randn(K, C, device=DEV)
    beta = 0.1 * torch.randn(K, C, device=DEV)
    y = torch.empty(K, B * N, C, device=DEV, dtype=torch.bfloat16)
    mean = torch.empty(K, B * N, device=DEV)
    rstd = torch.empty(K, B * N, device=DEV)
    x_cls = torch.stack([streams[i].view(B, N, C)[:, 0] for i in cls_src]).contiguous()
    ops.ln_fusion_fwd(streams, x_cls, gamma, beta, y, mean, rstd, B=B, N=N, C_=C, cls_src=cls_src, tok_src=tok_src)
    sd = streams.double().requires_grad_(True)
    s4 = sd.view(M, B, N, C)
    outs = []
    for k in range(K):
        tmp = torch.cat((s4[cls_src[k]][:, 0:1], s4[tok_src[k]][:, 1:]), dim=1)
        outs.append(torch.nn.functional.layer_norm(tmp, (C,), gamma[k].double(), beta[k].double(), 1e-5))
    want = torch.stack(outs).view(K, B * N, C)
    assert rel(y, want) < 4e-3
    dy = bf(torch.randn(K, B * N, C, device=DEV))
    want.backward(dy.double())
    dstreams = torch.randn(M, B * N, C, device=DEV)
    base = dstreams.clone()
    dg = torch.empty(K, C, device=DEV)
    db = torch.empty(K, C, device=DEV)
    ws = ops.ln_bwd_workspace(K, C, DEV)
    ops.ln_fusion_bwd(dy, streams, x_cls, mean, rstd, gamma, dstreams, dg, db, ws, B=B, N=N, C_=C, cls_src=cls_src,
                      tok_src=tok_src)
    assert rel(dstreams - base, sd.grad) < 1e-4
    # extra fp32 gradient on the CLS rows (the query path of the fusion)
    dy_cls = torch.randn(K, B, C, device=DEV)
    sd.grad = None
    outs2 = []
    for k in range(K):
        tmp = torch.cat((s4[cls_src[k]][:, 0:1], s4[tok_src[k]][:, 1:]), dim=1)
        outs2.append(torch.nn.functional.layer_norm(tmp, (C,), gamma[k].double(), beta[k].double(), 1e-5))
    want2 = torch.stack(outs2).view(K, B, N, C)
    dy2 = dy.double().view(K, B, N, C).clone()
    dy2[:, :, 0] += dy_cls.double()
    want2.backward(dy2)
    dstreams2 = torch.zeros(M, B * N, C, device=DEV)
    ops.ln_fusion_bwd(dy, streams, x_cls, mean, rstd, gamma, dstreams2, dg, db, ws, B=B, N=N, C_=C, cls_src=cls_src,
                      tok_src=tok_src, dy_cls=dy_cls)
    assert rel(dstreams2, sd.grad) < 1e-4
    status_ok()


@pytest.mark.parametrize("img_size,patch", [((16, 24, 8), (8, 8, 4)), ((32, 32, 1), (16, 16, 1)), ((16, 16, 16), (16, 8, 2)),
                                            ((16, 8, 4), (8, 8, 4)), ((8, 6, 2), (4, 2, 2)), ((224, 224, 1), (16, 16, 1))])
def test_patchify_bit_exact(img_size, patch):
    from cavit import ops
    from oracle.functional import patchify
    torch.manual_seed(6)
    B, M = 2, 3
    D, H, W = img_size
    img = torch.randn(B, M, 1, D, H, W, device=DEV)
    Np = (D // patch[0]) * (H // patch[1]) * (W // patch[2])
    P = patch[0] * patch[1] * patch[2]
    out = torch.empty(M, B * Np, P, device=DEV, dtype=torch.bfloat16)
    ops.patchify(img, out, patch_size=patch)
    for m in range(M):
        want = patchify(img[:, m].cpu(), patch).to(torch.bfloat16).reshape(B * Np, P)
        assert torch.equal(out[m].cpu(), want)   # bit-exact indexing (values are the same RNE cast)
    out2 = torch.empty(B, M * Np, P, device=DEV, dtype=torch.bfloat16)
    ops.patchify(img, out2, patch_size=patch, sample_major=True)
    for m in range(M):
        want = patchify(img[:, m].cpu(), patch).to(torch.bfloat16)
        assert torch.equal(out2[:, m * Np:(m + 1) * Np].cpu(), want)
    status_ok()


def test_embed_gemm_epilogue_and_cls_rows():
    from cavit import ops
    from cavit._abi import EPI_EMBED
    torch.manual_seed(7)
    M, B, Np, P, C = 2, 3, 8, 128, 128
    N = Np + 1
    patches = bf(torch.randn(1, M * B * Np, P, device=DEV))
    w = bf(torch.randn(1, C, P, device=DEV) / math.sqrt(P))
    bias = torch.randn(C, device=DEV)
    pos = torch.randn(N, C, device=DEV)
    cls = torch.randn(C, device=DEV)
    tokens = torch.full((M, B * N, C), float("nan"), device=DEV)
    ops.gemm(patches, w, tokens, M=M * B * Np, N=C, K=P, lda=P, ldb=P, ldo=C, epi=EPI_EMBED, bias=bias, resid=pos,
             ldr=C, embed_np=Np)
    ops.cls_rows(cls, pos, tokens, M=M, B=B, N=N, C_=C)
    emb = (patches[0].float() @ w[0].float().T + bias).view(M, B, Np, C) + pos[1:]
    want = torch.cat((cls.expand(M, B, 1, C) + pos[0], emb), dim=2).reshape(M, B * N, C)
    assert rel(tokens, want) < 1e-5
    dtok = torch.randn(M, B * N, C, device=DEV)
    dpos = torch.empty(N, C, device=DEV)
    dcls = torch.empty(C, device=DEV)
    ops.embed_param_grads(dtok, dpos, dcls, M=M, B=B, N=N, C_=C)
    assert rel(dpos, dtok.view(M * B, N, C).sum(0)) < 1e-5
    assert rel(dcls, dtok.view(M * B, N, C)[:, 0].sum(0)) < 1e-5
    status_ok()


def test_colsum_cast_gather():
    from cavit import ops
    torch.manual_seed(8)
    G, R, C = 3, 1500, 384
    x = bf(torch.randn(G, R, C, device=DEV))
    out = torch.empty(G, C, device=DEV)
    ops.colsum_bf16(x, out, rows=R, C_=C, groups=G)
    assert rel(out, x.float().sum(1)) < 1e-4
    # both thread mappings (rows-per-pass for widths that leave a 256-column block partly empty, column blocks otherwise),
    # ragged row counts, a single row
    for G2, R2, C2 in [(1, 1, 384), (2, 777, 64), (1, 333, 200), (4, 2049, 512), (2, 1000, 1152), (1, 4100, 1536), (3, 50, 8)]:
        x2 = bf(torch.randn(G2, R2, C2, device=DEV))
        o2 = torch.full((G2, C2), 7.0, device=DEV)
        ops.colsum_bf16(x2, o2, rows=R2, C_=C2, groups=G2)
        assert rel(o2, x2.float().sum(1)) < 1e-4, (G2, R2, C2)
    src = torch.randn(1001, device=DEV)
    dst = torch.empty(1001, device=DEV, dtype=torch.bfloat16)
    ops.cast_bf16(src, dst)
    assert torch.equal(dst, src.to(torch.bfloat16))
    s = torch.randn(G, 5 * 9, 128, device=DEV)
    d = torch.zeros(G, 5, 128, device=DEV)
    ops.gather_rows_f32(s, d, rows=5, C_=128, groups=G, src_row_stride=9 * 128, src_gs=45 * 128, dst_row_stride=128,
                        dst_gs=5 * 128)
    assert torch.equal(d, s.view(G, 5, 9, 128)[:, :, 0])
    s2 = s.clone()
    ops.gather_rows_f32(s2, d, rows=5, C_=128, groups=G, src_row_stride=9 * 128, src_gs=45 * 128, dst_row_stride=128,
                        dst_gs=5 * 128, accumulate=True, zero_src=True)
    assert torch.equal(d, 2 * s.view(G, 5, 9, 128)[:, :, 0])
    assert float(s2.view(G, 5, 9, 128)[:, :, 0].abs().max()) == 0.0
    assert torch.equal(s2.view(G, 5, 9, 128)[:, :, 1:], s.view(G, 5, 9, 128)[:, :, 1:])
    a = torch.randn(4096, device=DEV)
    b = bf(torch.randn(4096, device=DEV))
    o = torch.empty_like(a)
    ops.add_bf16_f32(a, b, o)
    assert torch.equal(o, a + b.float())
    u = bf(torch.randn(4096, device=DEV))
    dh = bf(torch.randn(4096, device=DEV))
    du = torch.empty_like(u)
    ops.gelu_bwd_bf16(dh, u, du)
    uu = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uu).backward(dh.float())
    assert rel(du, uu.grad) < 5e-3
    tok = bf(torch.randn(6 * 10, 64, device=DEV))
    comp = torch.empty(6 * 9, 64, device=DEV, dtype=torch.bfloat16)
    ops.compact_patch_rows_bf16(tok, comp, S=6, Np=9, C_=64)
    assert torch.equal(comp.view(6, 9, 64), tok.view(6, 10, 64)[:, 1:])
    status_ok()


def test_head_loss_fwd_bwd():
    from cavit import ops
    torch.manual_seed(9)
    M, B, F, K = 4, 5, 256, 2
    h = bf(torch.randn(M, B, F, device=DEV))
    W2 = torch.randn(M, K, F, device=DEV) / math.sqrt(F)
    b2 = torch.randn(M, K, device=DEV) * 0.1
    labels = torch.randint(0, K, (B,), device=DEV)
    for smoothing in (0.0, 0.1):
        logits = torch.empty(B, K, device=DEV)
        loss = torch.empty(1, device=DEV)
        ops.head_loss_fwd(h, W2, b2, labels, logits, loss, M=M, B=B, F=F, classes=K, smoothing=smoothing)
        hd = h.double().requires_grad_(True)
        Wd = W2.double().requires_grad_(True)
        bd = b2.double().requires_grad_(True)
        lg = (torch.einsum("mbf,mkf->mbk", hd, Wd) + bd[:, None]).mean(0)
        ls = torch.nn.functional.cross_entropy(lg, labels, label_smoothing=smoothing)
        assert rel(logits, lg) < 1e-5
        assert abs(float(loss) - float(ls)) < 1e-5
        ls.backward()
        dh = torch.empty(M, B, F, device=DEV, dtype=torch.bfloat16)
        dW2 = torch.empty_like(W2)
        db2 = torch.empty_like(b2)
        ops.head_loss_bwd(h, W2, labels, logits, dh, dW2, db2, M=M, B=B, F=F, classes=K, smoothing=smoothing)
        assert rel(dh, hd.grad) < 5e-3
        assert rel(dW2, Wd.grad) < 1e-4
        assert rel(db2, bd.grad) < 1e-4
    status_ok()


def _attn_ref(qkv, B, N, H):
    """fp64 reference on the bf16-rounded packed QKV: returns O [G, B*N, C], LSE [G,B,H,N], and a
    differentiable handle."""
    G = qkv.shape[0]
    C = H * 64
    x = qkv.double().view(G, B, N, 3, H, 64).requires_grad_(True)
    q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]            # [G,B,N,H,64]
    s = torch.einsum("gbqhd,gbkhd->gbhqk", q, k) * 64 ** -0.5
    lse = torch.logsumexp(s, dim=-1)
    o = torch.einsum("gbhqk,gbkhd->gbqhd", torch.softmax(s, dim=-1), v).reshape(G, B * N, C)
    return x, o, lse


# N = 785 (ModelVIT on cfg2 slices), 2251 (cfg3: 18 key tiles), 4097 (cfg5: 33 tiles) are the BASELINE.json sequence lengths
# of the long-sequence kernels (cross-item pipelining, three-warp MMA issue); several heads per launch so that the
# persistent CTAs walk more than one work item
LONG_ATTN = [(1, 1, 785, 2), (1, 1, 2251, 2), (1, 1, 4097, 1), (2, 1, 1025, 3), (1, 2, 2251, 12)]


@pytest.mark.parametrize("G,B,N,H", [(1, 1, 128, 1), (1, 2, 64, 2), (2, 2, 197, 3), (1, 1, 513, 2), (1, 3, 9, 2),
                                     (1, 1, 257, 1)] + LONG_ATTN)
def test_attention_fwd(G, B, N, H):
    from cavit import ops
    torch.manual_seed(10)
    C = H * 64
    qkv = bf(torch.randn(G, B * N, 3 * C, device=DEV))
    out = torch.full((G, B * N, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.full((G, B, H, N), float("nan"), device=DEV)
    ops.attn_fwd(qkv, out, lse, G=G, B=B, N=N, H=H, scale=64 ** -0.5)
    status_ok()
    _, o, l = _attn_ref(qkv, B, N, H)
    assert rel(out, o) < 1e-2
    assert rel(lse, l) < 1e-3


@pytest.mark.parametrize("G,B,N,H", [(1, 1, 128, 1), (1, 2, 64, 2), (2, 2, 197, 3), (1, 1, 513, 2), (1, 3, 9, 2),
                                     # short-sequence kernel (N <= 256): one / two key tiles, one / two query halves,
                                     # ragged tails, and enough heads that every persistent CTA walks several of them
                                     (1, 2, 16, 1), (1, 2, 17, 2), (1, 1, 31, 1), (1, 2, 129, 2), (1, 1, 144, 1),
                                     (1, 1, 200, 2), (1, 2, 255, 1), (1, 1, 256, 2), (2, 40, 197, 6), (1, 50, 65, 4),
                                     (1, 1, 257, 1)] + LONG_ATTN)
def test_attention_bwd(G, B, N, H):
    from cavit import ops
    torch.manual_seed(11)
    C = H * 64
    qkv = bf(torch.randn(G, B * N, 3 * C, device=DEV))
    x, o, l = _attn_ref(qkv, B, N, H)
    dout = bf(torch.randn(G, B * N, C, device=DEV))
    o.backward(dout.double())
    want = x.grad.reshape(G, B * N, 3 * C)
    # feed the kernel the reference O / LSE so this test is independent of the forward kernel
    dqkv = torch.full((G, B * N, 3 * C), float("nan"), device=DEV, dtype=torch.bfloat16)
    delta = torch.empty(G, B, H, N, device=DEV)
    dq_acc = torch.empty(G, B * N, C, device=DEV)
    ops.attn_bwd(qkv, bf(o.detach().float()), dout, l.detach().float().contiguous(), dqkv, delta, dq_acc, G=G, B=B, N=N,
                 H=H, scale=64 ** -0.5)
    status_ok()
    for i, nm in enumerate("qkv"):
        got = dqkv.view(G, B * N, 3, C)[:, :, i]
        ref = want.view(G, B * N, 3, C)[:, :, i]
        assert rel(got, ref) < 2e-2, nm


@pytest.mark.parametrize("K,B,N,H", [(1, 1, 9, 1), (4, 3, 197, 2), (2, 2, 513, 3)])
def test_single_query_cross_attention(K, B, N, H):
    from cavit import ops
    torch.manual_seed(12)
    C = H * 64
    q = torch.randn(K, B, C, device=DEV)
    kv = bf(torch.randn(K, B * N, 2 * C, device=DEV))
    out = torch.empty(K, B, C, device=DEV)
    probs = torch.empty(K, B, H, N, device=DEV)
    ops.xattn_fwd(q, kv, out, probs, K=K, B=B, N=N, H=H, scale=64 ** -0.5)
    qd = q.double().requires_grad_(True)
    kvd = kv.double().requires_grad_(True)
    kk = kvd.view(K, B, N, 2, H, 64)
    s = torch.einsum("kbhd,kbnhd->kbhn", qd.view(K, B, H, 64), kk[:, :, :, 0]) * 64 ** -0.5
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("kbhn,kbnhd->kbhd", p, kk[:, :, :, 1]).reshape(K, B, C)
    assert rel(out, o) < 1e-4
    assert rel(probs, p) < 1e-4
    dout = torch.randn(K, B, C, device=DEV)
    o.backward(dout.double())
    dq = torch.empty(K, B, C, device=DEV)
    dkv = torch.empty(K, B * N, 2 * C, device=DEV, dtype=torch.bfloat16)
    ops.xattn_bwd(q, kv, probs, dout, dq, dkv, K=K, B=B, N=N, H=H, scale=64 ** -0.5)
    assert rel(dq, qd.grad) < 1e-4
    assert rel(dkv, kvd.grad) < 5e-3
    status_ok()


# ------------------------------------------------------------------------------------------------ folded cross attention
def _xfold_reference(X, cls_src, tok_src, lnw, lnb, Wq, bq, Wk, bk, Wv, bv, H, dm=None):
    """fp64 restatement of LayerNorm(cat(cls_i, patches_j)) -> wq/wk/wv -> single-query attention, UNFOLDED
    (/root/reference/model_cross.py:88-99, 111-112; oracle/functional.py::cross_attention). Returns o [K][B][C]."""
    Kf = len(cls_src)
    M, B, N, C = X.shape
    outs = []
    for k in range(Kf):
        seq = torch.cat((X[cls_src[k]][:, 0:1], X[tok_src[k]][:, 1:]), dim=1)
        xn = torch.nn.functional.layer_norm(seq, (C,), lnw[k], lnb[k], 1e-5)
        q = (xn[:, 0:1] @ Wq[k].T + bq[k]).reshape(B, 1, H, 64).permute(0, 2, 1, 3)
        kk = (xn @ Wk[k].T + bk[k]).reshape(B, N, H, 64).permute(0, 2, 1, 3)
        v = (xn @ Wv[k].T + bv[k]).reshape(B, N, H, 64).permute(0, 2, 1, 3)
        attn = torch.softmax((q @ kk.transpose(-2, -1)) * 64 ** -0.5, dim=-1)
        if dm is not None:
            attn = attn * dm[k]
        outs.append((attn @ v).transpose(1, 2).reshape(B, C))
    return torch.stack(outs)


@pytest.mark.parametrize("M,B,N,H,cls_src,tok_src,p_drop", [
    (2, 3, 9, 1, [0], [1], 0.0), (3, 2, 37, 2, [0, 1], [1, 2], 0.0), (4, 5, 197, 6, [0, 1, 2, 3], [1, 2, 3, 0], 0.0),
    (2, 2, 65, 3, [0, 1], [1, 1], 0.0),      # two fusions reading the SAME token stream (atomic scatter)
    (2, 2, 130, 12, [1], [0], 0.0), (2, 1, 40, 16, [0], [1], 0.0)])
@pytest.mark.parametrize("tensor_cores", [0, 1])
def test_folded_cross_attention_fwd_bwd(M, B, N, H, cls_src, tok_src, p_drop, tensor_cores):
    """tensor_cores = 0: the all-fp32 CUDA-core forward (fp32-level agreement with the fp64 unfolded restatement);
    1: the tcgen05 forward where the shape qualifies (H in {2, 6} here: C % 128 == 0) — xhat and the probabilities enter the
    two contractions as bf16, so the agreement is bf16-level; the backward (fp32 kernel either way) then runs on that forward's
    saved probabilities / statistics."""
    from cavit import ops
    prev = ops.xfold_tensor_cores(tensor_cores)
    try:
        _folded_cross_attention_case(M, B, N, H, cls_src, tok_src, p_drop, tensor_cores and (H * 64) % 128 == 0 and H <= 8)
    finally:
        ops.xfold_tensor_cores(prev)


def _folded_cross_attention_case(M, B, N, H, cls_src, tok_src, p_drop, tc):
    from cavit import ops
    torch.manual_seed(21)
    C, Kf = H * 64, len(cls_src)
    f32 = dict(device=DEV, dtype=torch.float32)
    X = torch.randn(M, B, N, C, **f32) * 3.0 + 1.5
    lnw, lnb = 1.0 + 0.2 * torch.randn(Kf, C, **f32), 0.2 * torch.randn(Kf, C, **f32)
    Wq, Wk, Wv = (torch.randn(Kf, C, C, **f32) / math.sqrt(C) for _ in range(3))
    bq, bk, bv = (0.1 * torch.randn(Kf, C, **f32) for _ in range(3))
    cls = torch.stack([X[cls_src[k]][:, 0] for k in range(Kf)]).contiguous()      # [K][B][C]
    # host-side projections in fp32 torch (the engine does them with the tcgen05 GEMM): q, q' = Wk_h^T q_h
    xn0 = torch.stack([torch.nn.functional.layer_norm(cls[k], (C,), lnw[k], lnb[k], 1e-5) for k in range(Kf)])
    q = torch.einsum("kbc,kdc->kbd", xn0, Wq) + bq[:, None]
    qp = torch.einsum("kbhd,khdc->kbhc", q.view(Kf, B, H, 64), Wk.view(Kf, H, 64, C)).contiguous()
    zhat = torch.empty(Kf, B, H, C, **f32)
    z = torch.empty(Kf, B, H, C, device=DEV, dtype=torch.bfloat16)
    probs = torch.empty(Kf, B, H, N, **f32)
    mean, rstd = torch.empty(Kf, B, N, **f32), torch.empty(Kf, B, N, **f32)
    scratch = ops.xfold_scratch(Kf, B, N, H, DEV)
    seed = torch.tensor([77], dtype=torch.int64, device=DEV)
    kw = dict(K=Kf, B=B, N=N, C_=C, H=H, cls_src=cls_src, tok_src=tok_src, scale=64 ** -0.5, p_drop=p_drop, seed=seed, site=5)
    ops.xfold_fwd(X, cls, qp, lnw, lnb, zhat, z, probs, mean, rstd, scratch, **kw)
    status_ok()
    o = torch.einsum("kbhc,khdc->kbhd", (lnw[:, None, None] * zhat + lnb[:, None, None]), Wv.view(Kf, H, 64, C)).reshape(Kf, B, C) + bv[:, None]
    dm = None
    if p_drop > 0:
        mk = torch.empty(Kf * B * H * N, dtype=torch.uint8, device=DEV)
        ops.dropout(ops.DROP_MASK, None, None, mk, n=mk.numel(), p=p_drop, seed=seed, site=5)
        dm = (mk.view(Kf, B, H, 1, N).double() / (1.0 - p_drop))
    leaves = [t.double().requires_grad_(True) for t in (X, lnw, lnb, Wq, bq, Wk, bk, Wv, bv)]
    ref = _xfold_reference(leaves[0], cls_src, tok_src, *leaves[1:], H, dm=dm)
    assert rel(o, ref) < (6e-3 if tc else 2e-5)
    assert rel(z.float(), lnw[:, None, None] * zhat + lnb[:, None, None]) < 5e-3
    # ---- backward: upstream gradient do [K][B][C]
    do = torch.randn(Kf, B, C, **f32)
    ref.backward(do.double())
    gz = torch.einsum("kbhd,khdc->kbhc", do.view(Kf, B, H, 64), Wv.view(Kf, H, 64, C)).contiguous()
    dX = torch.zeros_like(X)
    dqp = torch.empty(Kf, B, H, C, **f32)
    dgam, dbet = torch.zeros(Kf, C, **f32), torch.zeros(Kf, C, **f32)
    ops.xfold_bwd(X, cls, qp, lnw, zhat, probs, mean, rstd, gz, scratch, dX, dqp, dgam, dbet, **kw)
    status_ok()
    # remaining (host-side) pieces of the chain: q path through LayerNorm of the CLS rows, weight gradients
    dq = torch.einsum("kbhc,khdc->kbhd", dqp, Wk.view(Kf, H, 64, C)).reshape(Kf, B, C)
    xn0_leaf = xn0.detach().clone()
    cls_leaf = cls.double().requires_grad_(True)
    lw, lb = lnw.double().requires_grad_(True), lnb.double().requires_grad_(True)
    xn0_d = torch.stack([torch.nn.functional.layer_norm(cls_leaf[k], (C,), lw[k], lb[k], 1e-5) for k in range(Kf)])
    xn0_d.backward(torch.einsum("kbd,kdc->kbc", dq.double(), Wq.double()))
    want_dX = leaves[0].grad
    got = dX.double().clone()
    for k in range(Kf):
        got[cls_src[k]][:, 0] += cls_leaf.grad[k]
    gtol = 1.5e-2 if tc else 1e-4     # tensor-core forward: backward inherits the bf16-level probabilities / zhat
    assert rel(got, want_dX) < gtol
    assert rel(dgam.double() + lw.grad, leaves[1].grad) < gtol
    assert rel(dbet.double() + lb.grad, leaves[2].grad) < gtol
    assert rel(torch.einsum("kbd,kbc->kdc", dq, xn0_leaf), leaves[3].grad) < gtol            # dWq
    assert rel(torch.einsum("kbhd,kbhc->khdc", q.view(Kf, B, H, 64), dqp).reshape(Kf, C, C), leaves[5].grad) < gtol  # dWk
    assert float(leaves[6].grad.abs().max()) < 1e-9                                           # dbk == 0 analytically
    zz = lnw[:, None, None] * zhat + lnb[:, None, None]
    assert rel(torch.einsum("kbhd,kbhc->khdc", do.view(Kf, B, H, 64), zz).reshape(Kf, C, C), leaves[7].grad) < gtol  # dWv
    assert rel(do.sum(1), leaves[8].grad) < 1e-5                                              # dbv


def test_expand_and_fold_heads():
    from cavit import ops
    torch.manual_seed(22)
    G, H = 3, 3
    C = H * 64
    W = bf(torch.randn(G, C, C, device=DEV))
    E = torch.empty(G, C, H * C, device=DEV, dtype=torch.bfloat16)
    ops.expand_heads(W, E, groups=G, C_=C, H=H)
    want = torch.zeros(G, C, H, C, device=DEV, dtype=torch.bfloat16)
    for h in range(H):
        want[:, h * 64:(h + 1) * 64, h] = W[:, h * 64:(h + 1) * 64]
    assert torch.equal(E.view(G, C, H, C), want)
    dE = torch.randn(G, C, H * C, device=DEV)
    dW = torch.empty(G, C, C, device=DEV)
    ops.fold_heads(dE, dW, groups=G, C_=C, H=H)
    ref = torch.stack([torch.cat([dE.view(G, C, H, C)[g, h * 64:(h + 1) * 64, h] for h in range(H)]) for g in range(G)])
    assert torch.equal(dW, ref)
    status_ok()


def test_gather_rows_indexed_moves_cls_rows_of_several_streams():
    from cavit import ops
    torch.manual_seed(12)
    M, B, N, C = 4, 5, 9, 128
    X = torch.randn(M, B * N, C, device=DEV)
    idx = [2, 0, 3]
    buf = torch.zeros(len(idx), B, C, device=DEV)
    ops.gather_rows_f32_indexed(X, buf, rows=B, C_=C, src_row_stride=N * C, src_gs=B * N * C, src_groups=idx,
                                dst_row_stride=C, dst_gs=B * C)
    for k, m in enumerate(idx):
        assert torch.equal(buf[k], X[m].view(B, N, C)[:, 0])
    X2 = X.clone()
    add = torch.randn(len(idx), B, C, device=DEV)
    ops.gather_rows_f32_indexed(add, X2, rows=B, C_=C, src_row_stride=C, src_gs=B * C, dst_row_stride=N * C,
                                dst_gs=B * N * C, dst_groups=idx, accumulate=True)
    want = X.clone().view(M, B, N, C)
    for k, m in enumerate(idx):
        want[m, :, 0] += add[k]
    assert torch.equal(X2.view(M, B, N, C), want)
    # "move": source rows are cleared
    X3 = X.clone()
    ops.gather_rows_f32_indexed(X3, buf, rows=B, C_=C, src_row_stride=N * C, src_gs=B * N * C, src_groups=idx,
                                dst_row_stride=C, dst_gs=B * C, zero_src=True)
    for k, m in enumerate(idx):
        assert float(X3[m].view(B, N, C)[:, 0].abs().max()) == 0.0
    assert torch.equal(X3[1], X[1])
    status_ok()
