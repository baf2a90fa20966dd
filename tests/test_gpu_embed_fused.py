"""GPU parity of K-EMBED (csrc/embed.cu): the TMA-staged patch unfold fused with the embedding GEMM + positional add, and
its weight gradient, against the oracle's restatement of
`rearrange('b c (d p1) (h p2) (w p3) -> b (h w d) (p1 p2 p3 c)')` + `patch_to_embedding` + `cat(cls)` + `pos_embedding`
(/root/reference/model_cross.py:189-198, modelv3.py:125-140).

Token and feature ORDER is checked bit-exactly: with integer voxels, weights and token gradients every product and every
fp32 partial sum is an exact integer, so the kernels must reproduce the integer reference to the last bit — any misplaced
voxel, token row or feature column changes the result. The forward's voxels go up to +-1000: exact in the TF32 operand
format that kernel reads the fp32 data in (11 significant bits), not in bf16 (8 bits) — a bf16 round trip would show."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
BF = torch.bfloat16

# B, M, (D, H, W), (dp, hp, wp), C, sample_major
GEOMS = [
    (2, 4, (32, 32, 16), (16, 16, 8), 128, False),      # cross_ring4
    (3, 3, (16, 32, 16), (8, 16, 8), 128, False),       # cross_chain3 (3 modalities)
    (2, 2, (48, 32, 1), (16, 16, 1), 192, False),       # 2-D slices stored as (D, H, 1): the slice frame
    (2, 2, (16, 16, 16), (8, 8, 8), 64, False),         # C < one N tile, features of two planes a per wgrad tile
    (2, 2, (16, 32, 16), (8, 16, 8), 128, True),        # ModelVIT row order (streams concatenated per sample)
    (3, 4, (224, 224, 1), (16, 16, 1), 384, False),     # BASELINE cfg2 geometry
    (1, 4, (128, 128, 64), (16, 16, 8), 1024, False),   # cfg1
    (1, 4, (64, 64, 64), (8, 8, 8), 512, False),        # cfg5 patches (8^3), half-size volume
    (1, 4, (48, 240, 160), (16, 16, 16), 768, False),   # cfg3 patches and H x W extents (15 x 10 patches per slab)
    (20, 4, (16, 16, 16), (8, 8, 8), 64, False),        # more volumes than one brick holds, ragged last bricks
]


@pytest.fixture(scope="module", autouse=True)
def _device():
    from cavit import _abi
    _abi.require_device(0)
    yield
    assert _abi.device_status() == 0


def _layout(B, M, Np, sample_major):
    """-> (number of streams G, tokens per sequence N, function (m, b) -> (g, first row of that sequence))."""
    if sample_major:
        N = M * Np + 1
        return 1, N, lambda m, b: (0, b * N + 1 + m * Np)
    N = Np + 1
    return M, N, lambda m, b: (m, b * N + 1)


def _ref_patches(img, patch):
    from oracle.functional import patchify
    return torch.stack([patchify(img[:, m].cpu(), patch) for m in range(img.shape[1])], 1)   # [B, M, Np, P]


@pytest.mark.parametrize("B,M,dims,patch,C,sample_major", GEOMS)
def test_fused_embedding_forward_is_bit_exact_on_integers(B, M, dims, patch, C, sample_major):
    from cavit import ops
    if not ops.embed_fused_supported((B, M, 1) + dims, patch, C):
        pytest.fail("geometry expected to be supported")
    g = torch.Generator().manual_seed(sum(dims) + C)
    img = torch.randint(-1000, 1001, (B, M, 1) + dims, generator=g).float()
    P = patch[0] * patch[1] * patch[2]
    W = torch.randint(-1, 2, (C, P), generator=g).float()
    bias = torch.randint(-3, 4, (C,), generator=g).float()
    pat = _ref_patches(img, patch)
    Np = pat.shape[2]
    G, N, where = _layout(B, M, Np, sample_major)
    pos = torch.randint(-5, 6, (N, C), generator=g).float()
    tokens = torch.full((G, B * N, C), float("nan"), device=DEV)
    ops.embed_fused_fwd(img.to(DEV), W.to(DEV), bias.to(DEV), pos.to(DEV), tokens, patch_size=patch, C_=C,
                        sample_major=sample_major)
    got = tokens.cpu()
    want = torch.einsum("bmtp,cp->bmtc", pat.double(), W.double()) + bias.double()
    for m in range(M):
        for b in range(B):
            gi, r0 = where(m, b)
            prow = (1 + m * Np) if sample_major else 1
            ref = (want[b, m] + pos[prow:prow + Np].double()).float()
            assert torch.equal(got[gi, r0:r0 + Np], ref), (m, b)
    # rows the kernel must not touch (CLS rows) are still NaN
    for b in range(B):
        for gi in range(G):
            assert torch.isnan(got[gi, b * N]).all()


@pytest.mark.parametrize("B,M,dims,patch,C,sample_major", GEOMS)
def test_fused_embedding_wgrad_is_bit_exact_on_integers(B, M, dims, patch, C, sample_major):
    from cavit import ops
    g = torch.Generator().manual_seed(sum(dims) + C + 1)
    img = torch.randint(-3, 4, (B, M, 1) + dims, generator=g).float()
    pat = _ref_patches(img, patch)
    Np, P = pat.shape[2], pat.shape[3]
    G, N, where = _layout(B, M, Np, sample_major)
    dtok = torch.randint(-2, 3, (G, B * N, C), generator=g).float()          # CLS rows carry values the kernel must skip
    dW = torch.full((C, P), float("nan"), device=DEV)
    ops.embed_fused_wgrad(img.to(DEV), dtok.to(DEV).to(BF), dW, patch_size=patch, C_=C, sample_major=sample_major)
    want = torch.zeros(C, P, dtype=torch.float64)
    for m in range(M):
        for b in range(B):
            gi, r0 = where(m, b)
            want += dtok[gi, r0:r0 + Np].double().T @ pat[b, m].double()
    assert torch.equal(dW.cpu(), want.float())


def test_fused_embedding_on_real_valued_data():
    """N(0,1) voxels / Xavier-sized weights against fp64: the fused forward (fp32 data read as TF32) is at least as close as
    patchify + EPI_EMBED GEMM on bf16 copies; the weight gradient likewise; the bias gradient from d(pos) equals the column
    sum of the patch-token gradients."""
    from cavit import ops
    from cavit._abi import EPI_EMBED
    torch.manual_seed(0)
    B, M, dims, patch, C = 4, 4, (32, 64, 32), (16, 16, 8), 256
    P = patch[0] * patch[1] * patch[2]
    Np = (dims[0] // patch[0]) * (dims[1] // patch[1]) * (dims[2] // patch[2])
    N = Np + 1
    img = torch.randn((B, M, 1) + dims, device=DEV)
    W = torch.randn(C, P, device=DEV) / P ** 0.5
    bias, pos = torch.randn(C, device=DEV), torch.randn(N, C, device=DEV)
    a = torch.zeros(M, B * N, C, device=DEV)
    b = torch.zeros(M, B * N, C, device=DEV)
    ops.embed_fused_fwd(img, W, bias, pos, a, patch_size=patch, C_=C)
    patches = torch.empty(M * B * Np, P, device=DEV, dtype=BF)
    ops.patchify(img, patches, patch_size=patch)
    ops.gemm(patches, W.to(BF), b, M=M * B * Np, N=C, K=P, lda=P, ldb=P, ldo=C, epi=EPI_EMBED, bias=bias, resid=pos, ldr=C,
             embed_np=Np)
    pat = _ref_patches(img, patch).double()                                          # [B, M, Np, P]
    want = torch.einsum("bmtp,cp->mbtc", pat, W.double().cpu()) + bias.double().cpu() + pos[1:].double().cpu()
    got_f = a.view(M, B, N, C)[:, :, 1:].double().cpu()
    got_u = b.view(M, B, N, C)[:, :, 1:].double().cpu()
    err_f, err_u = float((got_f - want).abs().max()), float((got_u - want).abs().max())
    assert err_f < 4e-3 * float(want.abs().max()), (err_f, err_u)        # TF32 operands: 2^-11 relative per word
    assert err_f < 1.5 * err_u, (err_f, err_u)
    # weight gradient on real-valued token gradients
    dY = torch.randn(M, B * N, C, device=DEV)
    dW = torch.empty(C, P, device=DEV)
    ops.embed_fused_wgrad(img, dY.to(BF), dW, patch_size=patch, C_=C)
    want_w = torch.einsum("mbtc,bmtp->cp", dY.view(M, B, N, C)[:, :, 1:].double().cpu(), pat)
    assert float((dW.double().cpu() - want_w).abs().max()) < 1e-2 * float(want_w.abs().max())     # bf16 operands
    dX = torch.randn(M, B * N, C, device=DEV)
    dpos, dcls, db = torch.empty(N, C, device=DEV), torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ops.embed_param_grads(dX, dpos, dcls, M=M, B=B, N=N, C_=C)
    ops.embed_bias_grad(dpos, db, N=N, C_=C)
    want = dX.view(M, B, N, C)[:, :, 1:].double().sum((0, 1, 2))
    assert float((db.double() - want).abs().max()) < 1e-4 * float(want.abs().max())


def test_unsupported_geometry_is_reported_and_engine_falls_back():
    from cavit import _abi, ops
    assert not ops.embed_fused_supported((2, 2, 1, 8, 8, 8), (4, 4, 4), 64)           # wp = 4: no 32-byte runs to box
    img = torch.zeros(2, 2, 1, 8, 8, 8, device=DEV)
    with pytest.raises(_abi.CavitError):
        ops.embed_fused_fwd(img, torch.zeros(64, 64, device=DEV), torch.zeros(64, device=DEV),
                            torch.zeros(9, 64, device=DEV), torch.zeros(2, 18, 64, device=DEV), patch_size=(4, 4, 4), C_=64)
