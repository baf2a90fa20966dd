"""Stand-alone forwards of the sub-modules (PreNorm / LayerNorm, FeedForward, Attention, CrossAttention and the blocks built
from them) routed through the same C-ABI kernels as the whole-model engine.

The engine (cavit/engine.py) is the fast path: it fuses across modules, keeps activations in flat pre-planned buffers
and replays CUDA graphs. These functions exist so that every sub-module of the drop-in remains individually callable and
testable against its reference twin (`PreNorm`, `FeedForward`, `Attention`, `SelfAttentionBlock`, `CrossAttention`,
`CrossAttentionBlock`, `MultiScaleBlock`, `Transformer`; /root/reference/model_cross.py:11-148, modelv3.py:18-88):
fp32 tensors in, fp32 tensors out, one `torch.autograd.Function` per leaf whose forward and backward are sequences of
`cavit.ops` launches (tcgen05 GEMMs with fused epilogues, LayerNorm, flash attention, single-query cross attention).
Dropout modules inside them are honoured only as the identity (p = 0 or eval mode); training with dropout > 0 goes through
the top-level model. There is no PyTorch implementation behind any of it: without the library or a B200 they raise.
"""
from __future__ import annotations

import torch

from . import _abi, ops
from ._abi import EPI_BIAS, EPI_BIAS_GELU, EPI_GELU_BWD, EPI_NONE

BF16, F32 = torch.bfloat16, torch.float32


def _check(x: torch.Tensor, C: int, what: str):
    if not x.is_cuda or x.dtype != F32:
        raise _abi.CavitError(f"{what}: float32 CUDA tensor expected (cavit has no CPU path)")
    if x.shape[-1] != C:
        raise _abi.CavitError(f"{what}: last dimension {x.shape[-1]} != {C}")
    if C % 64:
        raise _abi.CavitError(f"{what}: hidden size must be a multiple of 64")
    _abi.require_device(x.device.index)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    out = torch.empty(t.shape, dtype=BF16, device=t.device)
    ops.cast_bf16(t.contiguous(), out)
    return out


def _no_dropout(module_training: bool, p: float, what: str):
    if module_training and p > 0.0:
        raise _abi.CavitError(f"{what}: stand-alone sub-module forwards run with dropout as the identity (p = 0 or eval()); "
                              "train with dropout through the top-level model")


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        C = x.shape[-1]
        xr = x.contiguous().view(-1, C)
        T = xr.shape[0]
        y = torch.empty_like(xr)
        mean, rstd = torch.empty(T, dtype=F32, device=x.device), torch.empty(T, dtype=F32, device=x.device)
        with torch.cuda.device(x.device):
            ops.ln_fwd(xr, weight.detach().contiguous(), bias.detach().contiguous(), None, mean, rstd, rows_per_group=T, groups=1,
                       C=C, eps=eps, y_f32=y)
        ctx.save_for_backward(xr, weight, mean, rstd)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        xr, weight, mean, rstd = ctx.saved_tensors
        T, C = xr.shape
        dx = torch.empty_like(xr)
        dg, db = torch.empty(C, dtype=F32, device=xr.device), torch.empty(C, dtype=F32, device=xr.device)
        with torch.cuda.device(xr.device):
            ws = ops.ln_bwd_workspace(1, C, xr.device)
            ops.ln_bwd(dy.contiguous().view(T, C), xr, mean, rstd, weight.detach().contiguous(), dx, dg, db, ws, rows_per_group=T,
                       groups=1, C=C)
        return dx.view(dy.shape), dg, db, None


def layer_norm(x, weight, bias, eps=1e-5):
    """nn.LayerNorm(hidden_dim) of PreNorm (/root/reference/model_cross.py:14-17)."""
    _check(x, weight.shape[0], "layer_norm")
    return _LayerNorm.apply(x, weight, bias, eps)


class _Linear(torch.autograd.Function):
    """y = x W^T (+ b): fp32 in / out, bf16 operands on the tensor cores."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        K = x.shape[-1]
        N = weight.shape[0]
        xr = x.contiguous().view(-1, K)
        T = xr.shape[0]
        with torch.cuda.device(x.device):
            xb, wb = _bf16(xr), _bf16(weight.detach())
            y = torch.empty(T, N, dtype=F32, device=x.device)
            ops.gemm(xb, wb, y, M=T, N=N, K=K, lda=K, ldb=K, ldo=N, epi=EPI_BIAS if bias is not None else EPI_NONE,
                     bias=None if bias is None else bias.detach().contiguous())
        ctx.save_for_backward(xb, wb)
        ctx.has_bias = bias is not None
        return y.view(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, dy):
        xb, wb = ctx.saved_tensors
        T, K = xb.shape
        N = wb.shape[0]
        dev = xb.device
        with torch.cuda.device(dev):
            dyb = _bf16(dy.contiguous().view(T, N))
            dx = torch.empty(T, K, dtype=F32, device=dev)
            ops.gemm(dyb, wb, dx, M=T, N=K, K=N, b_mn=True, lda=N, ldb=K, ldo=K)
            dw = torch.empty(N, K, dtype=F32, device=dev)
            ops.gemm(dyb, xb, dw, M=N, N=K, K=T, a_mn=True, b_mn=True, lda=N, ldb=K, ldo=K)
            db = None
            if ctx.has_bias:
                db = torch.empty(N, dtype=F32, device=dev)
                ops.colsum_bf16(dyb, db, rows=T, C_=N, groups=1)
        return dx.view(dy.shape[:-1] + (K,)), dw, db


def linear(x, weight, bias=None):
    _check(x, weight.shape[1], "linear")
    return _Linear.apply(x, weight, bias)


class _FeedForward(torch.autograd.Function):
    """Linear(C, F) + bias, exact-erf GELU, Linear(F, C) + bias with the fused epilogues of the engine
    (FeedForward.net, /root/reference/model_cross.py:19-31)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        C, F = w1.shape[1], w1.shape[0]
        xr = x.contiguous().view(-1, C)
        T = xr.shape[0]
        dev = x.device
        with torch.cuda.device(dev):
            xb, w1b, w2b = _bf16(xr), _bf16(w1.detach()), _bf16(w2.detach())
            u, h = torch.empty(T, F, dtype=BF16, device=dev), torch.empty(T, F, dtype=BF16, device=dev)
            ops.gemm(xb, w1b, h, M=T, N=F, K=C, lda=C, ldb=C, ldo=F, epi=EPI_BIAS_GELU, bias=b1.detach().contiguous(), aux=u, ldaux=F)
            y = torch.empty(T, C, dtype=F32, device=dev)
            ops.gemm(h, w2b, y, M=T, N=C, K=F, lda=F, ldb=F, ldo=C, epi=EPI_BIAS, bias=b2.detach().contiguous())
        ctx.save_for_backward(xb, w1b, w2b, u, h)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        xb, w1b, w2b, u, h = ctx.saved_tensors
        T, C = xb.shape
        F = w1b.shape[0]
        dev = xb.device
        e = lambda *s: torch.empty(*s, dtype=F32, device=dev)   # noqa: E731
        with torch.cuda.device(dev):
            dyb = _bf16(dy.contiguous().view(T, C))
            du = torch.empty(T, F, dtype=BF16, device=dev)
            ops.gemm(dyb, w2b, du, M=T, N=F, K=C, b_mn=True, lda=C, ldb=F, ldo=F, epi=EPI_GELU_BWD, aux=u, ldaux=F)
            dw2, db2, dw1, db1, dx = e(C, F), e(C), e(F, C), e(F), e(T, C)
            ops.gemm(dyb, h, dw2, M=C, N=F, K=T, a_mn=True, b_mn=True, lda=C, ldb=F, ldo=F)
            ops.colsum_bf16(dyb, db2, rows=T, C_=C, groups=1)
            ops.gemm(du, w1b, dx, M=T, N=C, K=F, b_mn=True, lda=F, ldb=C, ldo=C)
            ops.gemm(du, xb, dw1, M=F, N=C, K=T, a_mn=True, b_mn=True, lda=F, ldb=C, ldo=C)
            ops.colsum_bf16(du, db1, rows=T, C_=F, groups=1)
        return dx.view(dy.shape), dw1, db1, dw2, db2


def feed_forward(x, w1, b1, w2, b2):
    _check(x, w1.shape[1], "feed_forward")
    if w1.shape[0] % 8:
        raise _abi.CavitError("feed_forward: mlp_dim must be a multiple of 8")
    return _FeedForward.apply(x, w1, b1, w2, b2)


class _SelfAttention(torch.autograd.Function):
    """to_qkv (no bias), fused flash attention over 'b n (h d) -> b h n d', optional to_out Linear
    (Attention.forward, /root/reference/model_cross.py:50-61). x: [B, N, C] fp32."""

    @staticmethod
    def forward(ctx, x, wqkv, wo, bo, heads, bqkv=None):
        B, N, C = x.shape
        T = B * N
        dev = x.device
        with torch.cuda.device(dev):
            xb, wqb = _bf16(x.contiguous().view(T, C)), _bf16(wqkv.detach())
            qkv = torch.empty(1, T, 3 * C, dtype=BF16, device=dev)
            ops.gemm(xb, wqb, qkv, M=T, N=3 * C, K=C, lda=C, ldb=C, ldo=3 * C, epi=EPI_BIAS if bqkv is not None else EPI_NONE,
                     bias=None if bqkv is None else bqkv.detach().contiguous())
            ao = torch.empty(1, T, C, dtype=BF16, device=dev)
            lse = torch.empty(1, B, heads, N, dtype=F32, device=dev)
            ops.attn_fwd(qkv, ao, lse, G=1, B=B, N=N, H=heads, scale=64 ** -0.5)
            y = torch.empty(T, C, dtype=F32, device=dev)
            if wo is not None:
                wob = _bf16(wo.detach())
                ops.gemm(ao, wob, y, M=T, N=C, K=C, lda=C, ldb=C, ldo=C, epi=EPI_BIAS, bias=bo.detach().contiguous())
            else:   # heads == 1: the reference's to_out is nn.Identity()
                wob = None
                y.copy_(ao.view(T, C))
        ctx.save_for_backward(xb, wqb, qkv, ao, lse, *( [wob] if wob is not None else []))
        ctx.dims, ctx.has_out, ctx.has_qkv_bias = (B, N, C, heads), wob is not None, bqkv is not None
        return y.view(B, N, C)

    @staticmethod
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        xb, wqb, qkv, ao, lse = saved[:5]
        B, N, C, heads = ctx.dims
        T = B * N
        dev = xb.device
        e = lambda *s: torch.empty(*s, dtype=F32, device=dev)   # noqa: E731
        dwo = dbo = None
        with torch.cuda.device(dev):
            dyb = _bf16(dy.contiguous().view(T, C))
            if ctx.has_out:
                wob = saved[5]
                dao = torch.empty(1, T, C, dtype=BF16, device=dev)
                ops.gemm(dyb, wob, dao, M=T, N=C, K=C, b_mn=True, lda=C, ldb=C, ldo=C)
                dwo, dbo = e(C, C), e(C)
                ops.gemm(dyb, ao, dwo, M=C, N=C, K=T, a_mn=True, b_mn=True, lda=C, ldb=C, ldo=C)
                ops.colsum_bf16(dyb, dbo, rows=T, C_=C, groups=1)
            else:
                dao = dyb.view(1, T, C)
            dqkv = torch.empty(1, T, 3 * C, dtype=BF16, device=dev)
            ops.attn_bwd(qkv, ao, dao, lse, dqkv, e(1, B, heads, N), e(1, T, C), G=1, B=B, N=N, H=heads, scale=64 ** -0.5)
            dx, dwq = e(T, C), e(3 * C, C)
            ops.gemm(dqkv, wqb, dx, M=T, N=C, K=3 * C, b_mn=True, lda=3 * C, ldb=C, ldo=C)
            ops.gemm(dqkv, xb, dwq, M=3 * C, N=C, K=T, a_mn=True, b_mn=True, lda=3 * C, ldb=C, ldo=C)
            dbq = None
            if ctx.has_qkv_bias:
                dbq = e(3 * C)
                ops.colsum_bf16(dqkv.view(T, 3 * C), dbq, rows=T, C_=3 * C, groups=1)
        return dx.view(B, N, C), dwq, dwo, dbo, None, dbq


def self_attention(x, wqkv, wo, bo, heads, bqkv=None):
    """wqkv: packed [3C, C] (q | k | v thirds); bqkv: optional packed bias (model.py's biased query / key / value)."""
    _check(x, wqkv.shape[1], "self_attention")
    if x.dim() != 3 or x.shape[-1] != 64 * heads:
        raise _abi.CavitError("self_attention: x must be [B, N, C] with C == 64 * heads (the kernels are specialised for head_dim 64)")
    return _SelfAttention.apply(x, wqkv, wo, bo, heads, bqkv)


class _CrossAttention(torch.autograd.Function):
    """Single-query cross attention: q from token 0, k / v from all N tokens, biased projections, output projection
    (CrossAttention.forward, /root/reference/model_cross.py:88-102). x: [B, N, C] fp32 -> [B, 1, C]."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, wp, bp, heads):
        B, N, C = x.shape
        T = B * N
        dev = x.device
        with torch.cuda.device(dev):
            xb = _bf16(x.contiguous().view(T, C))
            wqb, wpb = _bf16(wq.detach()), _bf16(wp.detach())
            wkvb = _bf16(torch.cat((wk.detach(), wv.detach()), 0))
            bkv = torch.cat((bk.detach(), bv.detach()), 0).contiguous()
            kv = torch.empty(1, T, 2 * C, dtype=BF16, device=dev)
            ops.gemm(xb, wkvb, kv, M=T, N=2 * C, K=C, lda=C, ldb=C, ldo=2 * C, epi=EPI_BIAS, bias=bkv)
            q = torch.empty(1, B, C, dtype=F32, device=dev)
            ops.gemm(xb, wqb, q, M=B, N=C, K=C, lda=N * C, ldb=C, ldo=C, epi=EPI_BIAS, bias=bq.detach().contiguous())   # CLS rows only
            o = torch.empty(1, B, C, dtype=F32, device=dev)
            probs = torch.empty(1, B, heads, N, dtype=F32, device=dev)
            ops.xattn_fwd(q, kv, o, probs, K=1, B=B, N=N, H=heads, scale=64 ** -0.5)
            ob = _bf16(o.view(B, C))
            y = torch.empty(B, C, dtype=F32, device=dev)
            ops.gemm(ob, wpb, y, M=B, N=C, K=C, lda=C, ldb=C, ldo=C, epi=EPI_BIAS, bias=bp.detach().contiguous())
        ctx.save_for_backward(xb, wqb, wkvb, wpb, kv, q, probs, ob)
        ctx.dims = (B, N, C, heads)
        return y.view(B, 1, C)

    @staticmethod
    def backward(ctx, dy):
        xb, wqb, wkvb, wpb, kv, q, probs, ob = ctx.saved_tensors
        B, N, C, heads = ctx.dims
        T = B * N
        dev = xb.device
        e = lambda *s: torch.empty(*s, dtype=F32, device=dev)   # noqa: E731
        with torch.cuda.device(dev):
            dyb = _bf16(dy.contiguous().view(B, C))
            do, dwp, dbp = e(1, B, C), e(C, C), e(C)
            ops.gemm(dyb, wpb, do, M=B, N=C, K=C, b_mn=True, lda=C, ldb=C, ldo=C)
            ops.gemm(dyb, ob, dwp, M=C, N=C, K=B, a_mn=True, b_mn=True, lda=C, ldb=C, ldo=C)
            ops.colsum_bf16(dyb, dbp, rows=B, C_=C, groups=1)
            dq, dkv = e(1, B, C), torch.empty(1, T, 2 * C, dtype=BF16, device=dev)
            ops.xattn_bwd(q, kv, probs, do, dq, dkv, K=1, B=B, N=N, H=heads, scale=64 ** -0.5)
            dqb = _bf16(dq.view(B, C))
            dx = e(T, C)
            ops.gemm(dkv, wkvb, dx, M=T, N=C, K=2 * C, b_mn=True, lda=2 * C, ldb=C, ldo=C)
            dxq, dwq, dbq, dwkv, dbkv = e(B, C), e(C, C), e(C), e(2 * C, C), e(2 * C)
            ops.gemm(dqb, wqb, dxq, M=B, N=C, K=C, b_mn=True, lda=C, ldb=C, ldo=C)
            ops.gather_rows_f32(dxq, dx, rows=B, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=N * C, dst_gs=0,
                                accumulate=True)                                                  # the query came from token 0
            ops.gemm(dqb, xb, dwq, M=C, N=C, K=B, a_mn=True, b_mn=True, lda=C, ldb=N * C, ldo=C)
            ops.colsum_bf16(dqb, dbq, rows=B, C_=C, groups=1)
            ops.gemm(dkv, xb, dwkv, M=2 * C, N=C, K=T, a_mn=True, b_mn=True, lda=2 * C, ldb=C, ldo=C)
            ops.colsum_bf16(dkv, dbkv, rows=T, C_=2 * C, groups=1)
        return (dx.view(B, N, C), dwq, dbq, dwkv[:C], dbkv[:C], dwkv[C:], dbkv[C:], dwp, dbp, None)


def cross_attention(x, wq, bq, wk, bk, wv, bv, wp, bp, heads):
    _check(x, wq.shape[1], "cross_attention")
    if x.dim() != 3 or x.shape[-1] != 64 * heads:
        raise _abi.CavitError("cross_attention: x must be [B, N, C] with C == 64 * heads")
    return _CrossAttention.apply(x, wq, bq, wk, bk, wv, bv, wp, bp, heads)
