"""The fp32-tolerance mode of the engine (`precision="fp32"`): ModelCross / ModelVIT forward and backward within ~1e-3 of
the reference's fp32 path (north_star; the reference trains in fp32: `L.Trainer` without `precision=`,
/root/reference/main_mist.py:211-218).

Same flat buffers, same static launch sequence idea, same CUDA-graph capture as the bf16 mode (cavit/engine.py) — what
changes is the number format between the kernels:

  * every GEMM operand is a bf16 hi + lo PAIR (x ~ hi + lo, 16 mantissa bits) and `cavit_gemm` multiplies pairs with three
    tcgen05 MMAs per product (Ah Bh + Al Bh + Ah Bl, one fp32 accumulation in TMEM): the projections stay on the tensor
    cores at ~2^-16 relative operand error instead of 2^-9;
  * GEMM outputs that feed anything but a residual add are fp32 (the fp32 instantiations of the EPI_NONE / EPI_BIAS
    epilogues); the consumer (LayerNorm, GELU, the attention kernels, the cast) produces the next operand pair;
  * self-attention runs in fp32 on the CUDA cores (csrc/attn_f32.cu: fused online softmax, no N x N matrix in memory);
  * GELU is the exact erf form (libdevice erff), LayerNorm / softmax statistics / residual stream / loss were fp32 already;
  * cross-attention takes the FOLDED route always (csrc/xfold.cu is fp32 end to end; its [B, .] GEMMs use operand pairs).

Cost: 3x the tensor work of the projections, fp32 activations between kernels (2x the bytes) and CUDA-core attention —
this is the accuracy mode, `bench.py --precision fp32` reports its throughput next to the bf16 headline.
Limits: ModelCross / ModelVIT only, dropout = 0 (the parity configuration), heads in {1, 2, 3, 4, 6, 8, 12, 16}.

Reference semantics followed: /root/reference/model_cross.py:186-212, 128-148, 111-114, 69-72; modelv3.py:123-147.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import _abi, ops
from ._abi import EPI_BIAS, EPI_BIAS_RESID, EPI_EMBED

BF16, F32 = torch.bfloat16, torch.float32


def check_supported(eng):
    if eng.kind not in ("cross", "vit"):
        raise _abi.CavitError("fp32-tolerance mode: ModelCross / ModelVIT only")
    if eng.K and not eng.fold_ok:
        raise _abi.CavitError("fp32-tolerance mode: the folded cross-attention needs heads in {1,2,3,4,6,8,12,16}, "
                              "hidden_dim <= 1024 and at most 16 fusions")


def plan(eng, B: int, train: bool) -> Dict:
    """Activation / scratch buffers of one (batch, mode). `sp` = bf16 hi plane carrying its lo plane (ops.split_pair)."""
    check_supported(eng)
    dev = eng.device
    G, N, C, F, H, K = eng.G, eng.N, eng.C, eng.F, eng.H, eng.K
    T = B * N
    HC = H * C
    R = eng.Mimg * B * eng.Np

    def e(shape, dt=F32):
        return torch.empty(shape, dtype=dt, device=dev)

    def sp(shape):
        return ops.split_pair(shape, dev)

    a: Dict = {}
    eng.fold = bool(K)
    a["patches"] = sp((R, eng.P))
    nL = eng.L if train else 1
    a["X"] = [e((G, T, C)) for _ in range(2 * eng.L + 1 if train else 3)]
    for nm, mk in [("xn1", lambda: sp((G, T, C))), ("mean1", lambda: e((G, T))), ("rstd1", lambda: e((G, T))),
                   ("qkv32", lambda: e((G, T, 3 * C))), ("ao32", lambda: e((G, T, C))), ("ao", lambda: sp((G, T, C))),
                   ("lse", lambda: e((G, B, H, N))), ("xn2", lambda: sp((G, T, C))), ("mean2", lambda: e((G, T))),
                   ("rstd2", lambda: e((G, T))), ("u32", lambda: e((G, T, F))), ("h", lambda: sp((G, T, F)))]:
        a[nm] = [mk() for _ in range(nL)]
    if K:
        nF = eng.cfg.num_multi_blocks if train else 1
        for nm, mk in [("f_cls", lambda: e((K, B, C))), ("f_xncls", lambda: sp((K, B, C))), ("f_mean0", lambda: e((K, B))),
                       ("f_rstd0", lambda: e((K, B))), ("f_q32", lambda: e((K, B, C))), ("f_qb", lambda: sp((K, B, C))),
                       ("f_qp", lambda: e((K, B, HC))), ("f_zhat", lambda: e((K, B, HC))), ("f_zb", lambda: sp((K, B, HC))),
                       ("f_Ekv", lambda: sp((K, 2, C, HC))), ("f_probs", lambda: e((K, B, H, N))),
                       ("f_mean", lambda: e((K, T))), ("f_rstd", lambda: e((K, T))), ("f_xo", lambda: e((K, B, C))),
                       ("f_xob", lambda: sp((K, B, C))), ("f_y", lambda: e((K, B, C))), ("f_yn", lambda: sp((K, B, C))),
                       ("f_meany", lambda: e((K, B))), ("f_rstdy", lambda: e((K, B))), ("f_u32", lambda: e((K, B, F))),
                       ("f_h", lambda: sp((K, B, F))), ("f_z", lambda: e((K, B, C)))]:
            a[nm] = [mk() for _ in range(nF)]
        a["xf_scratch"] = ops.xfold_scratch(K, B, N, H, dev)
    a["clsn"] = sp((G, B, C))
    a["meanc"], a["rstdc"] = e((G, B)), e((G, B))
    a["uh32"], a["hh32"] = e((G, B, F)), e((G, B, F))
    a["logits"], a["loss"] = e((B, eng.classes)), e((1,))
    if train:
        a["dX"], a["dXb"] = e((G, T, C)), sp((G, T, C))
        a["dbig32"], a["dbig"] = e((G, T, F)), sp((G, T, F))
        a["dmid32"] = e((G, T, C))
        a["dqkv32"], a["dqkv"] = e((G, T, 3 * C)), sp((G, T, 3 * C))
        a["delta"] = e((G, B, H, N))
        a["ln_ws"] = ops.ln_bwd_workspace(max(G, K, 1), C, dev)
        a["dhh32"], a["duh"], a["dclsn32"] = e((G, B, F)), sp((G, B, F)), e((G, B, C))
        a["dcomp"] = sp((R, C))
        if K:
            a["d_z"], a["d_zb"] = e((K, B, C)), sp((K, B, C))
            a["d_u32"], a["d_u"] = e((K, B, F)), sp((K, B, F))
            a["d_yn32"], a["d_y"], a["d_yb"] = e((K, B, C)), e((K, B, C)), sp((K, B, C))
            a["d_xo32"], a["d_xob"] = e((K, B, C)), sp((K, B, C))
            a["d_gz"], a["d_qp"], a["d_qpb"] = e((K, B, HC)), e((K, B, HC)), sp((K, B, HC))
            a["d_Ekv"] = e((K, 2, C, HC))
            a["d_lnA"], a["d_bv"] = e((2, K, C)), e((K, C))
            a["d_q32"], a["d_qb"] = e((K, B, C)), sp((K, B, C))
            a["d_xncls32"], a["d_clsq"] = e((K, B, C)), e((K, B, C))
    return a


# ---------------------------------------------------------------------------------------------------- forward
def forward_impl(eng, img, labels, train: bool):
    cfg = eng.cfg
    a, G, N, C, F, H, T, K, B = eng.a, eng.G, eng.N, eng.C, eng.F, eng.H, eng.T, eng.K, eng.B
    w, wb = eng.w, eng.wb
    X0 = a["X"][0]
    ops.patchify(img, a["patches"], patch_size=cfg.patch_size, sample_major=(eng.kind == "vit"))
    ops.gemm(a["patches"], wb("embed.w"), X0, M=eng.Mimg * B * eng.Np, N=C, K=eng.P, lda=eng.P, ldb=eng.P, ldo=C,
             epi=EPI_EMBED, bias=w("embed.b"), resid=w("pos"), ldr=C, embed_np=N - 1)
    ops.cls_rows(w("cls"), w("pos"), X0, M=G, B=B, N=N, C_=C)
    xi = 0
    for l in range(eng.L):
        s = l if train else 0
        if train:
            x_in, x_mid, x_out = a["X"][2 * l], a["X"][2 * l + 1], a["X"][2 * l + 2]
        else:
            x_in, x_mid, x_out = a["X"][xi], a["X"][(xi + 1) % 3], a["X"][(xi + 2) % 3]
            xi = (xi + 2) % 3
        tag = f"L{l}"
        ops.ln_fwd_split(x_in, w(f"{tag}.ln1.w"), w(f"{tag}.ln1.b"), a["xn1"][s], a["mean1"][s], a["rstd1"][s],
                         rows_per_group=T, groups=G, C=C, eps=eng.eps)
        eng._fwd(a["xn1"][s], wb(f"{tag}.wqkv"), a["qkv32"][s], G=G, T=T, N=3 * C, K=C)
        ops.attn_fwd_f32(a["qkv32"][s], a["ao32"][s], a["lse"][s], G=G, B=B, N=N, H=H, scale=eng.scale)
        if H != 1:
            ops.cast_split(a["ao32"][s], a["ao"][s])
            eng._fwd(a["ao"][s], wb(f"{tag}.wo"), x_mid, G=G, T=T, N=C, K=C, epi=EPI_BIAS_RESID, bias=w(f"{tag}.bo"),
                     resid=x_in)
        else:   # to_out = nn.Identity() when heads == 1 (model_cross.py:37,44-48): x_mid = x_in + attention output
            _add_rows(x_in, a["ao32"][s], x_mid, G * T, C)
        ops.ln_fwd_split(x_mid, w(f"{tag}.ln2.w"), w(f"{tag}.ln2.b"), a["xn2"][s], a["mean2"][s], a["rstd2"][s],
                         rows_per_group=T, groups=G, C=C, eps=eng.eps)
        eng._fwd(a["xn2"][s], wb(f"{tag}.w1"), a["u32"][s], G=G, T=T, N=F, K=C, epi=EPI_BIAS, bias=w(f"{tag}.b1"))
        ops.gelu_split(a["u32"][s], h=a["h"][s])
        eng._fwd(a["h"][s], wb(f"{tag}.w2"), x_out, G=G, T=T, N=C, K=F, epi=EPI_BIAS_RESID, bias=w(f"{tag}.b2"), resid=x_mid)
        if eng.kind == "cross" and K and (l + 1) % cfg.num_self_blocks == 0:
            _fusion_fwd(eng, l // cfg.num_self_blocks, x_out, train)
    x_fin = a["X"][2 * eng.L] if train else a["X"][xi]
    eng._x_fin = x_fin
    ops.ln_fwd_split(x_fin, w("fin.ln.w"), w("fin.ln.b"), a["clsn"], a["meanc"], a["rstdc"], rows_per_group=B, groups=G, C=C,
                     x_row_stride=N * C, x_gs=T * C)
    eng._fwd(a["clsn"], wb("head.w1"), a["uh32"], G=G, T=B, N=F, K=C, epi=EPI_BIAS, bias=w("head.b1"))
    ops.gelu_split(a["uh32"], h32=a["hh32"])
    ops.head_loss_fwd_f32(a["hh32"], w("head.w2"), w("head.b2"), labels, a["logits"], a["loss"], M=G, B=B, F=F,
                          classes=eng.classes, smoothing=eng.smoothing)
    eng._labels = labels
    eng.saved_valid = train
    return a["logits"], a["loss"]


def _add_rows(x, y, out, rows, C):
    """out = x + y (fp32 rows) with the row-copy kernel: copy, then accumulate."""
    kw = dict(rows=rows, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=C, dst_gs=0)
    ops.gather_rows_f32(x, out, **kw)
    ops.gather_rows_f32(y, out, accumulate=True, **kw)


def _fusion_fwd(eng, mb: int, X, train: bool):
    """CrossAttentionBlock over the K fusions of multi-block `mb` (folded single-query attention, model_cross.py:88-114,
    140-142); rewrites the CLS rows of the receiving streams in place."""
    a, N, C, F, H, T, K, B = eng.a, eng.N, eng.C, eng.F, eng.H, eng.T, eng.K, eng.B
    w, wb = eng.w, eng.wb
    s = mb if train else 0
    tag = f"X{mb}"
    HC = H * C
    f_cls = a["f_cls"][s]
    ops.gather_rows_f32_indexed(X, f_cls, rows=B, C_=C, src_row_stride=N * C, src_gs=T * C, src_groups=eng.cls_src,
                                dst_row_stride=C, dst_gs=B * C)
    ops.ln_fwd_split(f_cls, w(f"{tag}.lnA.w"), w(f"{tag}.lnA.b"), a["f_xncls"][s], a["f_mean0"][s], a["f_rstd0"][s],
                     rows_per_group=B, groups=K, C=C)
    eng._fwd(a["f_xncls"][s], wb(f"{tag}.wq"), a["f_q32"][s], G=K, T=B, N=C, K=C, epi=EPI_BIAS, bias=w(f"{tag}.bq"))
    ops.cast_split(a["f_q32"][s], a["f_qb"][s])
    Ekv = a["f_Ekv"][s]
    ops.expand_heads(wb(f"{tag}.wkv"), Ekv, groups=2 * K, C_=C, H=H)
    ops.gemm(a["f_qb"][s], Ekv, a["f_qp"][s], M=B, N=HC, K=C, groups=K, b_mn=True, lda=C, ldb=HC, ldo=HC, a_gs=B * C,
             b_gs=2 * C * HC, out_gs=B * HC)
    ops.xfold_fwd(X, f_cls, a["f_qp"][s], w(f"{tag}.lnA.w"), w(f"{tag}.lnA.b"), a["f_zhat"][s], a["f_zb"][s], a["f_probs"][s],
                  a["f_mean"][s], a["f_rstd"][s], a["xf_scratch"], K=K, B=B, N=N, C_=C, H=H, cls_src=eng.cls_src,
                  tok_src=eng.tok_src, scale=eng.scale)
    ops.gemm(a["f_zb"][s], ops.sub(Ekv, (slice(None), 1)), a["f_xo"][s], M=B, N=C, K=HC, groups=K, lda=HC, ldb=HC, ldo=C,
             a_gs=B * HC, b_gs=2 * C * HC, out_gs=B * C, epi=EPI_BIAS, bias=w(f"{tag}.bkv")[:, C:], bias_gs=2 * C)
    ops.cast_split(a["f_xo"][s], a["f_xob"][s])
    eng._fwd(a["f_xob"][s], wb(f"{tag}.wp"), a["f_y"][s], G=K, T=B, N=C, K=C, epi=EPI_BIAS_RESID, bias=w(f"{tag}.bp"),
             resid=f_cls)
    ops.ln_fwd_split(a["f_y"][s], w(f"{tag}.lnF.w"), w(f"{tag}.lnF.b"), a["f_yn"][s], a["f_meany"][s], a["f_rstdy"][s],
                     rows_per_group=B, groups=K, C=C)
    eng._fwd(a["f_yn"][s], wb(f"{tag}.w1"), a["f_u32"][s], G=K, T=B, N=F, K=C, epi=EPI_BIAS, bias=w(f"{tag}.b1"))
    ops.gelu_split(a["f_u32"][s], h=a["f_h"][s])
    eng._fwd(a["f_h"][s], wb(f"{tag}.w2"), a["f_z"][s], G=K, T=B, N=C, K=F, epi=EPI_BIAS_RESID, bias=w(f"{tag}.b2"),
             resid=a["f_y"][s])
    ops.gather_rows_f32_indexed(a["f_z"][s], X, rows=B, C_=C, src_row_stride=C, src_gs=B * C, dst_row_stride=N * C,
                                dst_gs=T * C, dst_groups=eng.cls_src)


# ---------------------------------------------------------------------------------------------------- backward
def backward_impl(eng, loss_scale: float, done, loss_scale_dev):
    cfg = eng.cfg
    a, G, N, C, F, H, T, K, B = eng.a, eng.G, eng.N, eng.C, eng.F, eng.H, eng.T, eng.K, eng.B
    w, wb, g = eng.w, eng.wb, eng.g
    ws = a["ln_ws"]
    # ---- loss, heads, final norm
    ops.head_loss_bwd_f32(a["hh32"], w("head.w2"), eng._labels, a["logits"], a["dhh32"], g("head.w2"), g("head.b2"), M=G, B=B,
                          F=F, classes=eng.classes, smoothing=eng.smoothing, loss_scale=loss_scale, loss_scale_dev=loss_scale_dev)
    ops.gelu_bwd_split(a["dhh32"], a["uh32"], a["duh"])
    eng._dgrad(a["duh"], wb("head.w1"), a["dclsn32"], G=G, T=B, N=F, K=C)
    eng._wgrad(a["duh"], a["clsn"], g("head.w1"), G=G, T=B, N=F, K=C)
    eng._colsum(a["duh"], g("head.b1"), G=G, T=B, N=F)
    dX, dXb = a["dX"], a["dXb"]
    dX.zero_()
    ops.ln_bwd_split(a["dclsn32"], eng._x_fin, a["meanc"], a["rstdc"], w("fin.ln.w"), dX, g("fin.ln.w"), g("fin.ln.b"), ws,
                     rows_per_group=B, groups=G, C=C, x_row_stride=N * C, x_gs=T * C, dx_row_stride=N * C, dx_gs=T * C)
    done("head")
    need_cast = True     # dXb must mirror dX before the first layer's GEMMs
    b2_done = False      # fc2 bias gradient of this layer already produced by the LayerNorm backward above it
    for l in reversed(range(eng.L)):
        tag = f"L{l}"
        if eng.kind == "cross" and K and (l + 1) % cfg.num_self_blocks == 0:
            _fusion_bwd(eng, l // cfg.num_self_blocks, dX)
            done(f"X{l // cfg.num_self_blocks}")
            need_cast = True
        if need_cast:
            ops.cast_split(dX, dXb)
        need_cast = False
        x_in, x_mid = a["X"][2 * l], a["X"][2 * l + 1]
        # FFN: x_out = x_mid + W2 gelu(W1 LN2(x_mid) + b1) + b2
        eng._dgrad(dXb, wb(f"{tag}.w2"), a["dbig32"], G=G, T=T, N=C, K=F)
        ops.gelu_bwd_split(a["dbig32"], a["u32"][l], a["dbig"])
        eng._wgrad(dXb, a["h"][l], g(f"{tag}.w2"), G=G, T=T, N=C, K=F)
        if not b2_done:
            eng._colsum(dXb, g(f"{tag}.b2"), G=G, T=T, N=C)
        eng._dgrad(a["dbig"], wb(f"{tag}.w1"), a["dmid32"], G=G, T=T, N=F, K=C)
        eng._wgrad(a["dbig"], a["xn2"][l], g(f"{tag}.w1"), G=G, T=T, N=F, K=C)
        eng._colsum(a["dbig"], g(f"{tag}.b1"), G=G, T=T, N=F)
        bo_fused = H != 1    # the column sums of this LayerNorm backward's output are the out-proj bias gradient
        ops.ln_bwd_split(a["dmid32"], x_mid, a["mean2"][l], a["rstd2"][l], w(f"{tag}.ln2.w"), dX, g(f"{tag}.ln2.w"),
                         g(f"{tag}.ln2.b"), ws, rows_per_group=T, groups=G, C=C, dresid=dX, dx_split=dXb,
                         dcol=g(f"{tag}.bo") if bo_fused else None)
        # attention: x_mid = x_in + Wo attn(LN1(x_in)) + bo
        if H != 1:
            eng._dgrad(dXb, wb(f"{tag}.wo"), a["dmid32"], G=G, T=T, N=C, K=C)
            eng._wgrad(dXb, a["ao"][l], g(f"{tag}.wo"), G=G, T=T, N=C, K=C)
            dao = a["dmid32"]
        else:
            dao = dX
        ops.attn_bwd_f32(a["qkv32"][l], a["ao32"][l], dao, a["lse"][l], a["dqkv32"], a["delta"], G=G, B=B, N=N, H=H,
                         scale=eng.scale)
        ops.cast_split(a["dqkv32"], a["dqkv"])
        eng._dgrad(a["dqkv"], wb(f"{tag}.wqkv"), a["dmid32"], G=G, T=T, N=3 * C, K=C)
        eng._wgrad(a["dqkv"], a["xn1"][l], g(f"{tag}.wqkv"), G=G, T=T, N=3 * C, K=C)
        fusion_next = eng.kind == "cross" and K and l % cfg.num_self_blocks == 0
        b2_done = l >= 1 and not fusion_next
        ops.ln_bwd_split(a["dmid32"], x_in, a["mean1"][l], a["rstd1"][l], w(f"{tag}.ln1.w"), dX, g(f"{tag}.ln1.w"),
                         g(f"{tag}.ln1.b"), ws, rows_per_group=T, groups=G, C=C, dresid=dX, dx_split=dXb,
                         dcol=g(f"L{l - 1}.b2") if b2_done else None)
        done(tag)
    # ---- embedding: d(pos), d(cls), dW = dY^T unfold(x), db
    ops.embed_param_grads(dX, g("pos"), g("cls"), M=G, B=B, N=N, C_=C)
    ops.compact_patch_rows_bf16(dXb, a["dcomp"], S=G * B, Np=N - 1, C_=C)
    R = eng.Mimg * B * eng.Np
    eng._wgrad(a["dcomp"], a["patches"], g("embed.w"), G=1, T=R, N=C, K=eng.P)
    eng._colsum(a["dcomp"], g("embed.b"), G=1, T=R, N=C)
    done("embed")
    return eng.grad


def _fusion_bwd(eng, mb: int, dX):
    a, N, C, F, H, T, K, B = eng.a, eng.N, eng.C, eng.F, eng.H, eng.T, eng.K, eng.B
    w, wb, g = eng.w, eng.wb, eng.g
    tag, s, ws, HC = f"X{mb}", mb, a["ln_ws"], H * C
    d_z = a["d_z"]
    # gradient of the new CLS rows; the old CLS rows of receiving streams get no pass-through gradient
    ops.gather_rows_f32_indexed(dX, d_z, rows=B, C_=C, src_row_stride=N * C, src_gs=T * C, src_groups=eng.cls_src,
                                dst_row_stride=C, dst_gs=B * C, zero_src=True)
    ops.cast_split(d_z, a["d_zb"])
    # FFN on the single CLS token
    eng._dgrad(a["d_zb"], wb(f"{tag}.w2"), a["d_u32"], G=K, T=B, N=C, K=F)
    ops.gelu_bwd_split(a["d_u32"], a["f_u32"][s], a["d_u"])
    eng._wgrad(a["d_zb"], a["f_h"][s], g(f"{tag}.w2"), G=K, T=B, N=C, K=F)
    eng._colsum(a["d_zb"], g(f"{tag}.b2"), G=K, T=B, N=C)
    eng._dgrad(a["d_u"], wb(f"{tag}.w1"), a["d_yn32"], G=K, T=B, N=F, K=C)
    eng._wgrad(a["d_u"], a["f_yn"][s], g(f"{tag}.w1"), G=K, T=B, N=F, K=C)
    eng._colsum(a["d_u"], g(f"{tag}.b1"), G=K, T=B, N=F)
    ops.ln_bwd_split(a["d_yn32"], a["f_y"][s], a["f_meany"][s], a["f_rstdy"][s], w(f"{tag}.lnF.w"), a["d_y"],
                     g(f"{tag}.lnF.w"), g(f"{tag}.lnF.b"), ws, rows_per_group=B, groups=K, C=C, dresid=d_z, dx_split=a["d_yb"])
    # y = proj(xattn) + cls_in; everything left of the token streams is a [B, .] problem (see csrc/xfold.cu)
    Ekv, dE = a["f_Ekv"][s], a["d_Ekv"]
    Ev = ops.sub(Ekv, (slice(None), 1))
    gkv, gbkv = g(f"{tag}.wkv"), g(f"{tag}.bkv").view(K, 2, C)
    eng._dgrad(a["d_yb"], wb(f"{tag}.wp"), a["d_xo32"], G=K, T=B, N=C, K=C)
    ops.cast_split(a["d_xo32"], a["d_xob"])
    eng._wgrad(a["d_yb"], a["f_xob"][s], g(f"{tag}.wp"), G=K, T=B, N=C, K=C)
    eng._colsum(a["d_yb"], g(f"{tag}.bp"), G=K, T=B, N=C)
    # value side: gz_h = Wv_h^T do_h, dWv (expanded), dbv = sum_b do, dbk = 0 (scores are shift invariant)
    ops.gemm(a["d_xob"], Ev, a["d_gz"], M=B, N=HC, K=C, groups=K, b_mn=True, lda=C, ldb=HC, ldo=HC, a_gs=B * C,
             b_gs=2 * C * HC, out_gs=B * HC)
    ops.gemm(a["d_xob"], a["f_zb"][s], dE[:, 1], M=C, N=HC, K=B, groups=K, a_mn=True, b_mn=True, lda=C, ldb=HC, ldo=HC,
             a_gs=B * C, b_gs=B * HC, out_gs=2 * C * HC)
    eng._colsum(a["d_xob"], a["d_bv"], G=K, T=B, N=C)
    ops.gather_rows_f32(a["d_bv"], gbkv[:, 1], rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=2 * C, dst_gs=0)
    a["d_lnA"].zero_()
    ops.gather_rows_f32(a["d_lnA"][0], gbkv[:, 0], rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=2 * C,
                        dst_gs=0)     # dbk = 0
    ops.xfold_bwd(eng._x_for_fusion(mb), a["f_cls"][s], a["f_qp"][s], w(f"{tag}.lnA.w"), a["f_zhat"][s], a["f_probs"][s],
                  a["f_mean"][s], a["f_rstd"][s], a["d_gz"], a["xf_scratch"], dX, a["d_qp"], a["d_lnA"][0], a["d_lnA"][1], K=K,
                  B=B, N=N, C_=C, H=H, cls_src=eng.cls_src, tok_src=eng.tok_src, scale=eng.scale, exact_fp32=True)
    # key side: dq_h = Wk_h dq'_h, dWk (expanded) = q^T dq'
    ops.cast_split(a["d_qp"], a["d_qpb"])
    ops.gemm(a["d_qpb"], Ekv, a["d_q32"], M=B, N=C, K=HC, groups=K, lda=HC, ldb=HC, ldo=C, a_gs=B * HC, b_gs=2 * C * HC,
             out_gs=B * C)
    ops.cast_split(a["d_q32"], a["d_qb"])
    ops.gemm(a["f_qb"][s], a["d_qpb"], dE, M=C, N=HC, K=B, groups=K, a_mn=True, b_mn=True, lda=C, ldb=HC, ldo=HC, a_gs=B * C,
             b_gs=B * HC, out_gs=2 * C * HC)
    ops.fold_heads(dE, gkv, groups=2 * K, C_=C, H=H)
    # query path through the LayerNorm of the CLS rows
    eng._dgrad(a["d_qb"], wb(f"{tag}.wq"), a["d_xncls32"], G=K, T=B, N=C, K=C)
    eng._wgrad(a["d_qb"], a["f_xncls"][s], g(f"{tag}.wq"), G=K, T=B, N=C, K=C)
    eng._colsum(a["d_qb"], g(f"{tag}.bq"), G=K, T=B, N=C)
    ops.ln_bwd_split(a["d_xncls32"], a["f_cls"][s], a["f_mean0"][s], a["f_rstd0"][s], w(f"{tag}.lnA.w"), a["d_clsq"],
                     g(f"{tag}.lnA.w"), g(f"{tag}.lnA.b"), ws, rows_per_group=B, groups=K, C=C)
    ops.gather_rows_f32(a["d_lnA"][0], g(f"{tag}.lnA.w"), rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=C,
                        dst_gs=0, accumulate=True)
    ops.gather_rows_f32(a["d_lnA"][1], g(f"{tag}.lnA.b"), rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=C,
                        dst_gs=0, accumulate=True)
    ops.gather_rows_f32_indexed(a["d_clsq"], dX, rows=B, C_=C, src_row_stride=C, src_gs=B * C, dst_row_stride=N * C,
                                dst_gs=T * C, dst_groups=eng.cls_src, accumulate=True)
    # residual path of the CLS token: y = ... + cls_in
    ops.gather_rows_f32_indexed(a["d_y"], dX, rows=B, C_=C, src_row_stride=C, src_gs=B * C, dst_row_stride=N * C, dst_gs=T * C,
                                dst_groups=eng.cls_src, accumulate=True)
