"""Kernel micro-benchmarks on one B200 (CUDA events, L2-cold by rotating over > 126 MB of buffers
where the working set is small). Prints TFLOP/s or GB/s per kernel for the shapes of a workload."""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))

import torch  # noqa: E402

from cavit import _abi, ops  # noqa: E402
from cavit._abi import EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID, EPI_GELU_BWD, EPI_NONE  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=384)
    ap.add_argument("--H", type=int, default=6)
    ap.add_argument("--F", type=int, default=1536)
    ap.add_argument("--N", type=int, default=197)
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--G", type=int, default=4)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    _abi.require_device(0)
    C, H, F, N, B, G = a.C, a.H, a.F, a.N, a.B, a.G
    T = B * N
    rnd = lambda *s: torch.randn(*s, device=DEV)  # noqa: E731
    x = rnd(G, T, C).to(BF)
    xf = rnd(G, T, F).to(BF)
    res = rnd(G, T, C)

    def gemm_line(name, fn, flops):
        ms = timeit(fn)
        print(f"{name:34s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)

    def mem_line(name, fn, bytes_):
        ms = timeit(fn)
        print(f"{name:34s} {ms:8.3f} ms  {bytes_ / ms / 1e6:8.1f} GB/s", flush=True)

    if not a.only or "gemm" in a.only:
        wqkv = (rnd(G, 3 * C, C) / math.sqrt(C)).to(BF)
        wo = (rnd(G, C, C) / math.sqrt(C)).to(BF)
        w1 = (rnd(G, F, C) / math.sqrt(C)).to(BF)
        w2 = (rnd(G, C, F) / math.sqrt(F)).to(BF)
        bC, bF_ = rnd(G, C), rnd(G, F)
        qkv = torch.empty(G, T, 3 * C, device=DEV, dtype=BF)
        y32 = torch.empty(G, T, C, device=DEV)
        h = torch.empty(G, T, F, device=DEV, dtype=BF)
        u = torch.empty(G, T, F, device=DEV, dtype=BF)
        yb = torch.empty(G, T, C, device=DEV, dtype=BF)
        gemm_line("fwd qkv   [T,C]x[3C,C] none", lambda: ops.linear_fwd(x, wqkv, qkv), 2 * G * T * 3 * C * C)
        gemm_line("fwd out   [T,C]x[C,C] bias+resid", lambda: ops.linear_fwd(x, wo, y32, epi=EPI_BIAS_RESID, bias=bC, resid=res), 2 * G * T * C * C)
        gemm_line("fwd fc1   [T,C]x[F,C] bias+gelu", lambda: ops.linear_fwd(x, w1, h, epi=EPI_BIAS_GELU, bias=bF_, aux=u), 2 * G * T * F * C)
        gemm_line("fwd fc2   [T,F]x[C,F] bias+resid", lambda: ops.linear_fwd(xf, w2, y32, epi=EPI_BIAS_RESID, bias=bC, resid=res), 2 * G * T * F * C)
        gemm_line("dgrad fc2 [T,C]x[C,F] gelu'", lambda: ops.linear_dgrad(x, w2, h, epi=EPI_GELU_BWD, aux=u), 2 * G * T * F * C)
        gemm_line("dgrad fc1 [T,F]x[F,C]", lambda: ops.linear_dgrad(xf, w1, yb), 2 * G * T * F * C)
        gemm_line("dgrad qkv [T,3C]x[3C,C]", lambda: ops.linear_dgrad(qkv, wqkv, yb), 2 * G * T * 3 * C * C)
        gemm_line("dgrad out [T,C]x[C,C]", lambda: ops.linear_dgrad(x, wo, yb), 2 * G * T * C * C)
        dw1 = torch.empty(G, F, C, device=DEV)
        dw2 = torch.empty(G, C, F, device=DEV)
        dwq = torch.empty(G, 3 * C, C, device=DEV)
        dwo = torch.empty(G, C, C, device=DEV)
        sweep = [int(v) for v in os.environ.get("CAVIT_KB_SPLITS", "").split(",") if v]
        for sk in sweep or (1, 2, 4, 8):
            gemm_line(f"wgrad fc1 dW[F,C] split_k={sk}", lambda: ops.linear_wgrad(xf, x, dw1, split_k=sk), 2 * G * T * F * C)
        for sk in sweep or (1, 2, 4):
            gemm_line(f"wgrad fc2 dW[C,F] split_k={sk}", lambda: ops.linear_wgrad(x, xf, dw2, split_k=sk), 2 * G * T * F * C)
        for sk in sweep or (1, 2, 4):
            gemm_line(f"wgrad qkv dW[3C,C] split_k={sk}", lambda: ops.linear_wgrad(qkv, x, dwq, split_k=sk), 2 * G * T * 3 * C * C)
        for sk in sweep or (1, 4, 8, 16):
            gemm_line(f"wgrad out dW[C,C] split_k={sk}", lambda: ops.linear_wgrad(x, x, dwo, split_k=sk), 2 * G * T * C * C)
    if not a.only or "attn" in a.only:
        qkv = rnd(G, T, 3 * C).to(BF)
        o = torch.empty(G, T, C, device=DEV, dtype=BF)
        lse = torch.empty(G, B, H, N, device=DEV)
        gemm_line("attn fwd", lambda: ops.attn_fwd(qkv, o, lse, G=G, B=B, N=N, H=H, scale=0.125), 4 * G * B * H * N * N * 64)
        do = rnd(G, T, C).to(BF)
        dqkv = torch.empty(G, T, 3 * C, device=DEV, dtype=BF)
        delta = torch.empty(G, B, H, N, device=DEV)
        acc = torch.empty(G, T, C, device=DEV)
        gemm_line("attn bwd", lambda: ops.attn_bwd(qkv, o, do, lse, dqkv, delta, acc, G=G, B=B, N=N, H=H, scale=0.125),
                  10 * G * B * H * N * N * 64)
    if not a.only or "ln" in a.only:
        xs = rnd(G, T, C)
        gamma, beta = rnd(G, C), rnd(G, C)
        y = torch.empty(G, T, C, device=DEV, dtype=BF)
        mean, rstd = torch.empty(G, T, device=DEV), torch.empty(G, T, device=DEV)
        mem_line("ln fwd", lambda: ops.ln_fwd(xs, gamma, beta, y, mean, rstd, rows_per_group=T, groups=G, C=C), G * T * C * 6)
        dy = rnd(G, T, C).to(BF)
        dx = rnd(G, T, C)
        dxb = torch.empty(G, T, C, device=DEV, dtype=BF)
        dg, db = torch.empty(G, C, device=DEV), torch.empty(G, C, device=DEV)
        ws = ops.ln_bwd_workspace(G, C, DEV)
        mem_line("ln bwd (+resid, +bf16 copy)", lambda: ops.ln_bwd(dy, xs, mean, rstd, gamma, dx, dg, db, ws, rows_per_group=T, groups=G, C=C,
                                                                  dresid=dx, dx_bf16=dxb), G * T * C * 16)
        out = torch.empty(G, C, device=DEV)
        mem_line("colsum bf16 [T,C]", lambda: ops.colsum_bf16(dy, out, rows=T, C_=C, groups=G), G * T * C * 2)
        big = rnd(G, T, F).to(BF)
        outF = torch.empty(G, F, device=DEV)
        mem_line("colsum bf16 [T,F]", lambda: ops.colsum_bf16(big, outF, rows=T, C_=F, groups=G), G * T * F * 2)
        src = rnd(G * T * C)
        dst = torch.empty(G * T * C, device=DEV, dtype=BF)
        mem_line("cast fp32->bf16", lambda: ops.cast_bf16(src, dst), G * T * C * 6)
    if not a.only or "xfold" in a.only:
        Kf = 4
        f32 = dict(device=DEV, dtype=torch.float32)
        X = torch.randn(G, B * N, C, **f32)
        cls = torch.randn(Kf, B, C, **f32)
        qp = torch.randn(Kf, B, H * C, **f32) * 0.05
        lnw, lnb = torch.ones(Kf, C, **f32), torch.zeros(Kf, C, **f32)
        zhat = torch.empty(Kf, B, H * C, **f32)
        z = torch.empty(Kf, B, H * C, device=DEV, dtype=BF)
        probs = torch.empty(Kf, B, H, N, **f32)
        mean, rstd = torch.empty(Kf, B, N, **f32), torch.empty(Kf, B, N, **f32)
        scratch = ops.xfold_scratch(Kf, B, N, H, DEV)
        kw = dict(K=Kf, B=B, N=N, C_=C, H=H, cls_src=[0, 1, 2, 3], tok_src=[1, 2, 3, 0], scale=0.125)
        mem_line("xfold fwd", lambda: ops.xfold_fwd(X, cls, qp, lnw, lnb, zhat, z, probs, mean, rstd, scratch, **kw), Kf * B * N * C * 4)
        gz = torch.randn(Kf, B, H * C, **f32) * 0.05
        dX = torch.zeros(G, B * N, C, **f32)
        dqp = torch.empty(Kf, B, H * C, **f32)
        dg, db = torch.zeros(Kf, C, **f32), torch.zeros(Kf, C, **f32)
        mem_line("xfold bwd", lambda: ops.xfold_bwd(X, cls, qp, lnw, zhat, probs, mean, rstd, gz, scratch, dX, dqp, dg, db, **kw),
                 Kf * B * N * C * 12)
    if not a.only or "embed" in a.only:
        # K-EMBED at the BASELINE geometries (cfg2 batch = --B; the volumetric ones at a batch that fits one tile wave)
        for name, Bv, dims, patch, Ce in (("cfg2", B, (224, 224, 1), (16, 16, 1), 384), ("cfg1 B=2", 2, (128, 128, 64), (16, 16, 8), 1024),
                                          ("cfg5 B=8", 8, (128, 128, 128), (8, 8, 8), 512)):
            Mv = 4
            img = rnd(Bv, Mv, 1, *dims)
            P = patch[0] * patch[1] * patch[2]
            Np = (dims[0] // patch[0]) * (dims[1] // patch[1]) * (dims[2] // patch[2])
            Ne = Np + 1
            We = rnd(Ce, P) / P ** 0.5
            be, pe = rnd(Ce), rnd(Ne, Ce)
            tok = torch.zeros(Mv, Bv * Ne, Ce, device=DEV)
            dYb = rnd(Mv, Bv * Ne, Ce).to(BF)
            dWe = torch.empty(Ce, P, device=DEV)
            vol, tokb = img.numel() * 4, Mv * Bv * Np * Ce
            mem_line(f"embed fused fwd {name}", lambda: ops.embed_fused_fwd(img, We, be, pe, tok, patch_size=patch, C_=Ce), vol + tokb * 4)
            mem_line(f"embed fused wgrad {name}", lambda: ops.embed_fused_wgrad(img, dYb, dWe, patch_size=patch, C_=Ce), vol + tokb * 2)
    print("status", _abi.device_status())


if __name__ == "__main__":
    main()
