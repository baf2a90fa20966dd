"""GPU bring-up diagnostics (run under gpurun): exercises each kernel family on small shapes and
prints max / relative errors and the kernel-side status word instead of asserting, so that one
call gives a full picture. Not part of the product or the test-suite."""
import math
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))

import torch  # noqa: E402

from cavit import _abi, ops  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(t):
    return t.to(torch.bfloat16)


def run(name, fn):
    t0 = time.time()
    try:
        msg = fn()
        torch.cuda.synchronize()
        st = _abi.device_status()
        print(f"[{name}] {msg} status={st} ({time.time() - t0:.2f}s)", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[{name}] EXCEPTION {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def gemm_case(G, T, N, K, kind):
    def f():
        torch.manual_seed(0)
        if kind == "fwd":
            x = bf(torch.randn(G, T, K, device=DEV)); w = bf(torch.randn(G, N, K, device=DEV) / math.sqrt(K))
            want = torch.einsum("gtk,gnk->gtn", x.float(), w.float())
            out = torch.full((G, T, N), float("nan"), device=DEV)
            ops.linear_fwd(x, w, out)
        elif kind == "dgrad":
            dy = bf(torch.randn(G, T, N, device=DEV)); w = bf(torch.randn(G, N, K, device=DEV) / math.sqrt(N))
            want = torch.einsum("gtn,gnk->gtk", dy.float(), w.float())
            out = torch.full((G, T, K), float("nan"), device=DEV)
            ops.linear_dgrad(dy, w, out)
        else:
            dy = bf(torch.randn(G, T, N, device=DEV)); x = bf(torch.randn(G, T, K, device=DEV) / math.sqrt(T))
            want = torch.einsum("gtn,gtk->gnk", dy.float(), x.float())
            out = torch.full((G, N, K), float("nan"), device=DEV)
            ops.linear_wgrad(dy, x, out)
        torch.cuda.synchronize()
        nan = int(torch.isnan(out).sum())
        return f"G={G} T={T} N={N} K={K} rel={rel(torch.nan_to_num(out), want):.3e} nan={nan}"
    return f


def attn_case(G, B, N, H, bwd=False):
    def f():
        torch.manual_seed(1)
        C = H * 64
        qkv = bf(torch.randn(G, B * N, 3 * C, device=DEV))
        x = qkv.double().view(G, B, N, 3, H, 64).requires_grad_(True)
        q, k, v = x[:, :, :, 0], x[:, :, :, 1], x[:, :, :, 2]
        s = torch.einsum("gbqhd,gbkhd->gbhqk", q, k) * 64 ** -0.5
        l = torch.logsumexp(s, dim=-1)
        o = torch.einsum("gbhqk,gbkhd->gbqhd", torch.softmax(s, dim=-1), v).reshape(G, B * N, C)
        if not bwd:
            out = torch.full((G, B * N, C), float("nan"), device=DEV, dtype=torch.bfloat16)
            lse = torch.full((G, B, H, N), float("nan"), device=DEV)
            ops.attn_fwd(qkv, out, lse, G=G, B=B, N=N, H=H, scale=64 ** -0.5)
            torch.cuda.synchronize()
            return (f"G={G} B={B} N={N} H={H} rel_o={rel(torch.nan_to_num(out.float()), o):.3e} "
                    f"rel_lse={rel(torch.nan_to_num(lse), l):.3e} nan={int(torch.isnan(out.float()).sum())}")
        dout = bf(torch.randn(G, B * N, C, device=DEV))
        o.backward(dout.double())
        want = x.grad.reshape(G, B * N, 3, C)
        dqkv = torch.full((G, B * N, 3 * C), float("nan"), device=DEV, dtype=torch.bfloat16)
        delta = torch.empty(G, B, H, N, device=DEV)
        acc = torch.empty(G, B * N, C, device=DEV)
        ops.attn_bwd(qkv, bf(o.detach().float()), dout, l.detach().float().contiguous(), dqkv, delta, acc, G=G, B=B, N=N, H=H,
                     scale=64 ** -0.5)
        torch.cuda.synchronize()
        got = torch.nan_to_num(dqkv.float()).view(G, B * N, 3, C)
        return (f"G={G} B={B} N={N} H={H} rel_dq={rel(got[:, :, 0], want[:, :, 0]):.3e} "
                f"rel_dk={rel(got[:, :, 1], want[:, :, 1]):.3e} rel_dv={rel(got[:, :, 2], want[:, :, 2]):.3e}")
    return f


if __name__ == "__main__":
    _abi.require_device(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    print(torch.cuda.get_device_name(0), flush=True)
    for kind in ("fwd", "dgrad", "wgrad"):
        for shp in [(1, 128, 128, 64), (1, 128, 128, 256), (1, 256, 256, 128), (2, 200, 384, 192), (1, 77, 136, 72),
                    (4, 1026, 1536, 384)]:
            run(f"gemm-{kind}", gemm_case(*shp, kind))
    for shp in [(1, 1, 128, 1), (1, 2, 64, 2), (2, 2, 197, 3), (1, 1, 513, 2)]:
        run("attn-fwd", attn_case(*shp))
    for shp in [(1, 1, 128, 1), (1, 2, 64, 2), (2, 2, 197, 3), (1, 1, 513, 2)]:
        run("attn-bwd", attn_case(*shp, bwd=True))
