"""Runs a single GEMM shape a few times (target for ncu captures)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))
import torch  # noqa: E402

from cavit import _abi, ops  # noqa: E402
from cavit._abi import EPI_BIAS_GELU, EPI_NONE  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "qkv"
G, T, C, F = 4, 256 * 197, 384, 1536
x = torch.randn(G, T, C, device="cuda").to(torch.bfloat16)
if kind == "qkv":
    w = (torch.randn(G, 3 * C, C, device="cuda") / math.sqrt(C)).to(torch.bfloat16)
    out = torch.empty(G, T, 3 * C, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.linear_fwd(x, w, out)  # noqa: E731
else:
    w = (torch.randn(G, F, C, device="cuda") / math.sqrt(C)).to(torch.bfloat16)
    b = torch.randn(G, F, device="cuda")
    out = torch.empty(G, T, F, device="cuda", dtype=torch.bfloat16)
    u = torch.empty(G, T, F, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.linear_fwd(x, w, out, epi=EPI_BIAS_GELU, bias=b, aux=u)  # noqa: E731
for _ in range(5):
    fn()
torch.cuda.synchronize()
print("status", _abi.device_status())
