"""Runs the attention forward/backward kernels a few times (target for ncu).
usage: one_attn.py [G B N H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))
import torch  # noqa: E402

from cavit import _abi, ops  # noqa: E402

G, B, N, H = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 256, 197, 6)
C = H * 64
T = B * N
qkv = torch.randn(G, T, 3 * C, device="cuda").to(torch.bfloat16)
o = torch.empty(G, T, C, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(G, B, H, N, device="cuda")
do = torch.randn(G, T, C, device="cuda").to(torch.bfloat16)
dqkv = torch.empty(G, T, 3 * C, device="cuda", dtype=torch.bfloat16)
delta = torch.empty(G, B, H, N, device="cuda")
acc = torch.empty(G, T, C, device="cuda")
for _ in range(3):
    ops.attn_fwd(qkv, o, lse, G=G, B=B, N=N, H=H, scale=0.125)
    ops.attn_bwd(qkv, o, do, lse, dqkv, delta, acc, G=G, B=B, N=N, H=H, scale=0.125)
torch.cuda.synchronize()
print("status", _abi.device_status())
