"""Per-shape GEMM timings inside one real cfg2 training step (CUDA events around every launch)."""
import os
import sys
from collections import defaultdict

os.environ["CAVIT_NO_GRAPHS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from cavit import ops  # noqa: E402
from cavit.config import make_config  # noqa: E402
from cavit.modules import ModelCross  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B = wl["batch"]
cfg = make_config(**wl["cfg"])
torch.manual_seed(0)
model = ModelCross(cfg).cuda().train()
D, H, W = cfg.img_size
img = torch.randn(B, cfg.num_modalities, 1, D, H, W, device="cuda")
labels = torch.randint(0, cfg.num_classes, (B,), device="cuda")


def step():
    logits, loss = model(img, labels)
    loss.backward()
    for p in model.parameters():
        p.grad = None


for _ in range(3):
    step()
agg = defaultdict(lambda: [0.0, 0, 0.0])
for rep in range(3):
    ops.PROFILE = []
    step()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    for name, a, b, info in prof:
        if name == "gemm":
            e = agg[info["shape"]]
            e[0] += a.elapsed_time(b) / 3
            e[1] += 1
            e[2] = info["flops"]
tot = 0.0
print("   M      N      K   G a b    n/step   ms/call   ms/step   TFLOP/s")
for shape, (ms, n, fl) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    n //= 3
    tot += ms
    print("%6d %6d %6d %3d %d %d   %4d   %8.3f  %8.3f  %8.1f" % (*shape, n, ms / n, ms, fl / (ms / n) / 1e9))
print("total gemm ms/step", tot)
