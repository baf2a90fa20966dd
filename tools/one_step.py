"""One cfg2 training step for ncu (eager launches, CUDA graphs off). The measured step is bracketed by
cudaProfilerStart/Stop:   ncu --profile-from-start off ... python tools/one_step.py [workload] [batch]"""
import os
import sys

os.environ["CAVIT_NO_GRAPHS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from cavit import _abi  # noqa: E402
from cavit.config import make_config  # noqa: E402
from cavit.modules import ModelCross  # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["batch"]
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
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = _abi.launch_count()
torch.cuda.profiler.start()
loss = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in the profiled step:", _abi.launch_count() - n0, "loss", float(loss), "status", _abi.device_status())
