"""SURVEY.md §8d "GPU before" number: the reference's formulation (the oracle port of model_cross.py, i.e. the same ATen ops
the reference issues: einsum-free matmul / softmax / layer_norm / gelu, materialised scores) run EAGERLY on the same B200,
in fp32 and under torch.autocast(bfloat16), fwd+bwd on the bench workload. Context for the bench line, not a bench arm:
   python tools/gpu_eager_baseline.py [workload] [batch]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))

import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from cavit.config import make_config  # noqa: E402
from cavit.modules import ModelCross  # noqa: E402
from oracle import functional as OF  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    wl = WORKLOADS[name]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else wl["batch"]
    cfg = make_config(**wl["cfg"])
    torch.manual_seed(0)
    state = {k: v.detach().clone() for k, v in ModelCross(cfg).state_dict().items()}   # reference initialiser
    D, H, W = cfg.img_size
    img = torch.randn(B, cfg.num_modalities, 1, D, H, W, device="cuda")
    labels = torch.randint(0, cfg.num_classes, (B,), device="cuda")
    out = {"workload": name, "batch": B}
    for mode in ("fp32", "autocast_bf16"):
        params = {k: v.cuda().requires_grad_(True) for k, v in state.items()}

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode != "fp32")):
                logits, loss = OF.model_cross_forward(params, img, labels, cfg)
            loss.backward()
            for p in params.values():
                p.grad = None
            return loss

        try:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 3
            for _ in range(n):
                loss = step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[mode] = {"ms_per_step": ms, "volumes_per_s": B / (ms * 1e-3), "loss": float(loss),
                         "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        except torch.cuda.OutOfMemoryError as ex:   # noqa: PERF203
            out[mode] = {"error": "out of memory: " + str(ex)[:80]}
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
