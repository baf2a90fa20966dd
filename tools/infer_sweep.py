"""BASELINE.json configs[3]: inference-only forward of the cfg2 model, batch sweep 1..1024 on one B200
(latency vs throughput). eval() + torch.no_grad(), weights frozen (no per-forward operand cast), CUDA-graph replay.
Prints one JSON line per batch size and a markdown table; run under gpurun and copy into profiles/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))

import torch  # noqa: E402

from bench import WORKLOADS, flops_per_volume  # noqa: E402
from cavit import _abi  # noqa: E402
from cavit.config import make_config  # noqa: E402
from cavit.modules import ModelCross  # noqa: E402


def main():
    _abi.require_device(0)
    wl = WORKLOADS["cfg2"]
    cfg = make_config(**wl["cfg"])
    torch.manual_seed(0)
    model = ModelCross(cfg).cuda().eval()
    model.engine().freeze_operands(True)
    D, H, W = cfg.img_size
    fpv = flops_per_volume(cfg)
    rows = []
    batches = [int(b) for b in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]
    for B in batches:
        img = torch.randn(B, cfg.num_modalities, 1, D, H, W, device="cuda")
        labels = torch.zeros(B, dtype=torch.long, device="cuda")
        with torch.no_grad():
            for _ in range(5):
                model(img, labels)
            torch.cuda.synchronize()
            iters = max(5, min(200, int(2000 / max(B, 8))))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                model(img, labels)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        assert _abi.device_status() == 0
        r = {"batch": B, "latency_ms": ms, "volumes_per_s": B / (ms * 1e-3), "model_tflops": B * fpv / (ms * 1e-3) / 1e12}
        rows.append(r)
        print(json.dumps(r), flush=True)
        del img
    print("\n| batch | latency ms | volumes/s | model TFLOP/s (fwd) |\n|---|---|---|---|")
    for r in rows:
        print(f"| {r['batch']} | {r['latency_ms']:.3f} | {r['volumes_per_s']:.0f} | {r['model_tflops']:.1f} |")


if __name__ == "__main__":
    main()
