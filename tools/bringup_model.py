"""GPU bring-up of the whole model path: prints errors per parameter vs the fp64 oracle."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))

import torch  # noqa: E402

from cavit import _abi  # noqa: E402
from cavit.modules import ModelCross, ModelVIT  # noqa: E402
from oracle import functional as OF  # noqa: E402
from oracle.cases import CASES, build_case  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


for name in (sys.argv[1:] or list(CASES)):
    try:
        kind, cfg, state, img, labels = build_case(name)
        model = (ModelCross if kind == "cross" else ModelVIT)(cfg)
        model.load_state_dict(state)
        model = model.cuda().train()
        logits, loss = model(img.cuda(), labels.cuda())
        torch.cuda.synchronize()
        print(f"== {name}: fwd status {_abi.device_status()}", flush=True)
        ref_logits, ref_loss, ref_grads = OF.forward_backward(state, img, labels, cfg, kind, torch.float64)
        print(f"   logits rel {rel(logits, ref_logits):.3e}  loss {float(loss):.6f} vs {float(ref_loss):.6f}", flush=True)
        if kind == "cross":
            _, _, toks = OF.model_cross_forward({k: v.double() for k, v in state.items()}, img.double(), labels, cfg,
                                                return_tokens=True)
            xf = model.engine()._x_fin.view(cfg.num_modalities, img.shape[0], -1, cfg.hidden_dim)
            for m in range(cfg.num_modalities):
                print(f"   final tokens stream {m}: rel {rel(xf[m], toks[m]):.3e} cls {rel(xf[m][:, 0], toks[m][:, 0]):.3e}")
        loss.backward()
        torch.cuda.synchronize()
        print(f"   bwd status {_abi.device_status()}", flush=True)
        worst = []
        for k, p in model.named_parameters():
            g = ref_grads[k]
            worst.append((rel(p.grad, g), k, float(g.norm())))
        worst.sort(reverse=True)
        for r, k, n in worst[:12]:
            print(f"   grad {k}: rel {r:.3e} (|g|={n:.3e})")
        tot = sum(float((p.grad.double().cpu() - ref_grads[k]).norm()) ** 2 for k, p in model.named_parameters())
        ref = sum(float(g.norm()) ** 2 for g in ref_grads.values())
        print(f"   total grad rel {(tot / ref) ** 0.5:.3e}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"== {name}: EXCEPTION {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()
