"""Debug helper: run the encoder parity cases eagerly and print the norm of every saved buffer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))
os.environ["CAVIT_NO_GRAPHS"] = "1"
import torch
from oracle import encoders as E
from cavit.encoders import ViT, ViT3D

for name in sys.argv[1:] or list(E.ENC_CASES):
    kind, _, ctor, B, M, sseed, _ = E.ENC_CASES[name]
    cfg = E.enc_config(name)
    model = ViT(cfg) if kind == "cnnvit" else ViT3D({}, 1e-4, 0.0, M, cfg, **ctor)
    schema = {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}
    state = E.make_state_generic(schema, sseed)
    model.load_state_dict(state)
    model = model.cuda().train()
    x, labels = E.enc_inputs(name)
    logits, loss = model(x.cuda(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print(name, "logits", logits, "loss", float(loss))
    eng = model._engine_obj
    for k, v in eng.a.items():
        vs = v if isinstance(v, list) else [v]
        for i, t in enumerate(vs):
            if torch.is_tensor(t) and t.is_floating_point():
                f = t.float()
                print(f"  {k}[{i}] shape {tuple(t.shape)} norm {float(f.norm()):.4e} nan {int(torch.isnan(f).sum())}")
    rl, rs, rg = E.enc_forward_backward(name, state, x, labels)
    print("  ref logits", rl, "loss", float(rs))
    for k, p in model.named_parameters():
        g = rg[k]
        d = (p.grad.double().cpu() - g).norm() / g.norm().clamp_min(1e-30)
        print(f"  grad {k:60s} rel {float(d):.3e} |g| {float(g.norm()):.3e}")
