import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import functional as OF
from oracle.cases import build_case
from test_gpu_dropout import export_masks, rel
from cavit.modules import ModelCross, ModelVIT
for name in ["cross_noattn_h1", "cross_ring4", "vit_small"]:
    for p in [0.0, 0.1, 0.5]:
        kind, cfg, state, img, labels = build_case(name)
        cfg.dropout = p
        torch.manual_seed(123)
        model = (ModelCross if kind == "cross" else ModelVIT)(cfg)
        model.load_state_dict(state, strict=True)
        model = model.cuda().train()
        for step in range(5):
            model.zero_grad(set_to_none=True)
            logits, loss = model(img.cuda(), labels.cuda())
            loss.backward()
            eng = model.engine()
            dm = export_masks(eng, kind, cfg, img.shape[0]) if p > 0 else None
            rl, rloss, rg = OF.forward_backward(state, img, labels, cfg, kind, torch.float64, dm=dm)
            te = sum(float((prm.grad.double().cpu() - rg[k]).norm()) ** 2 for k, prm in model.named_parameters())
            tr = sum(float(rg[k].norm()) ** 2 for k in rg)
            print(name, p, step, "logits rel %.4f" % rel(logits, rl), "loss %.5f vs %.5f" % (float(loss), float(rloss)), "grad rel %.4f" % ((te / tr) ** 0.5), flush=True)
