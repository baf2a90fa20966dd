"""GPU parity with dropout > 0 in training mode (nn.Dropout sites of /root/reference/model_cross.py:25,27,47,
84,86,170,180,182 and modelv3.py).

ATen's Philox stream cannot be reproduced by another implementation, so parity is checked by REPLAY: the
counter-based masks the CUDA path used for this step (exported with cavit_dropout mode 4 from the step's
device seed and the site ids) are fed to the fp64 oracle, whose dropout call sites are pinned against the
reference modules in tests/test_oracle_vs_reference.py::test_dropout_sites_match_reference.
Tolerances are the bf16-mode ones of tests/test_gpu_model.py (2e-2 on logits, 3e-2 on the gradient vector);
because a dropped-out step can leave the 2x2 logits of these tiny cases close to zero (loss ~ ln 2), the logits
error is measured against max(||ref||, sqrt(numel)), i.e. 2e-2 relative with a 2e-2-per-element absolute floor."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import functional as OF                      # noqa: E402
from oracle.cases import CASES, build_case               # noqa: E402


def _mask(eng, site, shape):
    from cavit import ops
    n = 1
    for s in shape:
        n *= s
    out = torch.empty(n, dtype=torch.uint8, device=eng.device)
    ops.dropout(ops.DROP_MASK, None, None, out, n=n, p=eng.p_drop, seed=eng.seed_buf, site=site)
    return (out.view(shape).double() / (1.0 - eng.p_drop)).cpu()


def export_masks(eng, kind, cfg, B):
    """Multiplier tensors of every dropout site of the step the engine just ran, keyed like oracle.functional."""
    G, N, C, F, H, K = eng.G, eng.N, eng.C, eng.F, eng.H, eng.K
    dm = {}
    emb = _mask(eng, eng.SITE_EMBED, (G, B, N, C))
    for m in range(G):
        dm[("embed", m)] = emb[m]
    for l in range(eng.L):
        for which, width in (("out", C), ("gelu", F), ("fc2", C)):
            mk = _mask(eng, eng.site_layer(l, which), (G, B, N, width))
            for m in range(G):
                dm[(which, l, m)] = mk[m]
    if kind == "cross":
        for mb in range(cfg.num_multi_blocks):
            if not K:
                break
            shapes = {"attn": (K, B, H, 1, N), "proj": (K, B, 1, C), "gelu": (K, B, 1, F), "fc2": (K, B, 1, C)}
            for which, shp in shapes.items():
                mk = _mask(eng, eng.site_fusion(mb, which), shp)
                for k in range(K):
                    dm[("f_" + which, mb, k)] = mk[k]
    hg = _mask(eng, eng.SITE_HEAD_GELU, (G, B, F))
    hl = _mask(eng, eng.SITE_HEAD_LOGITS, (G, B, cfg.num_classes))
    for m in range(G):
        dm[("head_gelu", m)] = hg[m]
        dm[("head_logits", m)] = hl[m]
    return dm


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rel_floor(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), b.numel() ** 0.5))


@pytest.mark.parametrize("name", ["cross_ring4", "cross_chain3", "cross_noattn_h1", "vit_small"])
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_training_step_matches_oracle_with_replayed_masks(name, p):
    from cavit import _abi
    from cavit.modules import ModelCross, ModelVIT
    kind, cfg, state, img, labels = build_case(name)
    cfg.dropout = p
    torch.manual_seed(123)
    model = (ModelCross if kind == "cross" else ModelVIT)(cfg)
    model.load_state_dict(state, strict=True)
    model = model.cuda().train()
    B = img.shape[0]
    for step in range(4):            # steps 0,1 run eagerly, step 2 captures the CUDA graphs, step 3 replays them
        model.zero_grad(set_to_none=True)
        logits, loss = model(img.cuda(), labels.cuda())
        loss.backward()
        assert _abi.device_status() == 0
        eng = model.engine()
        dm = export_masks(eng, kind, cfg, B)
        drop_frac = float((dm[("embed", 0)] == 0).double().mean())
        assert abs(drop_frac - p) < 0.05, drop_frac
        ref_logits, ref_loss, ref_grads = OF.forward_backward(state, img, labels, cfg, kind, torch.float64, dm=dm)
        assert rel_floor(logits, ref_logits) < 2e-2, step
        assert abs(float(loss.detach()) - float(ref_loss)) < 2e-2 * max(1.0, abs(float(ref_loss))), step
        tot_err, tot_ref = 0.0, 0.0
        for k, prm in model.named_parameters():
            d = prm.grad.double().cpu() - ref_grads[k]
            tot_err += float(d.norm()) ** 2
            tot_ref += float(ref_grads[k].norm()) ** 2
        assert (tot_err / tot_ref) ** 0.5 < 3e-2, step
        if step == 0:
            first = {k: v.clone() for k, v in dm.items()}
    # a fresh mask every step, also under graph replay (the seed is a device scalar the graph reads)
    assert not torch.equal(first[("embed", 0)], dm[("embed", 0)])


def test_eval_mode_ignores_dropout_and_masks_are_reproducible():
    from cavit.modules import ModelCross
    kind, cfg, state, img, labels = build_case("cross_chain3")
    cfg.dropout = 0.3
    runs = []
    for _ in range(2):
        torch.manual_seed(7)
        model = ModelCross(cfg)
        model.load_state_dict(state)
        model = model.cuda().train()
        logits, loss = model(img.cuda(), labels.cuda())
        runs.append(logits.clone())
    assert torch.equal(runs[0], runs[1])                  # torch.manual_seed controls the dropout stream
    model.eval()
    with torch.no_grad():
        l_eval, _ = model(img.cuda(), labels.cuda())
    ref_logits, _ = OF.model_cross_forward({k: v.double() for k, v in state.items()}, img.double(), labels, cfg)
    assert rel(l_eval, ref_logits) < 2e-2                 # identity in eval(), like nn.Dropout
    assert not torch.equal(l_eval, runs[0])


def test_dropout_mask_statistics():
    """keep-rate and pairwise independence of the counter-based generator across sites / seeds."""
    from cavit import ops
    n = 1 << 20
    seed = torch.tensor([0x1234567], dtype=torch.int64, device="cuda")
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    masks = []
    for site in (1, 2, 4096):
        ops.dropout(ops.DROP_MASK, None, None, out, n=n, p=0.3, seed=seed, site=site)
        masks.append(out.clone().float())
        assert abs(float(masks[-1].mean()) - 0.7) < 3e-3
    seed.fill_(0x1234568)
    ops.dropout(ops.DROP_MASK, None, None, out, n=n, p=0.3, seed=seed, site=1)
    masks.append(out.clone().float())
    for i in range(len(masks)):
        for j in range(i + 1, len(masks)):
            a, b = masks[i] - masks[i].mean(), masks[j] - masks[j].mean()
            corr = float((a * b).mean() / (a.std() * b.std()))
            assert abs(corr) < 5e-3, (i, j, corr)
    # neighbouring elements (the two halves of one hash) are independent too
    a, b = masks[0][0::2] - 0.7, masks[0][1::2] - 0.7
    assert abs(float((a * b).mean() / (a.std() * b.std()))) < 5e-3
