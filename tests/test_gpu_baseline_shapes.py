"""GPU parity at the EXACT BASELINE.json model shapes (cfg2, cfg1, cfg3, cfg5, and ModelVIT on the cfg2
slices with N = 4*196 + 1) at a reduced batch, against outputs of the UNMODIFIED reference frozen by
oracle/gen_golden_full.py into tests/golden/full_<case>.pt (fp64 run of /root/reference/model_cross.py:186-212 /
modelv3.py:123-147 on the same seeded weights and inputs).

Error definitions: logits  ||ours - ref||_2 / ||ref||_2 over the batch; gradients the same norm ratio per
parameter tensor on the frozen sample of positions (all positions for tensors <= 1024 elements) and over the concatenation
of all samples, plus the full-tensor norm of every gradient against the frozen reference norm.

Tolerances — bf16 mode (north_star "about 2e-2"): logits <= 2.5e-2, whole-gradient <= 2.5e-2; per tensor <= 5e-2 for tensors
carrying more than 1e-3 of the largest gradient norm (small tensors see the same absolute noise against a smaller norm; the
one tensor above 3e-2 is ModelVIT's pos_embedding, 2.2e-2 - 4.2e-2 from run to run).
cfg2_b2 is the exception on logits: the four reference logits there have ||logits||_2 = 0.088, and the bf16 rounding noise of
the network (2e-3 - 4e-3 absolute) is 2e-2 - 4.5e-2 of that: five runs with the input scaled by 1 + {0, +-1e-6, +-3e-6} give
1.9e-2 ... 4.5e-2 with either embedding path (tools/logit_noise_floor.py, profiles/logit_noise_floor_r02.txt). Its bound is
therefore absolute, ||ours - ref||_2 <= 5e-3, with the relative figure reported.
Measured on a B200 (round 2): logits 1.9e-2 - 4.0e-2 (cfg2), 1.2e-2 (cfg1), 1.8e-3 (cfg3), 1.1e-2 (cfg5), 8.0e-3 (ModelVIT);
whole gradient 0.8e-2 / 1.5e-2 / 1.0e-2 / 2.2e-2 / 0.7e-2; worst tensor 3.0e-2 (cfg5).
fp32 mode (north_star "about 1e-3"): logits <= 1e-3, whole-gradient <= 2e-3, per tensor <= 5e-3."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.full_cases import FULL_CASES, build_full_case, sample_index   # noqa: E402
from oracle.weights import state_checksum                                  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REPORT = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out", "baseline_shape_parity.jsonl")


def _measure(name, precision):
    from cavit import _abi
    from cavit.config import make_config
    from cavit.modules import ModelCross, ModelVIT
    path = os.path.join(GOLD, "full_" + name + ".pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    rec = torch.load(path, weights_only=False)
    kind, cfg, model, img, labels = build_full_case(name, ModelCross, ModelVIT, make_config)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cs = state_checksum(state)
    assert abs(cs - rec["state_checksum"]) <= 1e-9 * abs(rec["state_checksum"]), "seeded weights differ from the generator's"
    assert abs(float(img.double().sum()) - rec["img_checksum"]) <= 1e-9 * max(1.0, abs(rec["img_checksum"]))
    assert torch.equal(labels, rec["labels"])
    if precision != "bf16":
        model.set_precision(precision)
    model = model.cuda().train()
    logits, loss = model(img.cuda(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert _abi.device_status() == 0
    ref = rec["logits64"]
    out = {"case": name, "precision": precision,
           "logits_rel": float((logits.detach().double().cpu() - ref).norm() / ref.norm()),
           "logits_abs": float((logits.detach().double().cpu() - ref).norm()), "logits_ref_norm": float(ref.norm()),
           "loss_abs": abs(float(loss) - float(rec["loss64"])), "per_tensor": {}}
    gmax = max(rec["grad_norm"].values())
    num = den = 0.0
    worst, worst_norm = ("", 0.0), ("", 0.0)
    for i, (k, p) in enumerate(model.named_parameters()):
        g = p.grad.detach()
        ref_s = rec["grad_sample"][k].double()
        got_s = g.flatten()[sample_index(g.numel(), i).to(g.device)].double().cpu()
        w = g.numel() / ref_s.numel()              # every sample stands for numel / samples elements
        e2, r2 = float((got_s - ref_s).norm()) ** 2, float(ref_s.norm()) ** 2
        num += w * e2
        den += w * r2
        gn = rec["grad_norm"][k]
        if gn > 1e-3 * gmax:
            r = (e2 / max(r2, 1e-300)) ** 0.5
            if r > worst[1]:
                worst = (k, r)
            rn = abs(float(g.double().norm()) - gn) / gn
            if rn > worst_norm[1]:
                worst_norm = (k, rn)
        else:   # analytically ~zero gradients (fusion wk.bias) must stay negligible
            assert float(g.double().norm()) < 2e-3 * gmax + 1e-6, k
    out.update(grad_rel=(num / den) ** 0.5, worst_tensor=worst[0], worst_tensor_rel=worst[1],
               worst_norm_tensor=worst_norm[0], worst_norm_rel=worst_norm[1])
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(out) + "\n")
    except OSError:
        pass
    print(json.dumps(out))
    del model
    torch.cuda.empty_cache()
    return out


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_baseline_shape_bf16_mode_matches_reference(name):
    m = _measure(name, "bf16")
    if name == "cfg2_b2":     # ||ref logits|| = 0.088: the bf16 noise floor is 2e-2 - 4.5e-2 of it (module docstring)
        assert m["logits_abs"] < 5e-3 and m["logits_rel"] < 6e-2, m
    else:
        assert m["logits_rel"] < 2.5e-2, m
    assert m["loss_abs"] < 2e-3, m
    assert m["grad_rel"] < 2.5e-2, m
    assert m["worst_tensor_rel"] < 5e-2, m
    assert m["worst_norm_rel"] < 4e-2, m


@pytest.mark.parametrize("name", list(FULL_CASES))
def test_baseline_shape_fp32_mode_matches_reference(name):
    m = _measure(name, "fp32")
    assert m["logits_rel"] < 1e-3, m
    assert m["loss_abs"] < 1e-4, m
    assert m["grad_rel"] < 2e-3, m
    assert m["worst_tensor_rel"] < 5e-3, m
