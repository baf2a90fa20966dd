"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — recipe for ``oracle/_ref/`` and the real-voxel fixtures.

The reference is pure Python: nothing compiles. What this recipe places into the git-ignored (but gpurun-shipped)
``oracle/_ref/`` directory, from the sources where they lie under /root/reference, is

  * the reference's own model files, byte for byte (so that on the GPU box — which has no /root/reference — the
    ``cpu_baseline`` / ``bench.py --impl reference`` legs can time the UNMODIFIED ``ModelCross``: ``kind: "reference"``).
    They never enter the git history and the product never imports them;
  * ``ucsf_cfg1_int16.npz``: BASELINE.json ``configs[0]``'s inputs — the STORED int16 voxels of the centre-crop window
    [56:184, 56:184, 45:109] of the T1, T1c, T2, FLAIR volumes of the six bundled UCSF-PDGM cases (scl_slope / scl_inter
    kept beside them; /root/reference/dataset_ucsf.py:81-89,121-158 with ``img_size`` (128, 128, 64)) and their MGMT labels
    (labels.csv, column "MGMT status": negative 0, positive 1);

and, committed because it is small, ``tests/golden/ucsf_small.pt``: the same for two cases at a (64, 64, 32) window, plus
the fp64 / fp32 outputs of the unmodified reference ``ModelCross`` (config2.py defaults, M = 4 ring) on those real voxels.

Usage (build container only):  python -m oracle.make_ref
"""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))

REF = os.environ.get("CAVIT_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ("model_cross.py", "modelv3.py", "model.py", "modelv2.py", "config.py", "config2.py", "utils.py",
         "modify_model.py", "labels.csv")
CASES = ("0085", "0279", "0381", "0392", "0451", "0516")
TYPES = ("T1", "T1c", "T2", "FLAIR")
RING4 = {"0": "1", "1": "2", "2": "3", "3": "0"}
SMALL_CFG = dict(hidden_dim=1024, mlp_dim=4096, num_heads=16, num_multi_blocks=2, num_self_blocks=2,
                 patch_size=(16, 16, 8), img_size=(64, 64, 32), num_modalities=4, attn_order=RING4, num_classes=2,
                 dropout=0.0, label_smoothing=0.0)


def copy_reference_files() -> int:
    os.makedirs(OUT, exist_ok=True)
    n = 0
    for f in FILES:
        src = os.path.join(REF, f)
        if os.path.isfile(src):
            shutil.copyfile(src, os.path.join(OUT, f))
            n += 1
    return n


def labels_of(cases):
    import pandas as pd
    d = pd.read_csv(os.path.join(REF, "labels.csv"))
    out = []
    for c in cases:   # labels.csv writes the case number without the zero padding of the folder names
        row = d[d["ID"] == f"UCSF-PDGM-{int(c):03d}"]
        assert len(row) == 1, c
        out.append(1 if row.iloc[0]["MGMT status"] == "positive" else 0)
    return np.asarray(out, dtype=np.int64)


def stored_windows(cases, img_size):
    """-> int16 [len(cases), len(TYPES), *window] in FILE axis order (i, j, k), slope / inter [cases, types]."""
    from cavit.staging import plan_batch, read_nifti
    vols = [read_nifti(os.path.join(REF, "ucsf-data", f"UCSF-PDGM-{c}_nifti", f"UCSF-PDGM-{c}_{t}.nii.gz"))
            for c in cases for t in TYPES]
    _, wins, _ = plan_batch(vols, img_size)
    ext = tuple(n for _, n in wins[0])
    data = np.zeros((len(vols),) + ext, dtype=vols[0].data.dtype)
    for i, (v, w) in enumerate(zip(vols, wins)):
        assert tuple(n for _, n in w) == ext
        data[i] = v.data.reshape(v.dims, order="F")[tuple(slice(a, a + n) for a, n in w)]
    shp = (len(cases), len(TYPES))
    return (data.reshape(shp + ext), np.asarray([v.slope for v in vols], np.float32).reshape(shp),
            np.asarray([v.inter for v in vols], np.float32).reshape(shp), wins[0])


def volumes_fp32(data, slope, inter, img_size):
    """Stored windows -> the reference's fp32 batch [B, M, 1, D, H, W] (oracle.staging: nibabel scaling, MONAI crop / pad)."""
    from oracle.staging import stage_batch
    samples = [[(data[b, m].reshape(-1, order="F"), data[b, m].shape, float(slope[b, m]), float(inter[b, m]))
                for m in range(data.shape[1])] for b in range(data.shape[0])]
    return torch.from_numpy(stage_batch(samples, img_size))


def make_small_golden():
    from oracle import ref_loader
    from oracle.full_cases import perturb_1d, sample_index
    from oracle.functional import make_config
    from oracle.weights import state_checksum
    cases = CASES[:2]
    cfg = make_config(**SMALL_CFG)
    data, slope, inter, win = stored_windows(cases, cfg.img_size)
    labels = torch.from_numpy(labels_of(cases))
    img = volumes_fp32(data, slope, inter, cfg.img_size)
    mc = ref_loader.load("model_cross")
    torch.manual_seed(0)
    model = mc.ModelCross(ref_loader.to_config_dict(cfg))
    perturb_1d(model.named_parameters(), 0)
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    rec = {"cases": cases, "types": TYPES, "cfg": SMALL_CFG, "window": win, "stored": torch.from_numpy(data.copy()),
           "slope": torch.from_numpy(slope), "inter": torch.from_numpy(inter), "labels": labels,
           "state_checksum": state_checksum(state), "img_checksum": float(img.double().sum()), "torch": torch.__version__}
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        model.load_state_dict(state)
        m = model.to(dt).train()
        for p in m.parameters():
            p.grad = None
        logits, loss = m(img.to(dt), labels)
        loss.backward()
        rec["logits" + tag], rec["loss" + tag] = logits.detach().clone(), loss.detach().clone()
        if dt == torch.float64:
            rec["grad_norm"] = {k: float(p.grad.norm()) for k, p in m.named_parameters()}
            rec["grad_sample"] = {k: p.grad.flatten()[sample_index(p.numel(), i)].to(torch.float32).clone()
                                  for i, (k, p) in enumerate(m.named_parameters())}
    path = os.path.join(ROOT, "tests", "golden", "ucsf_small.pt")
    torch.save(rec, path)
    print(f"ucsf_small: img mean {float(img.mean()):.1f} max {float(img.max()):.0f}; loss64 {float(rec['loss64']):.9f} "
          f"|logits| {float(rec['logits64'].norm()):.5f} fp32-vs-fp64 {float((rec['logits32'].double() - rec['logits64']).norm() / rec['logits64'].norm()):.2e}"
          f" -> {path} ({os.path.getsize(path)} B)")


def main(force: bool = True, golden: bool = True, quiet: bool = False):
    """force=False (what __graft_entry__.build() does): only what is missing is produced."""
    say = (lambda *a: None) if quiet else print
    if not os.path.isfile(os.path.join(REF, "model_cross.py")):
        say(f"make_ref: {REF} not present, nothing to do")
        return
    n = copy_reference_files()
    say(f"make_ref: {n} reference files -> {OUT}")
    path = os.path.join(OUT, "ucsf_cfg1_int16.npz")
    if force or not os.path.exists(path):
        data, slope, inter, win = stored_windows(CASES, (128, 128, 64))
        np.savez_compressed(path, stored=data, slope=slope, inter=inter, labels=labels_of(CASES), window=np.asarray(win),
                            cases=np.asarray(CASES), types=np.asarray(TYPES))
        say(f"make_ref: {data.shape} {data.dtype} window {win} -> {path} ({os.path.getsize(path) >> 20} MiB)")
    if golden and (force or not os.path.exists(os.path.join(ROOT, "tests", "golden", "ucsf_small.pt"))):
        make_small_golden()


if __name__ == "__main__":
    main(force=True, golden="--no-golden" not in sys.argv)
