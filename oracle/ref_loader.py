"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — import the UNMODIFIED reference.

The reference's model files import third-party packages that are absent from this image
(``lightning``, ``ml_collections``, ``torchmetrics``). This loader places
minimal stub modules in ``sys.modules`` and imports ``model_cross`` / ``modelv3`` straight
from the reference tree: /root/reference in the build container, or — on the GPU box, which has
no /root/reference — the byte-for-byte copies that ``oracle/make_ref.py`` placed into the
git-ignored ``oracle/_ref/`` (used there by the timed CPU-baseline legs of bench.py only).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch.nn as nn

# /root/reference in the build container; on the GPU box the byte-for-byte copies that oracle/make_ref.py placed
# into the git-ignored, gpurun-shipped oracle/_ref/ (test infrastructure: cpu_baseline / --impl reference only)
REFERENCE_DIRS = [os.environ.get("CAVIT_REFERENCE_DIR", ""), "/root/reference",
                  os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]


def reference_dir():
    for d in REFERENCE_DIRS:
        if d and os.path.isfile(os.path.join(d, "model_cross.py")):
            return d
    return None


class ConfigDict(dict):
    """Stand-in for ml_collections.ConfigDict: attribute access, nested dict wrapping."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v


def _install_stubs():
    if "lightning" not in sys.modules:
        L = types.ModuleType("lightning")

        class LightningModule(nn.Module):
            def log(self, *a, **k):
                return None

        L.LightningModule = LightningModule
        sys.modules["lightning"] = L
    if "ml_collections" not in sys.modules:
        mc = types.ModuleType("ml_collections")
        mc.ConfigDict = ConfigDict
        sys.modules["ml_collections"] = mc
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tmf = types.ModuleType("torchmetrics.functional")
        tmc = types.ModuleType("torchmetrics.classification")
        for nm in ("BinaryAccuracy", "BinaryPrecision", "BinaryRecall", "BinarySpecificity",
                   "BinaryF1Score", "BinaryConfusionMatrix"):
            setattr(tmc, nm, type(nm, (), {}))
        tm.functional = tmf
        tm.classification = tmc
        sys.modules["torchmetrics"] = tm
        sys.modules["torchmetrics.functional"] = tmf
        sys.modules["torchmetrics.classification"] = tmc


def _install_encoder_stubs():
    """Extra stubs for model.py / modelv2.py: MONAI's DenseNet121 (only used with
    pretrained_cnn=True) and the import-time dataset construction of model.py:15,338-344."""
    if "monai" not in sys.modules:
        mo, mn, mnn = types.ModuleType("monai"), types.ModuleType("monai.networks"), types.ModuleType("monai.networks.nets")
        mnn.DenseNet121 = object
        mo.networks, mn.nets = mn, mnn
        sys.modules.update({"monai": mo, "monai.networks": mn, "monai.networks.nets": mnn})
    if "dataset_ucsf" not in sys.modules:
        ds = types.ModuleType("dataset_ucsf")

        class BrainDataset:
            def __init__(self, *a, **k):
                pass

        ds.BrainDataset = BrainDataset
        sys.modules["dataset_ucsf"] = ds


def load(module_name: str):
    """Import `module_name` (e.g. 'model_cross', 'modelv3') from the reference tree."""
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("reference tree not found (set CAVIT_REFERENCE_DIR)")
    _install_stubs()
    if d not in sys.path:
        sys.path.insert(0, d)
    if module_name in ("model", "modelv2"):
        _install_encoder_stubs()
    if module_name == "model":   # model.py:338 reads labels.csv relative to the working directory at import time
        cwd = os.getcwd()
        os.chdir(d)
        try:
            return importlib.import_module(module_name)
        finally:
            os.chdir(cwd)
    return importlib.import_module(module_name)


def to_config_dict(cfg) -> ConfigDict:
    """Attribute bag -> ConfigDict stub (what the reference constructors expect)."""
    out = ConfigDict()
    for k, v in vars(cfg).items():
        setattr(out, k, v)
    return out
