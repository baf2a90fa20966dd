"""Model configuration bag — the host-side mirror of the reference's config objects.

The reference builds an ``ml_collections.ConfigDict`` in ``get_mgmt_config()``
(/root/reference/config2.py:3-28) and overlays the hyper-parameters of its drivers with
``modify_config`` (/root/reference/config2.py:30-35, /root/reference/main_mist.py:150-170).
``ml_collections`` is not a dependency here: the model classes only read attributes, so any
attribute bag works (an ``ml_collections.ConfigDict`` included). ``make_config`` returns a
``types.SimpleNamespace`` with the same field names and the reference's defaults.
"""
from __future__ import annotations

from types import SimpleNamespace

RING4 = {"0": "1", "1": "2", "2": "3", "3": "0"}


def get_mgmt_config() -> SimpleNamespace:
    """Defaults of /root/reference/config2.py:3-28 (fields the transformer path reads)."""
    return SimpleNamespace(hidden_dim=1024, mlp_dim=4096, num_heads=16, num_multi_blocks=2, num_self_blocks=2,
                           patch_size=(16, 16, 8), num_classes=2, img_size=(128, 128, 64), in_channels=1,
                           spacing=(2, 2, 2), target="MGMT status")


def modify_config(config, params):
    """/root/reference/config2.py:30-35: overlay a dict / namedtuple of values on a config."""
    if not isinstance(params, dict):
        params = params._asdict()
    for key, value in params.items():
        setattr(config, key, value)
    return config


def make_config(**kw) -> SimpleNamespace:
    """get_mgmt_config() + the driver-side fields the model constructors read
    (/root/reference/model_cross.py:153-183, /root/reference/main_mist.py:150-170), overridden by kw."""
    cfg = get_mgmt_config()
    modify_config(cfg, dict(num_layers=4, dropout=0.0, lr=1e-4, weight_decay=5e-4,
                            optim_params={"T_max": 250, "eta_min": 1e-6}, label_smoothing=0.0,
                            num_modalities=4, attn_order=dict(RING4)))
    return modify_config(cfg, kw)
