"""CPU: the drop-in module classes keep the reference's state_dict schema / init, and the flat
parameter layout of the engine covers every parameter exactly once."""
import pytest
import torch

from oracle.cases import CASES
from oracle.functional import make_config
from oracle.weights import state_schema_cross, state_schema_vit


def _cfgs():
    for name, (kind, kw, *_rest) in CASES.items():
        yield name, kind, make_config(**kw)


@pytest.mark.parametrize("name,kind,cfg", list(_cfgs()))
def test_state_dict_schema(name, kind, cfg):
    from cavit.modules import ModelCross, ModelVIT
    model = (ModelCross if kind == "cross" else ModelVIT)(cfg)
    schema = state_schema_cross(cfg) if kind == "cross" else state_schema_vit(cfg)
    sd = model.state_dict()
    assert list(sd.keys()) == list(schema.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(schema[k]), k


@pytest.mark.parametrize("name,kind,cfg", list(_cfgs()))
def test_flat_layout_covers_every_parameter_once(name, kind, cfg):
    from cavit.engine import build_layout
    schema = state_schema_cross(cfg) if kind == "cross" else state_schema_vit(cfg)
    lay = build_layout(kind, cfg)
    assert set(lay.slots) == set(schema)
    spans = []
    for k, (off, shp) in lay.slots.items():
        n = 1
        for s in shp:
            n *= s
        assert tuple(shp) == tuple(schema[k]), k
        assert off % 8 == 0 or len(shp) == 1, k   # GEMM operands (bf16 TMA) need 16-byte alignment
        spans.append((off, off + n))
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    assert spans[-1][1] <= lay.total
    # layer ranges tile the buffer
    rs = sorted((s, e) for _, s, e in lay.layer_ranges)
    assert rs[0][0] == 0 and rs[-1][1] == lay.total
    for (a0, a1), (b0, b1) in zip(rs, rs[1:]):
        assert a1 == b0


def test_same_seed_gives_reference_init():
    from oracle import ref_loader
    if ref_loader.reference_dir() is None:
        pytest.skip("reference tree not present")
    from cavit.modules import ModelCross, ModelVIT
    for kind, kw in [("cross", CASES["cross_chain3"][1]), ("vit", CASES["vit_small"][1]),
                     ("cross", CASES["cross_noattn_h1"][1])]:
        cfg = make_config(**kw)
        mod = ref_loader.load("model_cross" if kind == "cross" else "modelv3")
        torch.manual_seed(123)
        ref = (mod.ModelCross if kind == "cross" else mod.ModelVIT)(ref_loader.to_config_dict(cfg))
        torch.manual_seed(123)
        ours = (ModelCross if kind == "cross" else ModelVIT)(cfg)
        rsd, osd = ref.state_dict(), ours.state_dict()
        assert list(rsd.keys()) == list(osd.keys())
        for k in rsd:
            assert torch.equal(rsd[k], osd[k]), k


def test_cpu_forward_fails_loudly():
    from cavit import CavitError
    from cavit.modules import ModelCross
    cfg = make_config(**CASES["cross_chain3"][1])
    model = ModelCross(cfg)
    img = torch.zeros(1, cfg.num_modalities, 1, *cfg.img_size)
    with pytest.raises(CavitError):
        model(img, torch.zeros(1, dtype=torch.long))


def test_cosine_annealing_matches_torch_scheduler():
    """cavit.optim.CosineAnnealing (closed form) against torch's CosineAnnealingLR stepped per epoch, the schedule of
    configure_optimizers (/root/reference/model_cross.py:280-291)."""
    from types import SimpleNamespace
    from cavit.optim import CosineAnnealing
    opt = SimpleNamespace(base_lr=1e-4, lr=1e-4)
    sched = CosineAnnealing(opt, T_max=250, eta_min=1e-6)
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.Adam([p], lr=1e-4)
    tsched = torch.optim.lr_scheduler.CosineAnnealingLR(topt, T_max=250, eta_min=1e-6)
    for _ in range(300):       # past T_max too: torch's recursive form keeps following the cosine
        topt.step()
        tsched.step()
        sched.step()
        assert abs(opt.lr - tsched.get_last_lr()[0]) < 1e-12
