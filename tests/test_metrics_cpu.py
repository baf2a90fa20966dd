"""CPU: the metric restatement (oracle/metrics.py; /root/reference/utils.py:18-62, model_cross.py:243-255) against
scikit-learn, and the checkpoint helpers."""
import numpy as np
import pytest
import torch

from oracle import metrics as OM


def _batch(seed, B, saturate=False, one_class=None):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, 2, generator=g) * (12.0 if saturate else 1.5)
    if saturate:
        logits[::3] = logits[::3].round()        # repeated scores: ties in the ROC curve
    labels = torch.randint(0, 2, (B,), generator=g)
    if one_class is not None:
        labels[:] = one_class
    return logits, labels


@pytest.mark.parametrize("seed,B,saturate", [(0, 8, False), (1, 64, False), (2, 256, True), (3, 37, True)])
def test_against_scikit_learn(seed, B, saturate):
    from sklearn import metrics as SK
    logits, labels = _batch(seed, B, saturate)
    m = OM.batch_metrics(logits, labels)
    pred = logits.argmax(1).numpy()
    y = labels.numpy()
    prob = torch.softmax(logits, 1)[:, 1].numpy()
    tn, fp, fn, tp = SK.confusion_matrix(y, pred, labels=[0, 1]).ravel()
    ref = {"acc": SK.accuracy_score(y, pred), "prec": SK.precision_score(y, pred, zero_division=0),
           "rec": SK.recall_score(y, pred, zero_division=0), "spec": tn / (tn + fp) if tn + fp else 0.0,
           "f1": SK.f1_score(y, pred, zero_division=0), "npv": tn / (tn + fn) if tn + fn else 0.0,
           "auc_roc": SK.roc_auc_score(y, prob)}
    for k, v in ref.items():
        assert abs(m[k] - v) < 1e-6, (k, m[k], v)


def test_degenerate_batches_follow_the_zero_conventions():
    logits, labels = _batch(5, 16, one_class=1)
    m = OM.batch_metrics(logits, labels)
    assert m["auc_roc"] == 0.0 and m["spec"] == 0.0 and m["npv"] == 0.0       # no negatives at all
    logits[:, 0] = 10.0                                                        # never predicts the positive class
    m = OM.batch_metrics(logits, labels)
    assert m["prec"] == 0.0 and m["rec"] == 0.0 and m["f1"] == 0.0 and m["acc"] == 0.0
    z = torch.zeros(4, 2)                                                      # equal logits: argmax keeps class 0
    m = OM.batch_metrics(z, torch.tensor([0, 1, 0, 1]))
    assert m["acc"] == 0.5 and m["auc_roc"] == 0.5 and m["spec"] == 1.0 and m["rec"] == 0.0


def test_epoch_reduction_is_the_weighted_mean_of_batch_values():
    b1, b2 = _batch(7, 8), _batch(8, 24)
    e = OM.epoch_metrics([(b1[0], b1[1], 0.7), (b2[0], b2[1], 0.3)], prefix="val")
    m1, m2 = OM.batch_metrics(*b1), OM.batch_metrics(*b2)
    assert abs(e["val_acc"] - (8 * m1["acc"] + 24 * m2["acc"]) / 32) < 1e-12
    assert abs(e["val_loss"] - (8 * 0.7 + 24 * 0.3) / 32) < 1e-12
    assert set(e) == {f"val_{n}" for n in OM.NAMES}


def test_checkpoint_round_trip_and_reference_ckpt(tmp_path):
    from oracle.cases import CASES
    from oracle.functional import make_config
    from cavit.checkpoint import load_reference_checkpoint, save_reference_checkpoint
    from cavit.modules import ModelCross
    cfg = make_config(**CASES["cross_chain3"][1])
    torch.manual_seed(1)
    a = ModelCross(cfg)
    torch.manual_seed(2)
    b = ModelCross(cfg)
    p = str(tmp_path / "epoch=3.ckpt")
    save_reference_checkpoint(a, p, epoch=3, global_step=120)
    rest = load_reference_checkpoint(b, p)
    assert rest["epoch"] == 3 and rest["global_step"] == 120 and "pytorch-lightning_version" in rest
    # weights_only loading: a checkpoint that pickles arbitrary objects is refused unless the caller vouches for it
    import types
    from cavit import CavitError
    bad = str(tmp_path / "pickled.ckpt")
    torch.save({"state_dict": a.state_dict(), "hyper_parameters": types.SimpleNamespace(lr=1e-4)}, bad)
    with pytest.raises(CavitError):
        load_reference_checkpoint(b, bad)
    load_reference_checkpoint(b, bad, trusted=True)
    for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(x, y), k
    # a checkpoint of the REAL reference model (container only): Lightning's layout is {"state_dict": module.state_dict(), ...}
    from oracle import ref_loader
    if ref_loader.reference_dir() is None:
        return
    mod = ref_loader.load("model_cross")
    torch.manual_seed(9)
    ref = mod.ModelCross(ref_loader.to_config_dict(cfg))
    q = str(tmp_path / "ref.ckpt")
    torch.save({"state_dict": ref.state_dict(), "epoch": 0, "pytorch-lightning_version": "2.x"}, q)
    load_reference_checkpoint(b, q)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k
    save_reference_checkpoint(b, q)
    ref.load_state_dict(torch.load(q, weights_only=False)["state_dict"])     # and back into the reference module


def test_metrics_have_no_cpu_path():
    from cavit import CavitError
    from cavit.metrics import EpochMetrics
    with pytest.raises(CavitError):
        EpochMetrics(torch.device("cpu"))
