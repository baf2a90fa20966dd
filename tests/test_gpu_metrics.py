"""GPU: cavit_batch_metrics through the C ABI against the restatement of the reference's per-batch metrics
(oracle/metrics.py; /root/reference/utils.py:18-62, model_cross.py:243-255). Counts are exact; the seven values are fp32
quotients of them (tolerance 1e-6 covers the trapezoid-vs-pair-count rounding of AUROC)."""
import pytest
import torch

from oracle import metrics as OM

pytestmark = pytest.mark.gpu
TOL = 1e-6


def _batch(seed, B, scale=1.5, ties=False, one_class=None):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, 2, generator=g) * scale
    if ties:
        logits[::3] = logits[::3].round()
    labels = torch.randint(0, 2, (B,), generator=g)
    if one_class is not None:
        labels[:] = one_class
    return logits, labels


@pytest.mark.parametrize("B,scale,ties", [(1, 1.5, False), (2, 1.5, False), (8, 1.5, False), (256, 1.5, False),
                                          (257, 12.0, True), (1000, 30.0, True), (8192, 2.0, True)])
def test_single_batch_matches_oracle(B, scale, ties):
    from cavit.metrics import EpochMetrics
    logits, labels = _batch(B, B, scale, ties)
    em = EpochMetrics("cuda:0", prefix="val")
    em.update(logits.cuda(), labels.cuda(), loss=torch.tensor(0.625).cuda())
    got = em.compute()
    want = OM.epoch_metrics([(logits, labels, 0.625)], prefix="val")
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) < TOL, (k, got[k], want[k])
    assert em.accum[8].item() == B and em.accum[9].item() == 1
    assert em.accum[10:].abs().sum().item() == 0            # scratch words are left at zero


def test_degenerate_batches():
    from cavit.metrics import EpochMetrics
    for kw in [dict(one_class=1), dict(one_class=0)]:
        logits, labels = _batch(3, 33, **kw)
        em = EpochMetrics("cuda:0")
        em.update(logits.cuda(), labels.cuda())
        got, want = em.compute(), OM.epoch_metrics([(logits, labels, 0.0)])
        for k in want:
            assert abs(got[k] - want[k]) < TOL, (k, got[k], want[k])
        assert got["train_auc_roc"] == 0.0
    z = torch.zeros(4, 2)
    em = EpochMetrics("cuda:0")
    em.update(z.cuda(), torch.tensor([0, 1, 0, 1]).cuda())
    got = em.compute()
    assert got["train_acc"] == 0.5 and got["train_auc_roc"] == 0.5 and got["train_rec"] == 0.0


def test_epoch_accumulation_and_reset():
    from cavit.metrics import EpochMetrics
    em = EpochMetrics("cuda:0")
    batches = []
    for i, B in enumerate([8, 8, 8, 5]):                     # ragged last batch, as a DataLoader leaves it
        logits, labels = _batch(10 + i, B, ties=(i % 2 == 0))
        loss = 0.7 - 0.1 * i
        batches.append((logits, labels, loss))
        em.update(logits.cuda(), labels.cuda(), loss=torch.tensor(loss, device="cuda"))
    got, want = em.compute(), OM.epoch_metrics(batches)
    for k in want:
        assert abs(got[k] - want[k]) < TOL, (k, got[k], want[k])
    em.reset()
    assert em.accum.abs().sum().item() == 0
    assert all(v == 0.0 for v in em.compute().values())


def test_update_inside_training_step_needs_no_sync():
    """Metrics of a real step: logits / loss straight from the model, labels on the device."""
    from oracle.cases import CASES
    from oracle.functional import make_config
    from cavit.metrics import EpochMetrics
    from cavit.modules import ModelCross
    cfg = make_config(**CASES["cross_chain3"][1])
    torch.manual_seed(0)
    model = ModelCross(cfg).cuda().train()
    B = 6
    img = torch.randn(B, cfg.num_modalities, 1, *cfg.img_size, device="cuda")
    labels = torch.randint(0, cfg.num_classes, (B,), device="cuda")
    em = EpochMetrics("cuda:0")
    logits, loss = model(img, labels)
    em.update(logits, labels, loss)
    loss.backward()
    got = em.compute()
    want = OM.epoch_metrics([(logits.detach().cpu(), labels.cpu(), float(loss.detach()))])
    for k in want:
        assert abs(got[k] - want[k]) < TOL, (k, got[k], want[k])


def test_bad_arguments_fail_loudly():
    from cavit import CavitError
    from cavit.metrics import EpochMetrics
    em = EpochMetrics("cuda:0")
    with pytest.raises(CavitError):
        em.update(torch.zeros(4, 3, device="cuda"), torch.zeros(4, dtype=torch.long, device="cuda"))   # 3 classes
    with pytest.raises(CavitError):
        em.update(torch.zeros(4, 2, device="cuda"), torch.zeros(5, dtype=torch.long, device="cuda"))
    with pytest.raises(CavitError):
        em.update(torch.zeros(9000, 2, device="cuda"), torch.zeros(9000, dtype=torch.long, device="cuda"))


def test_test_outputs_gather_once():
    """test_step / on_test_epoch_end (model_cross.py:294-308): logits of every batch, in order, on the host at the end —
    also when the model reuses its output buffer between steps."""
    from oracle.cases import build_case
    from cavit.metrics import TestOutputs
    from cavit.modules import ModelCross
    kind, cfg, state, img, labels = build_case("cross_chain3")
    model = ModelCross(cfg)
    model.load_state_dict(state)
    model = model.cuda().eval()
    outs, want = TestOutputs(), []
    with torch.no_grad():
        for step in range(3):
            x = (img + 0.25 * step).cuda()
            logits, _ = model(x, labels.cuda())
            outs.append(logits, labels.cuda())
            want.append(logits.cpu().clone())
    got_logits, got_targets = outs.finish()
    assert torch.equal(got_logits, torch.cat(want)) and torch.equal(got_targets, labels.repeat(3))
    assert not torch.equal(want[0], want[1])
