"""Host -> device staging helpers (cavit.data): ordering, slot reuse and value integrity under overlap."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_yields_every_batch_in_order_with_slot_reuse():
    from cavit.data import DevicePrefetcher
    n, shape = 7, (4, 2, 1, 8, 8, 8)
    host = [(torch.full(shape, float(i)).pin_memory(), torch.full((4,), i, dtype=torch.int64).pin_memory()) for i in range(n)]
    feed = DevicePrefetcher(iter(host), "cuda:0", depth=2)
    sums, ptrs = [], set()
    big = torch.randn(4096, 4096, device="cuda")
    for i, (x, y) in enumerate(feed):
        assert x.is_cuda and y.is_cuda and x.shape == shape
        for _ in range(3):                     # keep the compute stream busy while the next copy is in flight
            big = (big @ big).clamp_(-1, 1)
        sums.append((x.sum() + y.sum().float()))
        ptrs.add(x.data_ptr())
    torch.cuda.synchronize()
    numel = 1
    for s in shape:
        numel *= s
    assert [float(s) for s in sums] == [float(i * numel + 4 * i) for i in range(n)]
    assert len(ptrs) == 2                      # two device slots, reused
    assert feed.h2d_bytes == n * (numel * 4 + 4 * 8)


def test_scalar_readback_is_fifo_and_lagged():
    from cavit.data import ScalarReadback
    from cavit import CavitError
    rb = ScalarReadback("cuda:0", depth=3)
    buf = torch.zeros(1, device="cuda")
    got = []
    for i in range(10):
        buf.fill_(float(i))                    # the same device scalar is overwritten every step (like the loss buffer)
        rb.push(buf)
        if rb.pending() > 1:
            got.append(rb.pop())
    while rb.pending():
        got.append(rb.pop())
    assert got == [float(i) for i in range(10)]
    with pytest.raises(CavitError):
        rb.pop()


def test_training_through_prefetcher_matches_direct_call():
    from cavit.data import DevicePrefetcher
    from cavit.modules import ModelCross
    from oracle.cases import build_case
    kind, cfg, state, img, labels = build_case("cross_chain3")
    model = ModelCross(cfg)
    model.load_state_dict(state)
    model = model.cuda().train()
    logits0, loss0 = model(img.cuda(), labels.cuda())
    feed = DevicePrefetcher(((img.pin_memory(), labels.pin_memory()) for _ in range(4)), "cuda:0")
    for x, y in feed:                          # eager, eager, graph capture, graph replay
        logits, loss = model(x, y)
        loss.backward()
        assert torch.equal(logits, logits0) and torch.equal(loss, loss0)
