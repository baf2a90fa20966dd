"""CPU, world_size 2, gloo: the slab bookkeeping of cavit.ddp (coalescing of backward-ordered flat
ranges, in-place averaging) — the host-side logic of the multi-GPU path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cavit.ddp import SlabReducer
    from cavit.engine import build_layout
    from oracle.cases import CASES
    from oracle.functional import make_config
    cfg = make_config(**CASES["cross_ring4"][1])
    lay = build_layout("cross", cfg)
    flat = torch.full((lay.total,), float(rank + 1))
    red = SlabReducer(min_elems=lay.total // 4)
    order = [r for r in lay.layer_ranges if r[0] == "head"]
    order += sorted([r for r in lay.layer_ranges if r[0] not in ("head", "embed")], key=lambda r: -r[1])
    order += [r for r in lay.layer_ranges if r[0] == "embed"]
    n_slabs = 0
    for tag, s, e in order:
        ready = red.add(s, e)
        if tag == "embed":
            ready += red.flush()
        for a, b in ready:
            red.reduce_(flat, a, b)
            n_slabs += 1
    covered = sorted(red.issued)
    ok = covered[0][0] == 0 and covered[-1][1] == lay.total and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    q.put((rank, bool(ok), n_slabs, float(flat.min()), float(flat.max())))
    dist.destroy_process_group()


def test_slab_reducer_covers_flat_buffer_and_averages():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, n_slabs, lo, hi in res:
        assert ok
        assert 2 <= n_slabs <= 6
        assert lo == hi == 1.5      # mean of 1 and 2 everywhere: every element reduced exactly once


def _stem_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cavit.ddp import TensorGradReducer
    from cavit.encoders import CNN3DEncoder
    torch.manual_seed(100 + rank)                       # different replicas before the broadcast
    stem = CNN3DEncoder(hidden_dim=16)
    red = TensorGradReducer(list(stem.parameters()))
    red.broadcast(list(stem.parameters()) + list(stem.buffers()))
    w0 = float(stem.conv1.weight.double().sum())
    x = torch.randn(2, 1, 32, 32, 32, generator=torch.Generator().manual_seed(7 + rank))  # different shards
    stem(x).square().mean().backward()
    g = stem.conv4.weight.grad.clone()
    gs = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(gs, g)
    same = all(torch.equal(gs[0], t) for t in gs)
    q.put((rank, w0, bool(same), red.reduced, len(red.params)))
    dist.destroy_process_group()


def test_stem_gradients_are_averaged_across_ranks():
    """The CNN stems of ViT / ViT3D are outside the flat buffer: their parameters are broadcast and their gradients
    all-reduced per tensor (cavit.ddp.TensorGradReducer) — world_size 2 over gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stem_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1]                        # same weights after the broadcast
    for rank, w0, same, reduced, n in res:
        assert same and reduced == n


def _metrics_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cavit.metrics import epoch_means
    # rank 0 saw 3 batches (40 samples), rank 1 one batch (8 samples): weighted sums as cavit_batch_metrics leaves them
    accum = torch.zeros(10, dtype=torch.float64)
    if rank == 0:
        accum[:8] = torch.arange(1, 9, dtype=torch.float64) * 40 * 0.1
        accum[8], accum[9] = 40, 3
    else:
        accum[:8] = torch.arange(1, 9, dtype=torch.float64) * 8 * 0.05
        accum[8], accum[9] = 8, 1
    q.put((rank, epoch_means(accum).tolist()))
    dist.destroy_process_group()


def test_epoch_metrics_average_rank_means():
    """Lightning's sync_dist reduction (/root/reference/model_cross.py:243-255): each rank's own weighted epoch mean,
    then the plain mean over ranks — NOT the sample-weighted global mean."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_metrics_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, vals in res:
        for k, v in enumerate(vals):
            assert abs(v - (k + 1) * (0.1 + 0.05) / 2) < 1e-12


def test_auto_mode_switches_on_gradient_volume():
    """mode="auto": one post-backward all-reduce for small models (cfg2: 66 M parameters), overlapped slabs from 1 GiB of fp32
    gradients on (cfg3: 525 M parameters) — measured choice, DESIGN.md section 5."""
    from cavit.ddp import DataParallel
    assert DataParallel.auto_mode(66_408_640) == "post"
    assert DataParallel.auto_mode(82_200_000) == "post"
    assert DataParallel.auto_mode(524_600_000) == "overlap"
    assert DataParallel.auto_mode((1 << 28)) == "overlap" and DataParallel.auto_mode((1 << 28) - 1) == "post"
