"""Worker of tests/test_gpu_ddp_nccl.py (one process per GPU, launched with torch.distributed.run): batch-sharded
data parallel over NCCL must reproduce the 1-GPU global-batch gradients (the reference gets this from
Lightning's DDP strategy, /root/reference/main_mist.py:211-219)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cross-attention-vit_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    mode = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cavit.ddp import DataParallel
    from cavit.modules import ModelCross
    from oracle.cases import build_case
    from oracle.weights import make_inputs
    kind, cfg, state, _, _ = build_case("cross_ring4")
    B = 4 * world
    img, labels = make_inputs(cfg, B, seed=99)

    def run(model, x, y, runner=None, steps=5):
        out = None
        for _ in range(steps):     # eager, eager, capture, replay, replay
            for p in model.parameters():
                p.grad = None
            logits, loss = (runner or model)(x, y)
            loss.backward()
            out = torch.cat([p.grad.detach().flatten() for p in model.parameters()]).clone(), logits.detach().clone()
        return out

    # the global batch on one GPU (every rank computes it for itself: no communication)
    ref_model = ModelCross(cfg)
    ref_model.load_state_dict(state)
    ref_model = ref_model.to(dev).train()
    g_ref, logits_ref = run(ref_model, img.to(dev), labels.to(dev))

    model = ModelCross(cfg)
    if rank == 0:
        model.load_state_dict(state)          # other ranks start from their own random init: broadcast must fix that
    model = model.to(dev).train()
    dp = DataParallel(model, mode=mode, min_slab_elems=1 << 16)
    per = B // world
    sl = slice(rank * per, (rank + 1) * per)
    g, logits = run(model, img[sl].to(dev), labels[sl].to(dev), runner=dp)
    torch.cuda.synchronize()
    eng = model.engine()
    rel = float((g.double() - g_ref.double()).norm() / g_ref.double().norm())
    rel_logits = float((logits.double() - logits_ref[sl].double()).norm() / logits_ref[sl].double().norm())
    captured = any(v.get("graph") is not None for v in eng._bwd_graphs.values())
    print(f"rank {rank} mode {dp.mode}: grad rel {rel:.3e} logits rel {rel_logits:.3e} bwd graph captured {captured} "
          f"capture_failed {eng._hook_capture_failed} {eng.hook_capture_error}", flush=True)
    ok = rel < 2e-3 and rel_logits < 1e-5 and captured
    # every rank must hold the same averaged gradient
    gathered = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gathered, g)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
    dp.close()          # the backward graph recorded NCCL collectives: it has to go before the communicator does
    del model, ref_model, dp
    dist.barrier()
    print(f"rank {rank} ok={ok}", flush=True)
    import threading
    t = threading.Timer(30.0, lambda: os._exit(0 if ok else 1))    # teardown must never outlive the check
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
