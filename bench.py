#!/usr/bin/env python
"""bench.py — MRI volumes/sec, forward+backward, of the cross-attention ViT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|...]

ours       the sm_100a kernel path (cavit) on N B200s, one process per GPU (torchrun for N > 1),
           batch-sharded data parallel with NCCL gradient all-reduce (weak scaling: per-GPU batch fixed).
reference  the reference's own algorithm on the box's host cores (the CPU oracle port of
           /root/reference/model_cross.py — /root/reference does not exist on the GPU box), same model config,
           bounded batch, all host threads. Rank 0 only.

One JSON line on stdout (rank 0). See DESIGN.md §Measurement for how each field is produced.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))

import torch  # noqa: E402

RING4 = {"0": "1", "1": "2", "2": "3", "3": "0"}
WORKLOADS = {
    # BASELINE.json configs[1]: 2D slice cross-attn ViT, 4 MRI sequences, 224x224, patch 16, dim 384, 6 layers, batch 256
    "cfg2": dict(cfg=dict(hidden_dim=384, mlp_dim=1536, num_heads=6, num_multi_blocks=3, num_self_blocks=2,
                          patch_size=(16, 16, 1), img_size=(224, 224, 1), num_modalities=4, attn_order=RING4,
                          num_classes=2, dropout=0.0, label_smoothing=0.0), batch=256, cpu_batch=2),
    # BASELINE.json configs[0] shape (config2.py defaults), synthetic volumes
    "cfg1": dict(cfg=dict(hidden_dim=1024, mlp_dim=4096, num_heads=16, num_multi_blocks=2, num_self_blocks=2,
                          patch_size=(16, 16, 8), img_size=(128, 128, 64), num_modalities=4, attn_order=RING4,
                          num_classes=2, dropout=0.0, label_smoothing=0.0), batch=32, cpu_batch=2),
    # BASELINE.json configs[2] per-GPU shard: 3D volumes 240x240x160 (155 padded), 16^3 patches, dim 768, 12 layers, 4 / GPU
    "cfg3": dict(cfg=dict(hidden_dim=768, mlp_dim=3072, num_heads=12, num_multi_blocks=6, num_self_blocks=2,
                          patch_size=(16, 16, 16), img_size=(240, 240, 160), num_modalities=4, attn_order=RING4,
                          num_classes=2, dropout=0.0, label_smoothing=0.0), batch=4, cpu_batch=1),
    # BASELINE.json configs[4]: long sequences, 8^3 patches, dim 512
    "cfg5": dict(cfg=dict(hidden_dim=512, mlp_dim=2048, num_heads=8, num_multi_blocks=2, num_self_blocks=2,
                          patch_size=(8, 8, 8), img_size=(128, 128, 128), num_modalities=4, attn_order=RING4,
                          num_classes=2, dropout=0.0, label_smoothing=0.0), batch=8, cpu_batch=1),
    "tiny": dict(cfg=dict(hidden_dim=128, mlp_dim=256, num_heads=2, num_multi_blocks=1, num_self_blocks=1,
                          patch_size=(16, 16, 1), img_size=(64, 64, 1), num_modalities=4, attn_order=RING4,
                          num_classes=2, dropout=0.0, label_smoothing=0.0), batch=8, cpu_batch=2),
}


def flops_per_volume(cfg) -> float:
    """Algorithmic forward FLOPs per volume; fwd+bwd = 3x."""
    C, F, M = cfg.hidden_dim, cfg.mlp_dim, cfg.num_modalities
    D, H, W = cfg.img_size
    dp, hp, wp = cfg.patch_size
    Np = (D // dp) * (H // hp) * (W // wp)
    P = dp * hp * wp
    N = Np + 1
    L = cfg.num_multi_blocks * cfg.num_self_blocks
    K = len(cfg.attn_order)
    embed = M * 2 * Np * P * C
    self_proj = M * L * (2 * N * C * 3 * C + 2 * N * C * C + 4 * N * C * F)
    self_attn = M * L * 4 * N * N * C
    cross_proj = cfg.num_multi_blocks * K * (4 * N * C * C + 4 * C * C + 4 * C * F)
    cross_attn = cfg.num_multi_blocks * K * 4 * N * C
    head = M * (2 * C * F + 2 * F * cfg.num_classes)
    return float(embed + self_proj + self_attn + cross_proj + cross_attn + head)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16_tflops": d.get("bf16_tflops", 1590.0),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_port_throughput(wl, steps: int, warmup: int):
    """The reference's algorithm (oracle port) fwd+bwd on the host cores, bounded batch. Only used when neither
    /root/reference nor oracle/_ref (oracle/make_ref.py) is present."""
    from oracle import functional as OF
    from oracle.weights import make_inputs, make_state, state_schema_cross
    cfg = OF.make_config(**wl["cfg"])   # oracle-side config: this leg IS the CPU checker being timed
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = make_state(state_schema_cross(cfg), seed=0, init="reference")
    B = wl["cpu_batch"]
    img, labels = make_inputs(cfg, B, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        lp = OF.leaf_params(state, torch.float32)
        logits, loss = OF.model_cross_forward(lp, img, labels, cfg)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return {"value": B / med, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} fwd+bwd steps of the same model at batch {B} (fp32, oracle port of model_cross.py), median",
            "s_per_step": med}


def _reference_model(cfg_kw):
    """The UNMODIFIED reference ModelCross (imported from /root/reference, or from its byte-for-byte copy in the
    git-ignored oracle/_ref on the GPU box), weights from its own initialiser under torch.manual_seed(0)."""
    from oracle import functional as OF, ref_loader
    if ref_loader.reference_dir() is None:
        return None, None, None
    cfg = OF.make_config(**cfg_kw)
    mod = ref_loader.load("model_cross")
    torch.manual_seed(0)
    model = mod.ModelCross(ref_loader.to_config_dict(cfg)).train()   # dropout 0.0: identity in train mode
    return model, cfg, ref_loader.reference_dir()


def cpu_reference_throughput(wl, steps: int, warmup: int):
    """The reference's own CPU implementation of the path (model_cross.py:186-212 forward + loss.backward()), fp32 as
    shipped, all host threads, on a bounded batch of the arm's workload."""
    from oracle.weights import make_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, cfg, where = _reference_model(wl["cfg"])
    if model is None:
        return cpu_port_throughput(wl, steps, warmup)
    B = wl["cpu_batch"]
    img, labels = make_inputs(cfg, B, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        logits, loss = model(img, labels)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return {"value": B / med, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"{len(times)} fwd+bwd steps of the unmodified reference ModelCross ({where}) on the same model "
                      f"config at batch {B} (fp32 as shipped, synthetic N(0,1) volumes), median",
            "s_per_step": med}


def configs0_real_voxels(dev):
    """BASELINE.json configs[0] / BASELINE.md section 5: the reference ModelCross (config2.py defaults, M = 4 ring) on the
    six bundled UCSF-PDGM cases (T1, T1c, T2, FLAIR; centre window 128 x 128 x 64; raw intensities), 3 batches of 2,
    on the host cores — and the same batches through the sm_100a path on `dev`, logits compared in the same run.
    Needs oracle/_ref (oracle/make_ref.py); returns None without it."""
    import numpy as np
    from oracle import make_ref
    path = os.path.join(ROOT, "oracle", "_ref", "ucsf_cfg1_int16.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    wl = WORKLOADS["cfg1"]
    model, cfg, where = _reference_model(wl["cfg"])
    if model is None:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    img = make_ref.volumes_fp32(z["stored"], z["slope"], z["inter"], cfg.img_size)      # [6, 4, 1, 128, 128, 64] fp32
    labels = torch.from_numpy(z["labels"])
    state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    times, ref_logits = [], []
    for i in (0, 0, 1, 2):          # one warm-up pass over the first batch, then the three batches
        x, y = img[2 * i:2 * i + 2], labels[2 * i:2 * i + 2]
        t0 = time.perf_counter()
        model.zero_grad(set_to_none=True)
        logits, loss = model(x, y)
        loss.backward()
        times.append(time.perf_counter() - t0)
        ref_logits.append(logits.detach())
    times, ref_logits = sorted(times[1:]), torch.cat(ref_logits[1:])
    out = {"cpu": {"value": 2 / times[1], "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "reference",
                   "sample": f"unmodified reference ModelCross ({where}), config2.py defaults, fp32, 3 batches of 2 real "
                             "UCSF-PDGM cases (T1, T1c, T2, FLAIR), median step", "s_per_step": times[1]}}
    if dev is not None:
        from cavit.config import make_config
        from cavit.modules import ModelCross
        for prec in ("bf16", "fp32"):
            ours = ModelCross(make_config(**wl["cfg"]))
            ours.load_state_dict(state)
            ours.set_precision(prec)
            ours = ours.to(dev).train()
            got, ms = [], []
            for rep in range(3):    # repetitions 0, 1 warm up (eager runs + graph capture); the third is timed
                got = []
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(3):
                    logits, loss = ours(img[2 * i:2 * i + 2].to(dev), labels[2 * i:2 * i + 2].to(dev))
                    loss.backward()
                    for p in ours.parameters():
                        p.grad = None
                    got.append(logits.detach().float().cpu())
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1) / 3)
            got = torch.cat(got)
            out[prec] = {"value": 2 / (ms[-1] * 1e-3), "unit": "volumes/s", "ms_per_step": ms[-1],
                         "logits_rel_vs_reference_fp32": float((got.double() - ref_logits.double()).norm() / ref_logits.double().norm()),
                         "note": "same six cases, H2D copies inside the timed steps"}
            del ours
    return out


def run_reference(args, wl, rank):
    if rank != 0:
        return
    res = cpu_reference_throughput(wl, max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": "MRI volumes/sec fwd+bwd", "value": res["value"], "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["s_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, wl, args.batch or wl["batch"], max(1, args.gpus), None, args.precision),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, wl, B, world, mode, precision="bf16"):
    """The `config` object of the JSON line: the same for both arms (the reference arm runs a bounded sample of it,
    described in its cpu_baseline.sample)."""
    c = wl["cfg"]
    prec = {"bf16": "bf16 GEMM/attention operands, fp32 accumulate, fp32 residual stream / LayerNorm / softmax statistics",
            "fp32": "fp32-tolerance mode: bf16 hi+lo split operands (3 MMAs per product), fp32 accumulate, fp32 SIMT attention"}
    return {"workload": f"{name}: ModelCross C={c['hidden_dim']} H={c['num_heads']} F={c['mlp_dim']} "
                        f"{c['num_multi_blocks']}x{c['num_self_blocks']} blocks, img {tuple(c['img_size'])} patch {tuple(c['patch_size'])}, "
                        f"M=4 ring cross-attention, per-GPU batch {B}",
            "global_batch": B * world, "parallelism": f"dp{world}" + (f" ({mode} all-reduce)" if world > 1 and mode else ""),
            "l2": "per-step working set (activations + weights, several GB) far exceeds the 126 MB L2; no explicit flush",
            "precision": prec[precision]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16: bf16 operands (headline); fp32: the fp32-tolerance mode (split operands)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-tolerance-mode leg")
    ap.add_argument("--no-configs0", action="store_true", help="skip the configs[0] real-voxel leg (CPU reference + GPU)")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if args.warmup < 5:
        args.warmup = 5   # 2 eager runs + graph capture + first (slow) replay happen inside the warm-up

    from cavit import _abi, ops
    from cavit.modules import ModelCross
    from cavit.config import make_config

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _abi.require_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL's log (version banner, ring / NVLS setup, rank count) goes to a per-rank file instead of stdout, where it
        # would sit next to the JSON line: gpurun_out/nccl_<host>_<pid>.log when NCCL_DEBUG is set by the caller
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            os.environ["NCCL_DEBUG_FILE"] = os.path.join(ROOT, "gpurun_out", "nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)
    cfg = make_config(**wl["cfg"])
    B = args.batch or wl["batch"]
    torch.manual_seed(0)
    model = ModelCross(cfg)
    model.set_precision(args.precision)
    model = model.cuda().train()
    runner = model
    if world > 1:
        from cavit.ddp import DataParallel
        runner = DataParallel(model, mode=os.environ.get("CAVIT_DDP_MODE", "auto"))
    g = torch.Generator().manual_seed(1234 + rank)
    D, H, W = cfg.img_size
    img_host = torch.randn((B, cfg.num_modalities, 1, D, H, W), generator=g).pin_memory()
    labels_host = torch.randint(0, cfg.num_classes, (B,), generator=g).pin_memory()
    img = img_host.to(dev, non_blocking=True)
    labels = labels_host.to(dev, non_blocking=True)

    def step_resident():
        logits, loss = runner(img, labels)
        loss.backward()
        for p in model.parameters():
            p.grad = None
        return loss

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    def n_launches():
        return _abi.launch_count() + model.engine().graph_launches
    launches0 = n_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step_resident()
    e1.record()
    barrier()
    launches = n_launches() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST buffers (pinned): every step's batch is copied host -> device
    # and every step's loss is read back to the host inside the timed region. cavit.data.DevicePrefetcher copies
    # batch i+1 on a copy stream while step i runs; cavit.data.ScalarReadback delivers the loss one step late.
    from cavit.data import DevicePrefetcher, ScalarReadback

    logits_host = torch.empty((4, B, cfg.num_classes), dtype=torch.float32).pin_memory()

    def run_e2e(nsteps):
        feed = DevicePrefetcher(((img_host, labels_host) for _ in range(nsteps)), dev)
        rb = ScalarReadback(dev)
        losses, d2h_logits = [], 0
        for i, (x, y) in enumerate(feed):
            logits, loss = runner(x, y)
            loss.backward()
            for p in model.parameters():
                p.grad = None
            # both results of forward() go back to the host every step: the [B, classes] logits into a pinned ring (their
            # copy is ordered before the loss copy on the same stream, so popping step i's loss implies step i's logits)
            logits_host[i % 4].copy_(logits.detach(), non_blocking=True)
            d2h_logits += logits.numel() * 4
            rb.push(loss)
            if rb.pending() > 1:
                losses.append(rb.pop())
        while rb.pending():
            losses.append(rb.pop())
        assert len(losses) == nsteps
        return feed.h2d_bytes // nsteps, (rb.d2h_bytes + d2h_logits) // nsteps, losses

    run_e2e(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    h2d_bytes, d2h_bytes, e2e_losses = run_e2e(args.steps)
    e3.record()
    barrier()
    e2e_ms = max_over_ranks(e2.elapsed_time(e3))
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): one extra instrumented step
    roofline, breakdown = None, None
    if not args.no_profile:
        ops.PROFILE = []
        step_resident()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        peaks = measured_peaks()
        gemm_bound_ms, gemm_bytes = 0.0, 0.0
        for name, a, b, info in prof:
            t = a.elapsed_time(b)
            d = agg.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0})
            d["ms"] += t
            d["n"] += 1
            d["flops"] += info.get("flops", 0.0)
            if name == "gemm":   # per launch: the slower of its tensor-pipe time and its HBM time at the measured peaks
                gemm_bytes += info.get("bytes", 0.0)
                gemm_bound_ms += max(info["flops"] / (peaks["bf16_tflops_sustained"] * 1e12),
                                     info.get("bytes", 0.0) / (peaks["hbm_gbs"] * 1e9)) * 1e3
        tot = sum(d["ms"] for d in agg.values())
        gm = agg.get("gemm", {"ms": 0.0, "n": 0, "flops": 0.0})
        achieved = gm["flops"] / (gm["ms"] * 1e-3) / 1e12 if gm["ms"] > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        # `traffic` cannot be measured outside a profiler: it is a CITATION of the committed ncu launch list of this
        # workload (mean dram__bytes_read + dram__bytes_write per GEMM launch of one step), newest round first
        traffic, traffic_file = None, None
        if args.workload == "cfg2" and B == wl["batch"] and args.precision == "bf16":
            for name in ("step_summary_r02.json", "step_summary_r01.json"):
                try:
                    with open(os.path.join(ROOT, "profiles", name)) as f:
                        traffic = json.load(f)["kernels"]["gemm_bf16_tcgen05_kernel"]["traffic_bytes_per_launch"]
                    traffic_file = "profiles/" + name
                    break
                except Exception:
                    continue
        roofline = {"kernel": "gemm_bf16_tcgen05_kernel", "bound": "tensor", "achieved": achieved, "peak": peak,
                    "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_cited_from_committed_ncu_file": (f"{traffic_file}: ncu dram__bytes_read + dram__bytes_write per GEMM "
                                                              "launch, mean over one step; not measured in this run") if traffic else None,
                    "peak_source": peaks["source"] + " (sustained)",
                    "launches_per_step": gm["n"], "share_of_step": gm["ms"] / tot if tot else None,
                    # K = 384 GEMMs with bf16 I/O sit at ~170 FLOP/B against a machine balance of ~210: several of them are
                    # bounded by HBM, not by the tensor pipe. Both limits per launch:
                    "two_limit": {"algorithmic_bytes_per_step": gemm_bytes,
                                  "hbm_gbs_achieved": gemm_bytes / (gm["ms"] * 1e-3) / 1e9 if gm["ms"] > 0 else None,
                                  "bound_ms_per_step": gemm_bound_ms,
                                  "frac_of_bound": gemm_bound_ms / gm["ms"] if gm["ms"] > 0 else None,
                                  "how": "sum over launches of max(flops / measured sustained bf16 peak, algorithmic bytes / "
                                         "measured HBM peak) divided by the summed measured durations"}}
        breakdown = {k: {"ms": round(v["ms"], 3), "n": v["n"],
                         **({"tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)} if v["flops"] and v["ms"] > 0 else {})}
                     for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}

    # ---- optimizer step, reported separately: the metric is fwd+bwd, the reference's Adam is not in it
    optimizer = None
    if not args.no_profile and world == 1:
        from cavit.optim import FusedAdam
        opt = FusedAdam(model, lr=1e-4, weight_decay=5e-4)
        step_resident()
        for _ in range(2):
            opt.step()
        torch.cuda.synchronize()
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(5):
            opt.step()
        o1.record()
        torch.cuda.synchronize()
        adam_ms = o0.elapsed_time(o1) / 5
        nparam = model.engine().flat.numel()
        for _ in range(4):          # graphs are re-captured without the per-step cast kernel
            step_resident()
            opt.step()
        torch.cuda.synchronize()
        o0.record()
        for _ in range(args.steps):
            step_resident()
            opt.step()
        o1.record()
        torch.cuda.synchronize()
        peaks_ = measured_peaks()
        optimizer = {"kernel": "adam_step_kernel", "ms": adam_ms, "params": nparam, "bytes_per_param": 30,
                     "achieved_gbs": nparam * 30 / (adam_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks_["hbm_gbs"],
                     "frac_of_hbm": nparam * 30 / (adam_ms * 1e-3) / 1e9 / peaks_["hbm_gbs"],
                     "train_step_ms_with_optimizer": o0.elapsed_time(o1) / args.steps,
                     "note": "one launch over the flat fp32 slabs (p, g, m, v) that also rewrites the bf16 operand copy"}

    # ---- the fp32-tolerance mode (north_star's ~1e-3 bar) on the same workload, reported next to the bf16 headline
    fp32_mode = None
    if world == 1 and args.precision == "bf16" and not args.no_fp32 and not args.no_profile:
        m32 = ModelCross(cfg)
        m32.load_state_dict(model.state_dict())
        m32.set_precision("fp32")
        m32 = m32.cuda().train()

        def step32():
            logits32, loss32 = m32(img, labels)
            loss32.backward()
            for p in m32.parameters():
                p.grad = None
            return logits32, loss32

        for _ in range(4):     # eager, eager, graph capture, first replay
            logits32, loss32 = step32()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nst = max(2, min(args.steps, 5))
        f0.record()
        for _ in range(nst):
            logits32, loss32 = step32()
        f1.record()
        torch.cuda.synchronize()
        ms32 = f0.elapsed_time(f1) / nst
        with torch.no_grad():
            lb, _ = model(img, labels)
        fp32_mode = {"value": B / (ms32 * 1e-3), "unit": "volumes/s", "ms_per_step": ms32,
                     "what": "same model / batch with precision='fp32': bf16 hi+lo split operands (3 tcgen05 MMAs per product), "
                             "fp32 CUDA-core attention, fp32 activations between kernels; parity ~1e-5..1e-4 on logits vs the fp64 "
                             "reference at the BASELINE shapes (tests/test_gpu_baseline_shapes.py)",
                     "bf16_vs_fp32_mode_logits_rel": float((lb.double() - logits32.double()).norm() / logits32.double().norm()),
                     "loss": float(loss32)}
        del m32, logits32, loss32
        torch.cuda.empty_cache()

    if rank == 0:
        cpu, configs0 = None, None
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_throughput(wl, 3, 1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            if not args.no_configs0:
                del model, runner, img   # free the bench model before the cfg1 models of this leg are built
                torch.cuda.empty_cache()
                try:
                    configs0 = configs0_real_voxels(dev)
                except Exception as exc:   # a reported extra: never lose the bench line over it
                    configs0 = {"error": repr(exc)}
        fpv = 3.0 * flops_per_volume(cfg)
        peaks = measured_peaks()
        line = {
            "metric": "MRI volumes/sec fwd+bwd", "value": value, "unit": "volumes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (hi+lo split operands, fp32 accumulate)",
            "data": "synthetic",
            "config": workload_config(args.workload, wl, B, world, None, args.precision),
            "ddp_mode": runner.mode if world > 1 else None,
            "model_tflops": value * fpv / 1e12,
            "model_flops_frac_of_peak": (value / world) * fpv / 1e12 / peaks["bf16_tflops_sustained"],
            "e2e": {"value": e2e_value, "unit": "volumes/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps,
                    "how": "model(img, labels) + loss.backward() per step; pinned host batch -> device by cavit.data.DevicePrefetcher "
                           "(copy of step i+1 overlaps step i), logits and loss -> pinned host memory every step (read one step late)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernel_breakdown_ms": breakdown,
            "optimizer": optimizer,
            "fp32_mode": fp32_mode,
            "configs0_real_voxels": configs0,
            "loss": float(loss.detach()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        # NCCL keeps a communicator alive while a captured CUDA graph references it (the overlapped slab all-reduces are
        # recorded into the backward graph): drop the graphs first, and never let teardown outlive the measurement
        def _force_exit():
            sys.stdout.flush()
            os._exit(0)
        t = threading.Timer(30.0, _force_exit)
        t.daemon = True
        t.start()
        if hasattr(runner, "close"):
            runner.close()
        dist.barrier()
        dist.destroy_process_group()
        t.cancel()


if __name__ == "__main__":
    main()
