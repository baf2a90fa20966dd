"""Batch-sharded data parallelism for the cavit engine: one process per GPU, full replica per
rank, gradients averaged with NCCL all-reduce over NVLink 5 / NVSwitch.

The reference gets this implicitly from Lightning's DDP strategy (`L.Trainer(devices=4,
num_nodes=2)`, /root/reference/main_mist.py:211-219 -> torch DDP -> 25 MB buckets). Here the
engine already writes every gradient into ONE flat fp32 buffer whose layout follows backward
completion order (head, layers L-1..0 with their fusion blocks, embedding), so there is nothing to
flatten or copy: as soon as the kernels that produce a layer's slab are enqueued, an event is
recorded and the slab is all-reduced (AVG) in place on a dedicated communication stream, which
overlaps the remaining backward kernels. `finish()` makes the compute stream wait for the
communication stream before the optimizer reads the gradients.

Works with any torch.distributed backend (NCCL on GPUs; gloo for the CPU tests of the slab
bookkeeping via `SlabReducer`).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class SlabReducer:
    """Coalesces contiguous flat ranges into slabs of at least `min_elems` and all-reduces them."""

    def __init__(self, min_elems: int = 8 << 20, group=None, average: bool = True):
        self.min_elems, self.group, self.average = min_elems, group, average
        self.pending: Optional[Tuple[int, int]] = None
        self.issued: List[Tuple[int, int]] = []

    def add(self, start: int, end: int) -> List[Tuple[int, int]]:
        """Register that flat[start:end) is final. Returns slabs that became ready to reduce."""
        if self.pending is None:
            self.pending = (start, end)
        else:
            ps, pe = self.pending
            if end == ps:
                self.pending = (start, pe)
            elif start == pe:
                self.pending = (ps, end)
            else:  # non-adjacent: flush what we have; the new range may itself already be a full slab
                out = [self.pending]
                self.pending = (start, end)
                if end - start >= self.min_elems:
                    out.append(self.pending)
                    self.pending = None
                self.issued += out
                return out
        if self.pending[1] - self.pending[0] >= self.min_elems:
            out = [self.pending]
            self.pending = None
            self.issued += out
            return out
        return []

    def flush(self) -> List[Tuple[int, int]]:
        out = [self.pending] if self.pending is not None else []
        self.pending = None
        self.issued += out
        return out

    def reduce_(self, flat: torch.Tensor, start: int, end: int):
        world = dist.get_world_size(self.group)
        if flat.is_cuda and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat[start:end], op=dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(flat[start:end], op=dist.ReduceOp.SUM, group=self.group)
            if self.average:
                flat[start:end].div_(world)


class TensorGradReducer:
    """Averages the gradients of ordinary (non-flat) parameters across ranks as autograd produces them: the CNN stems
    of `cavit.encoders.ViT` / `ViT3D` live outside the engine's flat buffer (torch modules); their
    gradients appear after the engine's backward, while autograd walks the stem. One all-reduce per parameter from a
    post-accumulate hook (a handful of small tensors); backend-agnostic (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, params, group=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.reduced = 0

    def _hook(self, p: torch.Tensor):
        world = dist.get_world_size(self.group)
        if p.grad.is_cuda and dist.get_backend(self.group) == "nccl":
            dist.all_reduce(p.grad, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
            p.grad.div_(world)
        self.reduced += 1

    def broadcast(self, tensors, src: int = 0):
        for t in tensors:
            dist.broadcast(t.data if isinstance(t, torch.nn.Parameter) else t, src=src, group=self.group)

    def remove(self):
        for h in self.handles:
            h.remove()


class DataParallel:
    """Wraps a cavit model: `dp = DataParallel(model); logits, loss = dp(img, labels); loss.backward()`."""

    AUTO_OVERLAP_BYTES = 1 << 30     # mode="auto": overlapped slab all-reduces from this gradient volume on, else one post all-reduce

    @classmethod
    def auto_mode(cls, gradient_elems: int) -> str:
        """What mode="auto" selects for a model with `gradient_elems` fp32 gradient elements."""
        return "overlap" if gradient_elems * 4 >= cls.AUTO_OVERLAP_BYTES else "post"

    def __init__(self, model, group=None, min_slab_elems: int = 8 << 20, mode: str = "auto"):
        """mode = "overlap": every finished gradient slab (>= min_slab_elems, in backward-completion order) is
        all-reduced on a communication stream while the remaining backward kernels run. The slab all-reduces are
        recorded INTO the backward CUDA graph (NCCL collectives are capturable), so overlap costs no eager launches;
        if this stack cannot capture them the engine falls back to eager launches and, for small models, this wrapper
        to "post";
        mode = "post": backward replays its CUDA graph and the flat gradient buffer is all-reduced afterwards in
        one call; "auto" = "overlap" from 1 GiB of fp32 gradients on, "post" below (measured, see below)."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.model, self.group = model, group
        self.stem = None
        self._stem_buffers = []
        if hasattr(model, "_stem_prefix"):     # cavit.encoders.ViT / ViT3D: engine exists after the first forward
            self.engine = model.__dict__.get("_engine_obj")
            if self.engine is None:
                raise RuntimeError("DataParallel(ViT / ViT3D): run one forward first so that the encoder engine exists")
            named = [(k, p) for k, p in model.named_parameters() if k.startswith(model._stem_prefix)]
            self.stem = TensorGradReducer([p for _, p in named], group)
            # BatchNorm running statistics of the ViT3D stem: torch / Lightning DDP (broadcast_buffers=True) re-broadcasts
            # rank 0's buffers at every forward; so does __call__ below (a few small tensors)
            self._stem_buffers = [b for k, b in model.named_buffers() if k.startswith(model._stem_prefix)]
            self.stem.broadcast([p for _, p in named] + self._stem_buffers)
        else:
            self.engine = model.engine()
        self.comm_stream = torch.cuda.Stream(device=self.engine.device)
        self.min_slab_elems = min_slab_elems
        self._reducer: Optional[SlabReducer] = None
        if mode == "auto":
            # Overlap pays when the all-reduce is long: cfg3 (2.1 GB of fp32 gradients, 2 GPUs) 56.87 ms overlapped vs 58.29 ms
            # post. For a small model the collective is short and running it next to the backward kernels costs more (HBM / SM
            # contention) than exposing it: cfg2 (266 MB) on 4 GPUs 29.87 ms overlapped vs 29.56 ms post, same box.
            mode = self.auto_mode(self.engine.layout.total)
        if mode not in ("overlap", "post"):
            raise ValueError(f"mode must be 'auto', 'overlap' or 'post', got {mode!r}")
        self.mode = mode
        self._set_mode(mode)
        self.broadcast_parameters()

    def _set_mode(self, mode: str):
        eng = self.engine
        eng.on_range_done = eng.post_backward = None
        eng.hook_capturable = False
        if mode == "overlap":
            eng.on_range_done = self._on_range_done
            eng.hook_capturable = True      # the hook only enqueues: event record, all-reduce on comm_stream, stream wait
        else:
            eng.post_backward = self._post_backward
        self.mode = mode

    def broadcast_parameters(self, src: int = 0):
        dist.broadcast(self.engine.flat, src=src, group=self.group)
        self.engine._bf16_version = -1

    def _on_range_done(self, tag: str, start: int, end: int):
        if self._reducer is None:
            self._reducer = SlabReducer(self.min_slab_elems, self.group)
        ready = self._reducer.add(start, end)
        if tag == "embed":  # last range of the backward pass
            ready += self._reducer.flush()
        for s, e in ready:
            self._launch(s, e)
        if tag == "embed":
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            self._reducer = None

    def _post_backward(self, flat: torch.Tensor):
        SlabReducer(0, self.group).reduce_(flat, 0, flat.numel())

    def _launch(self, start: int, end: int):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        flat = self.engine.grad
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self._reducer.reduce_(flat, start, end)

    def close(self):
        """Drop every CUDA graph that recorded collectives of this process group. NCCL keeps a communicator alive while a
        captured graph still references it: `dist.destroy_process_group()` blocks until those graphs are gone, so call
        this (or delete the model) before tearing the process group down."""
        eng = self.engine
        eng.on_range_done = eng.post_backward = None
        eng.hook_capturable = False
        eng.drop_graphs()
        if self.stem is not None:
            self.stem.remove()
        torch.cuda.synchronize(eng.device)

    def __call__(self, img, labels):
        eng = self.engine
        if self.mode == "overlap" and eng._hook_capture_failed and eng.layout.total * 4 < (512 << 20):
            self._set_mode("post")          # eager launches would cost a small model more than the exposed all-reduce
        if self._stem_buffers and self.model.training:
            self.stem.broadcast(self._stem_buffers)
        return self.model(img, labels)

    def parameters(self):
        return self.model.parameters()
