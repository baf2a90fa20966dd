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
            else:  # non-adjacent: flush what we have
                out = [self.pending]
                self.pending = (start, end)
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


class DataParallel:
    """Wraps a cavit model: `dp = DataParallel(model); logits, loss = dp(img, labels); loss.backward()`."""

    def __init__(self, model, group=None, min_slab_elems: int = 8 << 20, mode: str = "auto"):
        """mode = "overlap": backward runs eagerly and every finished gradient slab is all-reduced on a
        communication stream while the remaining backward kernels run (best when the all-reduce is a
        visible fraction of the step, i.e. large models / small per-GPU batches);
        mode = "post": backward replays its CUDA graph and the flat gradient buffer is all-reduced
        afterwards in one call (best when launch overhead would cost more than the un-overlapped
        all-reduce); "auto" picks "post" below 512 MB of gradients."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.model, self.group = model, group
        self.engine = model.engine()
        self.comm_stream = torch.cuda.Stream(device=self.engine.device)
        self.min_slab_elems = min_slab_elems
        self._reducer: Optional[SlabReducer] = None
        if mode == "auto":
            mode = "post" if self.engine.layout.total * 4 < (512 << 20) else "overlap"
        self.mode = mode
        if mode == "overlap":
            self.engine.on_range_done = self._on_range_done
        else:
            self.engine.post_backward = self._post_backward
        self.broadcast_parameters()

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

    def __call__(self, img, labels):
        return self.model(img, labels)

    def parameters(self):
        return self.model.parameters()
