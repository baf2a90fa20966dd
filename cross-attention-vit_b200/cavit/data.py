"""Host -> device batch staging for the training loop around the hot path.

The reference hands batches to the model through a ``DataLoader(pin_memory=True)`` and Lightning's
batch transfer (/root/reference/main_mist.py:184-207, 211-224): the host -> device copy of step i+1
is not overlapped with step i, and ``self.log(...)`` reads the loss back synchronously
(/root/reference/model_cross.py:236-244). Here both transfers leave the critical path:

* ``DevicePrefetcher`` copies batch i+1 from pinned host memory on a dedicated copy stream while
  the kernels of batch i run (two device slots, event-ordered, no host synchronisation);
* ``ScalarReadback`` returns each step's loss through a pinned host slot, one step late, so the
  host never waits for the step it has just enqueued.

Both are plain torch stream/event plumbing (no kernels of their own).
"""
from __future__ import annotations

from collections import deque
from typing import Iterable, Iterator, Optional, Tuple

import torch

from . import _abi


class DevicePrefetcher:
    """Iterate device copies of ``(img, labels)`` host batches, copying one batch ahead.

    ``batches``: iterable of (img, labels) CPU tensors (pinned memory for truly asynchronous copies).
    Tensors yielded for batch i stay valid until batch i + ``depth`` is requested.
    """

    def __init__(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.CavitError("DevicePrefetcher needs a CUDA device (there is no CPU path)")
        if depth < 2:
            raise _abi.CavitError("depth must be >= 2 (one slot in use, one being filled)")
        self._it: Iterator = iter(batches)
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots = [None] * depth          # (img_dev, labels_dev)
        self._copied = [None] * depth         # event: H2D copy of the slot finished (copy stream)
        self._released = [None] * depth       # event: consumer is past the slot (compute stream)
        self._queue: deque = deque()          # slot indices holding prefetched batches, oldest first
        self._next_slot = 0
        self._in_use: Optional[int] = None
        self.h2d_bytes = 0
        self._fill()

    def _issue(self) -> bool:
        try:
            img_h, labels_h = next(self._it)
        except StopIteration:
            return False
        k = self._next_slot
        self._next_slot = (k + 1) % self.depth
        slot = self._slots[k]
        if slot is None or slot[0].shape != img_h.shape or slot[0].dtype != img_h.dtype or slot[1].shape != labels_h.shape:
            slot = (torch.empty(img_h.shape, dtype=img_h.dtype, device=self.device),
                    torch.empty(labels_h.shape, dtype=labels_h.dtype, device=self.device))
            self._slots[k] = slot
        with torch.cuda.stream(self.copy_stream):
            if self._released[k] is not None:     # do not overwrite a slot the compute stream may still read
                self.copy_stream.wait_event(self._released[k])
            slot[0].copy_(img_h, non_blocking=True)
            slot[1].copy_(labels_h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._copied[k] = ev
        self.h2d_bytes += img_h.numel() * img_h.element_size() + labels_h.numel() * labels_h.element_size()
        self._queue.append(k)
        return True

    def _fill(self):
        while len(self._queue) < self.depth - 1 and self._issue():
            pass

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self._in_use is not None:              # everything enqueued so far has consumed the previous batch
            ev = torch.cuda.Event()
            ev.record(cur)
            self._released[self._in_use] = ev
            self._in_use = None
        if not self._queue and not self._issue():
            raise StopIteration
        k = self._queue.popleft()
        cur.wait_event(self._copied[k])
        self._in_use = k
        self._fill()                              # start the next copy before the caller launches this step
        return self._slots[k]


class ScalarReadback:
    """Device scalar -> host float through pinned memory, without stalling the enqueueing thread.

    ``push(t)`` enqueues an asynchronous copy of the 0-dim / 1-element tensor ``t``; ``pop()`` returns the
    oldest pushed value as a float (waiting only for that copy). Reading step i-1's loss after enqueueing
    step i keeps the GPU busy across steps while every step's result still reaches the host."""

    def __init__(self, device, depth: int = 4):
        self.device = torch.device(device)
        self._host = torch.empty(depth, dtype=torch.float32).pin_memory()
        self._events = [None] * depth
        self._head = 0
        self._tail = 0
        self.depth = depth
        self.d2h_bytes = 0

    def push(self, t: torch.Tensor):
        if self._head - self._tail >= self.depth:
            raise _abi.CavitError("ScalarReadback overflow: pop() before pushing more")
        k = self._head % self.depth
        self._host[k:k + 1].copy_(t.detach().reshape(1).float(), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._events[k] = ev
        self._head += 1
        self.d2h_bytes += 4

    def pending(self) -> int:
        return self._head - self._tail

    def pop(self) -> float:
        if self._head == self._tail:
            raise _abi.CavitError("ScalarReadback.pop() on an empty queue")
        k = self._tail % self.depth
        self._events[k].synchronize()
        self._tail += 1
        return float(self._host[k])
