"""Optimizer step on the engine's flat slabs.

The reference trains with ``torch.optim.Adam(self.parameters(), lr, weight_decay)`` and a per-epoch
``CosineAnnealingLR(T_max, eta_min)`` (/root/reference/model_cross.py:276-292, modelv3.py:211-227). torch's Adam
walks the 250+ parameter tensors (multi-tensor foreach kernels); here the parameters, their gradients and both
moment buffers are four flat fp32 slabs with identical offsets, so one launch of ``cavit_adam_step`` updates the whole
model (16 B read + 12 B written per parameter) and, in the same pass, refreshes the bf16 GEMM-operand copy that the
next forward reads — the per-step cast kernel disappears.

``FusedAdam`` covers the parameters that live in the engine's flat buffer (all of ModelCross / ModelVIT; the
transformer cores of ViT / ViT3D — their CNN stems stay with a torch optimizer).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _abi, ops


class FusedAdam:
    """Adam with L2-style weight decay (torch.optim.Adam semantics) over ``model``'s flat parameter slab."""

    def __init__(self, model, lr: Optional[float] = None, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: Optional[float] = None, own_operands: bool = True):
        eng = model.engine() if not hasattr(model, "_stem_prefix") else model._engine_obj
        if eng is None:
            raise _abi.CavitError("FusedAdam: run one forward first so that the encoder engine exists")
        self.engine = eng
        self.lr = float(model.lr if lr is None else lr)
        self.base_lr = self.lr
        self.weight_decay = float(getattr(model, "weight_decay", 0.0) if weight_decay is None else weight_decay)
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = torch.zeros_like(eng.flat)
        self.exp_avg_sq = torch.zeros_like(eng.flat)
        self.step_count = 0
        self._acc = None       # flat gradient slab assembled from `.grad` tensors that do not alias the engine's buffer
        # the step rewrites the bf16 operand copy itself; the engine then skips its per-forward cast
        self.own_operands = bool(own_operands) and not eng.split   # fp32 mode: the engine re-splits hi / lo planes itself
        if self.own_operands:
            eng.refresh_operands()
            eng.operands_external = True

    def _gradient_slab(self) -> torch.Tensor:
        """The flat gradient the update reads. Normally the engine's buffer of the last backward (every ``p.grad`` is a
        view of it). Under gradient accumulation (several backward() calls without zero_grad) autograd keeps SUMMING into
        the ``.grad`` tensors of the first backward while the engine writes each later backward into another buffer
        (Engine._next_grad_buffer), so the accumulated gradient lives in ``p.grad`` only: it is gathered into a flat
        scratch slab here (same offsets), so that the step never silently uses just the last micro-batch."""
        eng = self.engine
        base = eng.grad.data_ptr()
        stray = [(key, p) for key, p in eng.params.items()
                 if p.grad is not None and p.grad.data_ptr() != base + 4 * eng.layout.slots[key][0]]
        if not stray:
            return eng.grad
        if self._acc is None:
            self._acc = torch.empty_like(eng.flat)
        self._acc.copy_(eng.grad)
        for key, p in stray:
            off, shp = eng.layout.slots[key]
            self._acc[off:off + p.numel()].view(shp).copy_(p.grad)
        return self._acc

    def step(self, grad_scale: float = 1.0):
        """One update from the parameters' gradients: the engine's flat buffer of the last backward (after the
        data-parallel all-reduce when cavit.ddp is attached), or — when gradients were accumulated over several
        backward() calls — the sums autograd left in ``p.grad``. grad_scale multiplies the gradients
        (e.g. 1 / accumulation steps)."""
        eng = self.engine
        if getattr(eng, "grad", None) is None:
            raise _abi.CavitError("FusedAdam.step() before any backward()")
        if not eng._params_in_place():
            eng.adopt_parameters()
        self.step_count += 1
        with torch.cuda.device(eng.device):
            ops.adam_step(eng.flat, self._gradient_slab(), self.exp_avg, self.exp_avg_sq,
                          eng.flat_bf16 if self.own_operands else None, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1],
                          eps=self.eps, weight_decay=self.weight_decay, step=self.step_count, grad_scale=grad_scale)

    def zero_grad(self, set_to_none: bool = True):
        """Every backward overwrites the flat gradient buffer, so there is nothing to clear; `.grad` views are dropped
        so that the engine does not switch to a second buffer (see Engine._next_grad_buffer)."""
        for p in self.engine.params.values():
            p.grad = None

    def sync_operands(self):
        """Call after changing parameters outside this optimizer (load_state_dict, manual edits) when own_operands."""
        ops.cast_bf16(self.engine.flat, self.engine.flat_bf16)

    def state_dict(self):
        return {"step": self.step_count, "lr": self.lr, "base_lr": self.base_lr, "exp_avg": self.exp_avg.clone(),
                "exp_avg_sq": self.exp_avg_sq.clone()}

    def load_state_dict(self, sd):
        self.step_count, self.lr, self.base_lr = int(sd["step"]), float(sd["lr"]), float(sd["base_lr"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])


class CosineAnnealing:
    """``torch.optim.lr_scheduler.CosineAnnealingLR(T_max, eta_min)`` stepped once per epoch
    (/root/reference/model_cross.py:280-291), in closed form."""

    def __init__(self, optimizer: FusedAdam, T_max: int, eta_min: float = 0.0):
        self.opt, self.T_max, self.eta_min, self.epoch = optimizer, int(T_max), float(eta_min), 0

    def lr_at(self, epoch: int) -> float:
        return self.eta_min + (self.opt.base_lr - self.eta_min) * (1.0 + math.cos(math.pi * epoch / self.T_max)) / 2.0

    def step(self):
        self.epoch += 1
        self.opt.lr = self.lr_at(self.epoch)
        return self.opt.lr
