"""Whole-model execution engine: runs ModelCross / ModelVIT forward and backward as a static
sequence of C-ABI kernel launches over pre-allocated, stream-major buffers.

Layout decisions (B200-first):
  * token streams are stored stream-major, fp32 residual stream  X[g][b*N + n][c]  (g = MRI
    sequence for ModelCross, one group for ModelVIT); every per-stream op (LayerNorm, GEMM,
    attention) is ONE grouped launch over all streams (blockIdx / TMA coordinate = stream);
  * all parameters live in one flat fp32 master buffer, ordered so that the per-stream copies of
    the same weight are adjacent ([stream][out][in]); the nn.Parameters of the drop-in module are
    views into it (state_dict keys unchanged). One cast kernel per step refreshes the flat bf16
    operand copy; wgrad kernels write straight into a flat fp32 gradient buffer with the same
    layout (param.grad are views; the DP all-reduce works on contiguous slabs);
  * GEMM operands are bf16 (fp32 accumulate in TMEM); the residual stream, LayerNorm / softmax
    statistics, LSE, logits and loss stay fp32.

Reference semantics followed: /root/reference/model_cross.py:186-212 (ModelCross.forward),
:128-148 (MultiScaleBlock), :111-114 (CrossAttentionBlock), :69-72 (SelfAttentionBlock);
/root/reference/modelv3.py:123-147 (ModelVIT.forward).
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch

from . import _abi, ops
from ._abi import (EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RELU, EPI_BIAS_RESID, EPI_EMBED, EPI_GELU_BWD, EPI_NONE,
                   EPI_RELU_BWD)

BF16, F32 = torch.bfloat16, torch.float32
_ALIGN = 64  # elements; keeps every segment 16-byte aligned in bf16 and 256-byte aligned in fp32


def _num_patches(cfg) -> int:
    D, H, W = cfg.img_size
    dp, hp, wp = cfg.patch_size
    return (D // dp) * (H // hp) * (W // wp)


class ParamLayout:
    """Flat packing of the model parameters. `segments[name] = (offset, shape)`; `slots` maps every
    state_dict key to (segment, element offset inside the flat buffer, shape)."""

    def __init__(self):
        self.segments: "OrderedDict[str, Tuple[int, Tuple[int, ...]]]" = OrderedDict()
        self.slots: "OrderedDict[str, Tuple[int, Tuple[int, ...]]]" = OrderedDict()
        self.total = 0
        self.layer_ranges: List[Tuple[str, int, int]] = []  # (tag, start, end) in backward-completion order

    def add(self, name: str, shape: Tuple[int, ...], keys: List[Tuple[str, int, Tuple[int, ...]]]):
        """Add a packed segment; keys = [(state_dict key, element offset within the segment, shape)]."""
        n = 1
        for s in shape:
            n *= s
        off = self.total
        self.segments[name] = (off, tuple(shape))
        for key, rel, shp in keys:
            self.slots[key] = (off + rel, tuple(shp))
        self.total = off + ((n + _ALIGN - 1) // _ALIGN) * _ALIGN
        return off


def build_layout_encoder(kind: str, cfg) -> ParamLayout:
    """Flat packing for the transformer cores of `ViT` (kind 'cnnvit', /root/reference/model.py:79-286: pre-norm,
    separate biased q/k/v Linears packed here as one [3C, C] operand, eps 1e-6, final encoder norm, single-logit
    head) and `ViT3D` (kind 'vit3d', /root/reference/modelv2.py:61-87,187-241: nn.TransformerEncoderLayer,
    post-norm, packed biased in-proj, ReLU FFN; head LN, Linear(C, C/8), Linear(C/8, classes)). The CNN stems'
    parameters are not in this buffer (they stay with torch)."""
    C, F = cfg.hidden_dim, cfg.mlp_dim
    lay = ParamLayout()
    start = lay.total
    if kind == "cnnvit":
        Np, P = cfg.patches_per_modality, cfg.patch_dim
        lay.add("pos", (Np + 1, C), [("embeddings.positional_embedding", 0, (1, Np + 1, C))])
        lay.add("cls", (C,), [("embeddings.class_token", 0, (1, 1, C))])
        lay.add("embed.w", (C, P), [("embeddings.patch_embed.weight", 0, (C, cfg.in_channels) + tuple(cfg.grid))])
        lay.add("embed.b", (C,), [("embeddings.patch_embed.bias", 0, (C,))])
    else:
        N = cfg.num_tokens
        lay.add("pos", (N, C), [("pos_embed", 0, (1, N, C))])
        if getattr(cfg, "has_cls", True):
            lay.add("cls", (C,), [("cls_token", 0, (1, 1, C))])
    lay.layer_ranges.append(("embed", start, lay.total))
    for l in range(cfg.num_layers):
        start = lay.total
        tag = f"L{l}"
        if kind == "cnnvit":
            p = f"encoder.layers.{l}."
            lay.add(f"{tag}.ln1.w", (1, C), [(p + "attention_norm.weight", 0, (C,))])
            lay.add(f"{tag}.ln1.b", (1, C), [(p + "attention_norm.bias", 0, (C,))])
            lay.add(f"{tag}.wqkv", (1, 3 * C, C), [(p + f"multi_head.{nm}.weight", i * C * C, (C, C))
                                                   for i, nm in enumerate(("query", "key", "value"))])
            lay.add(f"{tag}.bqkv", (1, 3 * C), [(p + f"multi_head.{nm}.bias", i * C, (C,))
                                                for i, nm in enumerate(("query", "key", "value"))])
            lay.add(f"{tag}.wo", (1, C, C), [(p + "multi_head.out.weight", 0, (C, C))])
            lay.add(f"{tag}.bo", (1, C), [(p + "multi_head.out.bias", 0, (C,))])
            lay.add(f"{tag}.ln2.w", (1, C), [(p + "ffn_norm.weight", 0, (C,))])
            lay.add(f"{tag}.ln2.b", (1, C), [(p + "ffn_norm.bias", 0, (C,))])
            lay.add(f"{tag}.w1", (1, F, C), [(p + "ffn.fc1.weight", 0, (F, C))])
            lay.add(f"{tag}.b1", (1, F), [(p + "ffn.fc1.bias", 0, (F,))])
            lay.add(f"{tag}.w2", (1, C, F), [(p + "ffn.fc2.weight", 0, (C, F))])
            lay.add(f"{tag}.b2", (1, C), [(p + "ffn.fc2.bias", 0, (C,))])
        else:
            p = f"transformer.transformer.layers.{l}."
            lay.add(f"{tag}.wqkv", (1, 3 * C, C), [(p + "self_attn.in_proj_weight", 0, (3 * C, C))])
            lay.add(f"{tag}.bqkv", (1, 3 * C), [(p + "self_attn.in_proj_bias", 0, (3 * C,))])
            lay.add(f"{tag}.wo", (1, C, C), [(p + "self_attn.out_proj.weight", 0, (C, C))])
            lay.add(f"{tag}.bo", (1, C), [(p + "self_attn.out_proj.bias", 0, (C,))])
            lay.add(f"{tag}.ln1.w", (1, C), [(p + "norm1.weight", 0, (C,))])
            lay.add(f"{tag}.ln1.b", (1, C), [(p + "norm1.bias", 0, (C,))])
            lay.add(f"{tag}.w1", (1, F, C), [(p + "linear1.weight", 0, (F, C))])
            lay.add(f"{tag}.b1", (1, F), [(p + "linear1.bias", 0, (F,))])
            lay.add(f"{tag}.w2", (1, C, F), [(p + "linear2.weight", 0, (C, F))])
            lay.add(f"{tag}.b2", (1, C), [(p + "linear2.bias", 0, (C,))])
            lay.add(f"{tag}.ln2.w", (1, C), [(p + "norm2.weight", 0, (C,))])
            lay.add(f"{tag}.ln2.b", (1, C), [(p + "norm2.bias", 0, (C,))])
        lay.layer_ranges.append((tag, start, lay.total))
    start = lay.total
    if kind == "cnnvit":
        lay.add("fin.ln.w", (1, C), [("encoder.encoder_norm.weight", 0, (C,))])
        lay.add("fin.ln.b", (1, C), [("encoder.encoder_norm.bias", 0, (C,))])
        lay.add("head.w", (C,), [("final.weight", 0, (1, C))])
        lay.add("head.b", (1,), [("final.bias", 0, (1,))])
    else:
        Fh = cfg.head_dim_hidden
        lay.add("fin.ln.w", (1, C), [("mlp_head.0.weight", 0, (C,))])
        lay.add("fin.ln.b", (1, C), [("mlp_head.0.bias", 0, (C,))])
        lay.add("head.w1", (1, Fh, C), [("mlp_head.1.weight", 0, (Fh, C))])
        lay.add("head.b1", (1, Fh), [("mlp_head.1.bias", 0, (Fh,))])
        lay.add("head.w2", (1, cfg.num_classes, Fh), [("mlp_head.2.weight", 0, (cfg.num_classes, Fh))])
        lay.add("head.b2", (1, cfg.num_classes), [("mlp_head.2.bias", 0, (cfg.num_classes,))])
    lay.layer_ranges.append(("head", start, lay.total))
    return lay


def build_layout(kind: str, cfg) -> ParamLayout:
    if kind in ("cnnvit", "vit3d"):
        return build_layout_encoder(kind, cfg)
    C, F, H = cfg.hidden_dim, cfg.mlp_dim, cfg.num_heads
    M = cfg.num_modalities
    Np = _num_patches(cfg)
    P = cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
    lay = ParamLayout()
    N = (Np + 1) if kind == "cross" else (Np * M + 1)
    start = lay.total
    lay.add("pos", (N, C), [("pos_embedding", 0, (1, N, C))])
    lay.add("cls", (C,), [("cls_token", 0, (1, 1, C))])
    lay.add("embed.w", (C, P), [("patch_to_embedding.weight", 0, (C, P))])
    lay.add("embed.b", (C,), [("patch_to_embedding.bias", 0, (C,))])
    lay.layer_ranges.append(("embed", start, lay.total))

    def grouped(seg, shape_one, keys):
        n = 1
        for s in shape_one:
            n *= s
        lay.add(seg, (len(keys),) + tuple(shape_one), [(k, i * n, shape_one) for i, k in enumerate(keys)])

    def self_block(tag, prefixes):
        start = lay.total
        grouped(f"{tag}.ln1.w", (C,), [p + "0.norm.weight" if kind == "vit" else p + "attn.norm.weight" for p in prefixes])
        grouped(f"{tag}.ln1.b", (C,), [p + "0.norm.bias" if kind == "vit" else p + "attn.norm.bias" for p in prefixes])
        a = "0.fn." if kind == "vit" else "attn.fn."
        f = "2." if kind == "vit" else "ffn."
        grouped(f"{tag}.wqkv", (3 * C, C), [p + a + "to_qkv.weight" for p in prefixes])
        if H != 1:
            grouped(f"{tag}.wo", (C, C), [p + a + "to_out.0.weight" for p in prefixes])
            grouped(f"{tag}.bo", (C,), [p + a + "to_out.0.bias" for p in prefixes])
        grouped(f"{tag}.ln2.w", (C,), [p + f + "norm.weight" for p in prefixes])
        grouped(f"{tag}.ln2.b", (C,), [p + f + "norm.bias" for p in prefixes])
        grouped(f"{tag}.w1", (F, C), [p + f + "fn.net.0.weight" for p in prefixes])
        grouped(f"{tag}.b1", (F,), [p + f + "fn.net.0.bias" for p in prefixes])
        grouped(f"{tag}.w2", (C, F), [p + f + "fn.net.3.weight" for p in prefixes])
        grouped(f"{tag}.b2", (C,), [p + f + "fn.net.3.bias" for p in prefixes])
        lay.layer_ranges.append((tag, start, lay.total))

    if kind == "cross":
        K = len(cfg.attn_order)
        l = 0
        for mb in range(cfg.num_multi_blocks):
            for sb in range(cfg.num_self_blocks):
                self_block(f"L{l}", [f"transformer.{mb}.blocks.{m}.{sb}." for m in range(M)])
                l += 1
            if K:
                start = lay.total
                pre = [f"transformer.{mb}.fusion.{k}." for k in range(K)]
                tag = f"X{mb}"
                grouped(f"{tag}.lnA.w", (C,), [p + "attn.norm.weight" for p in pre])
                grouped(f"{tag}.lnA.b", (C,), [p + "attn.norm.bias" for p in pre])
                grouped(f"{tag}.wq", (C, C), [p + "attn.fn.wq.weight" for p in pre])
                grouped(f"{tag}.bq", (C,), [p + "attn.fn.wq.bias" for p in pre])
                # wk | wv of one fusion adjacent: one [2C, C] operand per fusion
                keys = []
                for k, p in enumerate(pre):
                    keys.append((p + "attn.fn.wk.weight", (2 * k) * C * C, (C, C)))
                    keys.append((p + "attn.fn.wv.weight", (2 * k + 1) * C * C, (C, C)))
                lay.add(f"{tag}.wkv", (K, 2 * C, C), keys)
                keys = []
                for k, p in enumerate(pre):
                    keys.append((p + "attn.fn.wk.bias", (2 * k) * C, (C,)))
                    keys.append((p + "attn.fn.wv.bias", (2 * k + 1) * C, (C,)))
                lay.add(f"{tag}.bkv", (K, 2 * C), keys)
                grouped(f"{tag}.wp", (C, C), [p + "attn.fn.proj.weight" for p in pre])
                grouped(f"{tag}.bp", (C,), [p + "attn.fn.proj.bias" for p in pre])
                grouped(f"{tag}.lnF.w", (C,), [p + "ffn.norm.weight" for p in pre])
                grouped(f"{tag}.lnF.b", (C,), [p + "ffn.norm.bias" for p in pre])
                grouped(f"{tag}.w1", (F, C), [p + "ffn.fn.net.0.weight" for p in pre])
                grouped(f"{tag}.b1", (F,), [p + "ffn.fn.net.0.bias" for p in pre])
                grouped(f"{tag}.w2", (C, F), [p + "ffn.fn.net.3.weight" for p in pre])
                grouped(f"{tag}.b2", (C,), [p + "ffn.fn.net.3.bias" for p in pre])
                lay.layer_ranges.append((tag, start, lay.total))
        start = lay.total
        grouped("fin.ln.w", (C,), [f"norm.{m}.weight" for m in range(M)])
        grouped("fin.ln.b", (C,), [f"norm.{m}.bias" for m in range(M)])
        grouped("head.w1", (F, C), [f"mlp_head.{m}.0.weight" for m in range(M)])
        grouped("head.b1", (F,), [f"mlp_head.{m}.0.bias" for m in range(M)])
        grouped("head.w2", (cfg.num_classes, F), [f"mlp_head.{m}.3.weight" for m in range(M)])
        grouped("head.b2", (cfg.num_classes,), [f"mlp_head.{m}.3.bias" for m in range(M)])
        lay.layer_ranges.append(("head", start, lay.total))
    else:
        for l in range(cfg.num_layers):
            self_block(f"L{l}", [f"transformer.layers.{l}."])
        start = lay.total
        grouped("fin.ln.w", (C,), ["mlp_head.0.weight"])
        grouped("fin.ln.b", (C,), ["mlp_head.0.bias"])
        grouped("head.w1", (F, C), ["mlp_head.1.weight"])
        grouped("head.b1", (F,), ["mlp_head.1.bias"])
        grouped("head.w2", (cfg.num_classes, F), ["mlp_head.4.weight"])
        grouped("head.b2", (cfg.num_classes,), ["mlp_head.4.bias"])
        lay.layer_ranges.append(("head", start, lay.total))
    return lay


class Engine:
    """Executes one model (kind = 'cross' | 'vit') on one device. Not thread-safe."""

    def __init__(self, kind: str, cfg, named_params: "OrderedDict[str, torch.nn.Parameter]", device, precision: str = "bf16"):
        assert kind in ("cross", "vit", "cnnvit", "vit3d")
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:   # an index-less device means the CURRENT device, not cuda:0
            device = torch.device("cuda", torch.cuda.current_device())
        _abi.require_device(device.index)
        self.kind, self.cfg, self.device = kind, cfg, device
        if precision not in ("bf16", "fp32"):
            raise _abi.CavitError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        if precision == "fp32" and kind not in ("cross", "vit"):
            raise _abi.CavitError("the fp32-tolerance mode covers ModelCross / ModelVIT (pre-norm blocks); ViT / ViT3D cores run in bf16 mode")
        self.precision = precision
        self.split = precision == "fp32"      # GEMM operands as bf16 hi + lo planes (3 MMAs per product), see engine_fp32.py
        self.C, self.F, self.H = cfg.hidden_dim, cfg.mlp_dim, cfg.num_heads
        if self.C != self.H * 64:
            raise _abi.CavitError(f"cavit attention kernels are specialised for head_dim 64 (hidden_dim {self.C}, heads {self.H})")
        if self.C % 64 or self.F % 8:
            raise _abi.CavitError("hidden_dim must be a multiple of 64 and mlp_dim a multiple of 8")
        self.Mimg = cfg.num_modalities
        self.classes = cfg.num_classes
        # per-kind block structure: pre-norm + bias-free QKV + GELU (model_cross.py / modelv3.py), pre-norm + biased
        # q/k/v + eps 1e-6 (model.py:107-214), post-norm + biased in-proj + ReLU (nn.TransformerEncoderLayer, modelv2.py:72-78)
        self.encoder = kind in ("cnnvit", "vit3d")
        self.qkv_bias = self.encoder
        self.post_norm = kind == "vit3d"
        self.eps = 1e-6 if kind == "cnnvit" else 1e-5
        if kind == "cnnvit":
            self.Np, self.P = cfg.patches_per_modality, cfg.patch_dim
        elif kind == "vit3d":
            self.Np, self.P = cfg.tokens_per_modality, 8
            if cfg.head_dim_hidden % 8:
                raise _abi.CavitError("ViT3D: hidden_dim // 8 must be a multiple of 8")
        else:
            self.Np = _num_patches(cfg)
            self.P = cfg.patch_size[0] * cfg.patch_size[1] * cfg.patch_size[2]
        if self.P % 8:
            raise _abi.CavitError("patch_dim must be a multiple of 8")
        self.has_cls = bool(getattr(cfg, "has_cls", True))   # ViT3D(add_cls_token=False): mean-pooled head, no CLS row
        if self.encoder:
            if float(cfg.dropout) != 0.0:
                raise _abi.CavitError("cavit encoder cores (ViT / ViT3D) support dropout = 0 only")
            self.G, self.N = 1, self.Np * self.Mimg + int(self.has_cls)
            self.L = cfg.num_layers
            self.cls_src, self.tok_src = [], []
            self.smoothing = float(getattr(cfg, "label_smoothing", 0.0))
        elif kind == "cross":
            self.G, self.N = self.Mimg, self.Np + 1
            self.L = cfg.num_multi_blocks * cfg.num_self_blocks
            order = sorted((int(i), int(j)) for i, j in dict(cfg.attn_order).items() if int(i) < self.Mimg)
            self.cls_src = [i for i, _ in order]   # fusion k: CLS of stream cls_src[k] ...
            self.tok_src = [j for _, j in order]   # ... attends patch tokens of stream tok_src[k]
            self.smoothing = float(cfg.label_smoothing)
        else:
            self.G, self.N = 1, self.Np * self.Mimg + 1
            self.L = cfg.num_layers
            self.cls_src, self.tok_src = [], []
            self.smoothing = 0.0
        self.K = len(self.cls_src)
        self.fold_ok = (self.H in (1, 2, 3, 4, 6, 8, 12, 16) and self.C <= 1024 and self.K <= 16
                        and os.environ.get("CAVIT_XFOLD", "1") != "0")
        self.fold = False
        self.embed_fused = False
        self._img = None                       # the forward's input volumes (the fused embedding wgrad re-reads them)
        self.scale = 64 ** -0.5
        self.p_drop = float(cfg.dropout)
        self.drop = False                      # dropout active for the current forward/backward pair
        self.seed_buf = torch.zeros(1, dtype=torch.int64, device=self.device)   # per-step dropout seed (device scalar)
        self._step = 0
        self.layout = build_layout(kind, cfg)
        missing = [k for k in named_params if k not in self.layout.slots]
        extra = [k for k in self.layout.slots if k not in named_params]
        if missing or extra:
            raise _abi.CavitError(f"parameter schema mismatch: missing {missing[:3]} extra {extra[:3]}")
        self.params = named_params
        self.flat = torch.zeros(self.layout.total, dtype=F32, device=self.device)
        self.flat_bf16 = torch.zeros(self.layout.total, dtype=BF16, device=self.device)
        # fp32-tolerance mode: second bf16 plane of the operand copy (w ~ flat_bf16 + flat_bf16_lo)
        self.flat_bf16_lo = torch.zeros(self.layout.total, dtype=BF16, device=self.device) if self.split else None
        self.grad_bufs = [torch.zeros(self.layout.total, dtype=F32, device=self.device)]
        self._grad_idx = 0
        self._bf16_version = -1
        self._frozen = False
        self.operands_external = False   # cavit.optim.FusedAdam rewrites flat_bf16 itself: no per-forward cast then
        self.grad = None
        self.adopt_parameters()
        self._plan_key = None
        # (batch, train, drop) -> (buffers, forward graphs, backward graphs): a train <-> eval switch or a short last batch
        # re-uses its plan (and its captured CUDA graphs) instead of re-allocating every activation buffer. Bounded LRU.
        self._plans: "OrderedDict[Tuple, Tuple]" = OrderedDict()
        self.max_plans = int(os.environ.get("CAVIT_MAX_PLANS", "3"))
        self.saved_valid = False
        # CUDA graphs: the forward / backward kernel sequences are static for a given (batch, mode), so after
        # two eager runs they are captured once and replayed (removes ~1.5k launch calls per step from the
        # critical path). Disabled while per-launch profiling is on or a data-parallel hook is installed.
        self.use_graphs = os.environ.get("CAVIT_NO_GRAPHS", "0") != "1"
        self._fwd_graphs: Dict = {}
        self._bwd_graphs: Dict = {}
        self.on_range_done = None
        self.hook_capturable = False        # set by the owner of on_range_done when the hook only enqueues stream work
        self._hook_capture_failed = False
        self.hook_capture_error = None
        self.post_backward = None
        self.graph_launches = 0   # kernels executed through graph replays (the library counter only sees eager launches)

    # ------------------------------------------------------------------ parameters
    def adopt_parameters(self):
        """Copy current parameter values into the flat master buffer and re-point every
        nn.Parameter at its slice (idempotent; called again if a param was re-allocated)."""
        with torch.no_grad():
            for key, p in self.params.items():
                off, shp = self.layout.slots[key]
                view = self.flat[off:off + p.numel()].view(shp)
                if p.data_ptr() != view.data_ptr():
                    view.copy_(p.detach().to(device=self.device, dtype=F32))
                    p.data = view
        self._bf16_version = -1

    def _params_in_place(self) -> bool:
        for key, p in self.params.items():
            off, _ = self.layout.slots[key]
            if p.data_ptr() != self.flat.data_ptr() + 4 * off:
                return False
        return True

    def refresh_operands(self):
        """Re-derive the bf16 operand copy from the fp32 master weights. Parameters are views of the flat
        buffer whose in-place updates (optimizer steps) are not visible through any version counter of
        the flat tensor, so the cast (one launch, ~6 bytes per parameter) runs on every forward unless
        the caller froze the weights with `freeze_operands()` (inference)."""
        if not self._params_in_place():
            self.adopt_parameters()
        if self.operands_external and self._bf16_version >= 0:
            return
        if not self._frozen or self._bf16_version < 0:
            self._cast_operands()
            self._bf16_version = 1

    def _cast_operands(self):
        if self.split:
            ops.cast_split(self.flat, ops.with_lo(self.flat_bf16, self.flat_bf16_lo))
        else:
            ops.cast_bf16(self.flat, self.flat_bf16)

    def freeze_operands(self, frozen: bool = True):
        """Inference helper: skip the per-forward fp32 -> bf16 weight cast until unfrozen."""
        self._frozen = frozen
        self._bf16_version = -1

    def _seg(self, buf, name):
        off, shp = self.layout.segments[name]
        n = 1
        for s in shp:
            n *= s
        return buf[off:off + n].view(shp)

    def w(self, name):
        return self._seg(self.flat, name)

    def wb(self, name):
        t = self._seg(self.flat_bf16, name)
        if self.split:     # the bf16 operand view carries its lo plane (see ops.lo_of)
            t._lo = self._seg(self.flat_bf16_lo, name)
        return t

    def g(self, name):
        return self._seg(self.grad, name)

    def has(self, name):
        return name in self.layout.segments

    # ------------------------------------------------------------------ dropout sites
    # One id per nn.Dropout module instance on the path (model_cross.py:25,27,47,84,86,170,180,182).
    SITE_EMBED, SITE_HEAD_GELU, SITE_HEAD_LOGITS = 1, 2, 3

    @staticmethod
    def site_layer(l: int, which: str) -> int:      # which in out | gelu | fc2
        return 16 + 4 * l + {"out": 0, "gelu": 1, "fc2": 2}[which]

    @staticmethod
    def site_fusion(mb: int, which: str) -> int:    # which in attn | proj | gelu | fc2
        return 4096 + 4 * mb + {"attn": 0, "proj": 1, "gelu": 2, "fc2": 3}[which]

    def _dropout(self, mode, a, b, out, site):
        ops.dropout(mode, a, b, out, n=out.numel(), p=self.p_drop, seed=self.seed_buf, site=site)

    def _new_seed(self):
        """Fresh per-step seed derived from torch's global seed (so torch.manual_seed controls it)."""
        self._step += 1
        x = (torch.initial_seed() * 0x9E3779B97F4A7C15 + self._step * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF
        self.seed_buf.fill_(x)

    # ------------------------------------------------------------------ planning
    def _plan(self, B: int, train: bool, drop: bool = False):
        key = (B, train, drop)
        if self._plan_key == key:
            return
        if self._plan_key is not None:      # park the current plan (LRU order: most recent last)
            self._plans[self._plan_key] = (self.a, self._fwd_graphs, self._bwd_graphs, (self.fold, self.embed_fused))
            self._plans.move_to_end(self._plan_key)
            while len(self._plans) > self.max_plans:
                self._plans.popitem(last=False)
        self.saved_valid = False
        if key in self._plans:
            self.a, self._fwd_graphs, self._bwd_graphs, (self.fold, self.embed_fused) = self._plans.pop(key)
            self._plan_key = key
            self.B, self.T = B, B * self.N
            return
        self._fwd_graphs, self._bwd_graphs = {}, {}
        self.embed_fused = False
        if self.split:
            from . import engine_fp32
            self.a = engine_fp32.plan(self, B, train)
            self._plan_key = key
            self.B, self.T = B, B * self.N
            return
        dev = self.device
        G, N, C, F, H, K = self.G, self.N, self.C, self.F, self.H, self.K
        T = B * N

        def e(shape, dt=F32):
            return torch.empty(shape, dtype=dt, device=dev)

        a: Dict[str, torch.Tensor] = {}
        if self.post_norm:
            self._plan_post(a, B, train)
            self.a = a
            self._plan_key = key
            self.B, self.T = B, T
            return
        # Folded single-query cross attention (csrc/xfold.cu): no K/V projection of the fused sequence, no materialised
        # LayerNorm of it. Attention dropout breaks the fold (probabilities stop summing to one): unfolded path then.
        # ... and the folded kernels run one CTA per (fusion, sample): below ~2 CTAs per SM (small batches of long sequences,
        # cfg1 / cfg3 / cfg5) the token-parallel unfolded path is faster (cfg3: 26 ms vs 3 ms per step for the fusion backward)
        min_ctas = int(os.environ.get("CAVIT_XFOLD_MIN_CTAS", "296"))
        self.fold = bool(K) and not drop and self.fold_ok and K * B >= min_ctas
        fold = self.fold
        # K-EMBED: TMA-staged unfold fused with the embedding GEMM (csrc/embed.cu) wherever its brick geometry fits;
        # then neither the unfolded patch tensor nor the compacted token gradient exists
        # ... and where it wins: the fused kernels stream the fp32 weight tile from L2 once per 128-token brick, which is
        # cheap for a 384 x 256 projection (cfg2: step 30.32 -> 30.23 ms) and a wash at 512 x 512 (cfg5), but costs more than
        # patchify + the paired bf16 GEMM once the weights are megabytes (cfg1, 1024 x 2048: 36.60 vs 36.31 ms; cfg3,
        # 768 x 4096: 53.95 vs 52.99 ms, same box). CAVIT_EMBED_FUSED = 0 / 1 forces the choice.
        want = os.environ.get("CAVIT_EMBED_FUSED", "auto")
        self.embed_fused = (self.kind in ("cross", "vit") and want != "0" and (want == "1" or C * self.P <= 512 * 512) and
                            ops.embed_fused_supported((B, self.Mimg, 1) + tuple(self.cfg.img_size), self.cfg.patch_size, C))
        if not self.embed_fused:
            a["patches"] = e((self.Mimg * B * self.Np, self.P), BF16)
        nL = self.L if train else 1
        nX = 2 * self.L + 1 if train else 3
        a["X"] = [e((G, T, C)) for _ in range(nX)]
        for nm, shp, dt in [("xn1", (G, T, C), BF16), ("mean1", (G, T), F32), ("rstd1", (G, T), F32),
                            ("qkv", (G, T, 3 * C), BF16), ("ao", (G, T, C), BF16), ("lse", (G, B, H, N), F32),
                            ("xn2", (G, T, C), BF16), ("mean2", (G, T), F32), ("rstd2", (G, T), F32),
                            ("u", (G, T, F), BF16), ("h", (G, T, F), BF16)]:
            a[nm] = [e(shp, dt) for _ in range(nL)]
        if K:
            nF = self.cfg.num_multi_blocks if train else 1
            unfolded = [("f_xn", (K, T, C), BF16), ("f_kv", (K, T, 2 * C), BF16), ("f_q", (K, B, C), F32)]
            folded = [("f_xncls", (K, B, C), BF16), ("f_mean0", (K, B), F32), ("f_rstd0", (K, B), F32),
                      ("f_qb", (K, B, C), BF16), ("f_qp", (K, B, H * C), F32), ("f_zhat", (K, B, H * C), F32),
                      ("f_zb", (K, B, H * C), BF16), ("f_Ekv", (K, 2, C, H * C), BF16)]
            for nm, shp, dt in (folded if fold else unfolded) + [
                                ("f_cls", (K, B, C), F32), ("f_mean", (K, T), F32), ("f_rstd", (K, T), F32),
                                ("f_probs", (K, B, H, N), F32), ("f_xo", (K, B, C), F32), ("f_xob", (K, B, C), BF16),
                                ("f_y", (K, B, C), F32), ("f_yn", (K, B, C), BF16), ("f_meany", (K, B), F32),
                                ("f_rstdy", (K, B), F32), ("f_u", (K, B, F), BF16), ("f_h", (K, B, F), BF16),
                                ("f_z", (K, B, C), F32)]:
                a[nm] = [e(shp, dt) for _ in range(nF)]
            if fold:
                a["xf_scratch"] = ops.xfold_scratch(K, B, N, H, dev)
        if drop:   # bf16 branch outputs that get dropped before the residual add
            a["br"] = e((G, T, C), BF16)
            if K:
                a["br_f"] = e((K, B, C), BF16)
        a["clsn"] = e((G, B, C), BF16)
        a["meanc"], a["rstdc"] = e((G, B)), e((G, B))
        a["uh"], a["hh"] = e((G, B, F), BF16), e((G, B, F), BF16)
        a["logits"], a["loss"] = e((B, self.classes)), e((1,))
        if self.kind == "cnnvit":
            a["logits"] = e((B,))
            a["pos_exp"], a["clsn32"] = torch.zeros((N, C), dtype=F32, device=dev), e((B, C))
            if train:
                a["dpos_exp"], a["dclsn32"] = e((N, C)), e((B, C))
                a["drows"] = e((self.Mimg * B * self.Np, self.P), BF16)
                a["dfeat"] = e((self.Mimg * B, self.cfg.in_channels) + tuple(self.cfg.feat_dims))
        if train:
            a["dX"], a["dXb"] = e((G, T, C)), e((G, T, C), BF16)
            a["dbig"] = e((G, T, F), BF16)
            a["dmid"] = e((G, T, C), BF16)
            a["dqkv"] = e((G, T, 3 * C), BF16)
            a["delta"] = e((G, B, H, N))
            a["dq_acc"] = e((G, T, C))
            a["ln_ws"] = ops.ln_bwd_workspace(max(G, K, 1), C, dev)
            a["dhh"], a["duh"], a["dclsn"] = e((G, B, F), BF16), e((G, B, F), BF16), e((G, B, C), BF16)
            if not self.embed_fused:
                a["dcomp"] = e((self.Mimg * B * self.Np, C), BF16)
            if K:
                a["d_z"], a["d_zb"] = e((K, B, C)), e((K, B, C), BF16)
                a["d_u"], a["d_yn"] = e((K, B, F), BF16), e((K, B, C), BF16)
                a["d_y"], a["d_yb"] = e((K, B, C)), e((K, B, C), BF16)
                a["d_qb"] = e((K, B, C), BF16)
                if fold:
                    a["d_xob"], a["d_gz"] = e((K, B, C), BF16), e((K, B, H * C))
                    a["d_qp"], a["d_qpb"] = e((K, B, H * C)), e((K, B, H * C), BF16)
                    a["d_Ekv"] = e((K, 2, C, H * C))
                    a["d_lnA"] = e((2, K, C))          # dgamma | dbeta of the folded LayerNorm (atomically accumulated)
                    a["d_bv"] = e((K, C))
                    a["d_xnclsb"], a["d_clsq"] = e((K, B, C), BF16), e((K, B, C))
                else:
                    a["d_xo"], a["d_q"] = e((K, B, C)), e((K, B, C))
                    a["d_kv"], a["d_xn"] = e((K, T, 2 * C), BF16), e((K, T, C), BF16)
                    a["d_xncls"] = e((K, B, C))
        self.a = a
        self._plan_key = key
        self.B, self.T = B, T

    # ------------------------------------------------------------------ GEMM helpers
    def _split_for(self, G, N_out, K_in, T):
        """Split-K factor of a wgrad GEMM dW[N_out, K_in] = dY^T X (reduction over T tokens): the (group, tile, split)
        work items should fill whole waves of the 148 persistent CTAs (a 2.2-wave launch wastes 27 % of the last
        wave) while every split keeps >= 8 k-blocks of 64 tokens."""
        bn = 256 if (K_in >= 256 and (K_in % 256 == 0 or K_in > 1024)) else 128     # mirrors cavit_gemm's N tile
        if K_in % 192 == 0 and K_in % 256 != 0 and os.environ.get("CAVIT_WGRAD_BN192", "1") != "0":
            bn = 192
        tiles = G * ((N_out + 127) // 128) * ((K_in + bn - 1) // bn)
        kb = (T + 63) // 64
        sms = 148
        best, best_score = 1, -1.0
        for s_ in range(1, 65):
            if s_ > 1 and kb // s_ < 8:
                break
            items = tiles * s_
            waves = -(-items // sms)
            eff = items / (waves * sms)
            # best wave efficiency, then the smaller split (a single full wave measured as good as two: wgrad out 144 items
            # 0.078 ms vs 288 items 0.110 ms, wgrad qkv 144 items 0.190 vs 288 items 0.196)
            score = eff - 0.004 * s_
            if score > best_score:
                best, best_score = s_, score
        return best

    def _fwd(self, x, w, out, *, G, T, N, K, epi=EPI_NONE, bias=None, resid=None, aux=None, lda=None, a_gs=None):
        lda = K if lda is None else lda
        a_gs = T * lda if a_gs is None else a_gs
        ops.gemm(x, w, out, M=T, N=N, K=K, groups=G, lda=lda, ldb=K, ldo=N, a_gs=a_gs, b_gs=N * K, out_gs=T * N,
                 epi=epi, bias=bias, bias_gs=N, resid=resid, ldr=N, resid_gs=T * N, aux=aux, ldaux=N, aux_gs=T * N)

    def _dgrad(self, dy, w, out, *, G, T, N, K, epi=EPI_NONE, aux=None, bias=None, resid=None):
        """dX[T,K] = dY[T,N] W[N,K] (+ epilogue: EPI_BIAS_RESID adds an fp32 gradient stream, `bias` then is a zero vector)."""
        ops.gemm(dy, w, out, M=T, N=K, K=N, groups=G, a_mn=False, b_mn=True, lda=N, ldb=K, ldo=K, a_gs=T * N,
                 b_gs=N * K, out_gs=T * K, epi=epi, aux=aux, ldaux=K, aux_gs=T * K, bias=bias, bias_gs=K, resid=resid,
                 ldr=K, resid_gs=T * K)

    def _wgrad(self, dy, x, dw, *, G, T, N, K, ldx=None, x_gs=None):
        """dW[N,K] = dY[T,N]^T X[T,K] (fp32, written into the flat gradient buffer)."""
        ldx = K if ldx is None else ldx
        x_gs = T * ldx if x_gs is None else x_gs
        ops.gemm(dy, x, dw, M=N, N=K, K=T, groups=G, a_mn=True, b_mn=True, lda=N, ldb=ldx, ldo=K, a_gs=T * N,
                 b_gs=x_gs, out_gs=N * K, split_k=self._split_for(G, N, K, T))

    def _colsum(self, x, out, *, G, T, N):
        ops.colsum_bf16(x, out, rows=T, C_=N, groups=G)

    # ------------------------------------------------------------------ forward
    def forward(self, img: torch.Tensor, labels: torch.Tensor, train: bool, drop: bool = False):
        """Validates inputs, then runs the forward kernel sequence (eagerly, or by replaying its CUDA graph).
        drop: apply the configured dropout (module in training mode with dropout > 0).
        Every launch goes to the current stream of THIS engine's device, whatever the caller's current device is."""
        with torch.cuda.device(self.device):
            return self._forward_checked(img, labels, train, drop)

    def _forward_checked(self, img: torch.Tensor, labels: torch.Tensor, train: bool, drop: bool = False):
        cfg = self.cfg
        if img.dtype != F32 or not img.is_cuda:
            raise _abi.CavitError("img must be a float32 CUDA tensor")
        if img.device != self.device:
            raise _abi.CavitError(f"img is on {img.device}, the model on {self.device}")
        if self.kind == "cnnvit":     # stem feature maps [M*B, Cin, A, Bd, Cd] (modality-major), float targets
            want = (cfg.in_channels,) + tuple(cfg.feat_dims)
            if img.dim() != 5 or tuple(img.shape[1:]) != want or img.shape[0] % self.Mimg:
                raise _abi.CavitError(f"feature maps must be [{self.Mimg}*B, {want}], got {tuple(img.shape)}")
            B = img.shape[0] // self.Mimg
            labels = labels.to(device=self.device, dtype=F32).contiguous()
        elif self.kind == "vit3d":    # stem features of all modalities on the token axis, channel-major [B, C, M*S]
            if img.dim() != 3 or img.shape[1] != self.C or img.shape[2] != self.N - int(self.has_cls):
                raise _abi.CavitError(f"features must be [B, {self.C}, {self.N - int(self.has_cls)}], got {tuple(img.shape)}")
            B = img.shape[0]
            labels = labels.to(device=self.device, dtype=torch.int64).contiguous()
        else:
            if img.dim() != 6 or img.shape[1] != self.Mimg or tuple(img.shape[3:]) != tuple(cfg.img_size) or img.shape[2] != 1:
                raise _abi.CavitError(f"img must be [B, {self.Mimg}, 1, {tuple(cfg.img_size)}], got {tuple(img.shape)}")
            B = img.shape[0]
            labels = labels.to(device=self.device, dtype=torch.int64).contiguous()
        img = img.contiguous()
        if labels.numel() != B:
            raise _abi.CavitError(f"labels must have {B} entries, got {labels.numel()}")
        drop = bool(drop and self.p_drop > 0.0)
        if drop and self.split:
            raise _abi.CavitError("the fp32-tolerance mode runs with dropout = 0 (the parity configuration); "
                                  "use precision='bf16' to train with dropout")
        self._plan(B, train, drop)
        self.drop = drop
        if drop:
            self._new_seed()
        if not self._params_in_place():
            self.adopt_parameters()
        if self.operands_external and self._bf16_version < 0:   # parameters were (re)loaded behind the optimizer's back
            self._cast_operands()
            self._bf16_version = 1
        if not (self.use_graphs and ops.PROFILE is None):
            return self._forward_impl(img, labels, train)
        st = self._fwd_graphs.setdefault((B, train, drop, self.operands_external), {"runs": 0, "graph": None})
        if st["graph"] is None:
            st["runs"] += 1
            if st["runs"] <= 2:   # eager warm-up: sets kernel attributes, fills the TMA descriptor cache
                return self._forward_impl(img, labels, train)
            st["img"], st["labels"] = torch.empty_like(img), torch.empty_like(labels)
            st["img"].copy_(img)
            st["labels"].copy_(labels)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            n0 = _abi.launch_count()
            with torch.cuda.graph(graph):
                self._forward_impl(st["img"], st["labels"], train, force_cast=True)
            st["launches"] = _abi.launch_count() - n0   # kernels recorded in the graph
            st["graph"] = graph
        st["img"].copy_(img)
        st["labels"].copy_(labels)
        st["graph"].replay()
        self.graph_launches += st["launches"]
        self._labels = st["labels"]
        self._img = st["img"]
        self.saved_valid = train
        return self.a["logits"], self.a["loss"]

    def _forward_impl(self, img: torch.Tensor, labels: torch.Tensor, train: bool, force_cast: bool = False):
        cfg = self.cfg
        B = self.B       # (the leading axis of `img` is modality x sample for the CNN-stem ViT)
        if force_cast:   # inside a graph the operand refresh is unconditional (unless an optimizer owns the bf16 copy)
            if not self.operands_external:
                self._cast_operands()
        else:
            self.refresh_operands()
        if self.split:
            from . import engine_fp32
            return engine_fp32.forward_impl(self, img, labels, train)
        a, G, N, C, F, H, T, K = self.a, self.G, self.N, self.C, self.F, self.H, self.T, self.K
        if self.post_norm:
            return self._forward_post(img, labels, train)
        # ---- tokenisation: unfold -> embedding GEMM (+bias +pos, CLS-skipping row map) -> CLS rows
        np_seq = self.N - 1
        X0 = a["X"][0]
        if self.kind == "cnnvit":
            # Conv3d(k = stride) patch embedding of the stem's feature maps (model.py:84,95-105,258): every modality gets
            # the SAME positional rows 1.., the CLS row exists once -> positional table expanded to the token axis
            ops.conv_patch_rows(img, a["patches"], M=self.Mimg, B=B, Cin=cfg.in_channels, dims=cfg.feat_dims, grid=cfg.grid)
            pos = a["pos_exp"]
            for m in range(self.Mimg):
                ops.gather_rows_f32(self.w("pos")[1:], pos[1 + m * self.Np:], rows=self.Np, C_=C, groups=1,
                                    src_row_stride=C, src_gs=0, dst_row_stride=C, dst_gs=0)
        elif self.embed_fused:
            self._img = img
            ops.embed_fused_fwd(img, self.w("embed.w"), self.w("embed.b"), self.w("pos"), X0, patch_size=cfg.patch_size,
                                C_=C, sample_major=(self.kind == "vit"))
        else:
            ops.patchify(img, a["patches"], patch_size=cfg.patch_size, sample_major=(self.kind == "vit"))
            pos = self.w("pos")
        if not self.embed_fused:
            ops.gemm(a["patches"], self.wb("embed.w"), X0, M=self.Mimg * B * self.Np, N=C, K=self.P, lda=self.P,
                     ldb=self.P, ldo=C, epi=EPI_EMBED, bias=self.w("embed.b"), resid=pos, ldr=C, embed_np=np_seq)
        ops.cls_rows(self.w("cls"), self.w("pos"), X0, M=G, B=B, N=N, C_=C)
        drop = self.drop
        if drop:   # x = dropout(x) after the positional add (model_cross.py:198)
            self._dropout(ops.DROP_F32, X0, None, X0, self.SITE_EMBED)
        xi = 0
        for l in range(self.L):
            s = l if train else 0
            if train:
                x_in, x_mid, x_out = a["X"][2 * l], a["X"][2 * l + 1], a["X"][2 * l + 2]
            else:
                x_in, x_mid, x_out = a["X"][xi], a["X"][(xi + 1) % 3], a["X"][(xi + 2) % 3]
                xi = (xi + 2) % 3
            tag = f"L{l}"
            ops.ln_fwd(x_in, self.w(f"{tag}.ln1.w"), self.w(f"{tag}.ln1.b"), a["xn1"][s], a["mean1"][s], a["rstd1"][s],
                       rows_per_group=T, groups=G, C=C, eps=self.eps)
            if self.qkv_bias:   # model.py:118-120: query / key / value are biased Linears
                self._fwd(a["xn1"][s], self.wb(f"{tag}.wqkv"), a["qkv"][s], G=G, T=T, N=3 * C, K=C, epi=EPI_BIAS,
                          bias=self.w(f"{tag}.bqkv"))
            else:
                self._fwd(a["xn1"][s], self.wb(f"{tag}.wqkv"), a["qkv"][s], G=G, T=T, N=3 * C, K=C)
            ops.attn_fwd(a["qkv"][s], a["ao"][s], a["lse"][s], G=G, B=B, N=N, H=H, scale=self.scale)
            if H != 1 and drop:   # to_out = Linear, Dropout: x_mid = x_in + D(ao Wo^T + bo)
                self._fwd(a["ao"][s], self.wb(f"{tag}.wo"), a["br"], G=G, T=T, N=C, K=C, epi=EPI_BIAS, bias=self.w(f"{tag}.bo"))
                self._dropout(ops.DROP_ADD, x_in, a["br"], x_mid, self.site_layer(l, "out"))
            elif H != 1:
                self._fwd(a["ao"][s], self.wb(f"{tag}.wo"), x_mid, G=G, T=T, N=C, K=C, epi=EPI_BIAS_RESID,
                          bias=self.w(f"{tag}.bo"), resid=x_in)
            else:  # to_out = nn.Identity() when heads == 1 (model_cross.py:37,44-48): no projection, no dropout
                ops.add_bf16_f32(x_in, a["ao"][s], x_mid)
            ops.ln_fwd(x_mid, self.w(f"{tag}.ln2.w"), self.w(f"{tag}.ln2.b"), a["xn2"][s], a["mean2"][s], a["rstd2"][s],
                       rows_per_group=T, groups=G, C=C, eps=self.eps)
            self._fwd(a["xn2"][s], self.wb(f"{tag}.w1"), a["h"][s], G=G, T=T, N=F, K=C, epi=EPI_BIAS_GELU,
                      bias=self.w(f"{tag}.b1"), aux=a["u"][s])
            if drop:
                self._dropout(ops.DROP_BF16, a["h"][s], None, a["h"][s], self.site_layer(l, "gelu"))
                self._fwd(a["h"][s], self.wb(f"{tag}.w2"), a["br"], G=G, T=T, N=C, K=F, epi=EPI_BIAS, bias=self.w(f"{tag}.b2"))
                self._dropout(ops.DROP_ADD, x_mid, a["br"], x_out, self.site_layer(l, "fc2"))
            else:
                self._fwd(a["h"][s], self.wb(f"{tag}.w2"), x_out, G=G, T=T, N=C, K=F, epi=EPI_BIAS_RESID,
                          bias=self.w(f"{tag}.b2"), resid=x_mid)
            if self.kind == "cross" and K and (l + 1) % cfg.num_self_blocks == 0:
                self._fusion_fwd(l // cfg.num_self_blocks, x_out, train)
        x_fin = a["X"][2 * self.L] if train else a["X"][xi]
        self._x_fin = x_fin
        if self.kind == "cnnvit":
            # encoder_norm (eps 1e-6) on the CLS rows, final Linear(C, 1), BCEWithLogits (model.py:202-206,224,273-286)
            ops.ln_fwd(x_fin, self.w("fin.ln.w"), self.w("fin.ln.b"), None, a["meanc"], a["rstdc"], rows_per_group=B, groups=1,
                       C=C, x_row_stride=N * C, x_gs=T * C, eps=self.eps, y_f32=a["clsn32"])
            ops.bce_head_fwd(a["clsn32"], self.w("head.w"), self.w("head.b"), labels, a["logits"], a["loss"], B=B, C_=C)
            self._labels = labels
            self.saved_valid = train
            return a["logits"], a["loss"]
        # ---- final norm on the CLS rows only (row 0 is all the reference consumes), heads, loss
        ops.ln_fwd(x_fin, self.w("fin.ln.w"), self.w("fin.ln.b"), a["clsn"], a["meanc"], a["rstdc"], rows_per_group=B,
                   groups=G, C=C, x_row_stride=N * C, x_gs=T * C)
        self._fwd(a["clsn"], self.wb("head.w1"), a["hh"], G=G, T=B, N=F, K=C, epi=EPI_BIAS_GELU, bias=self.w("head.b1"),
                  aux=a["uh"])
        if drop:
            self._dropout(ops.DROP_BF16, a["hh"], None, a["hh"], self.SITE_HEAD_GELU)
        ops.head_loss_fwd(a["hh"], self.w("head.w2"), self.w("head.b2"), labels, a["logits"], a["loss"], M=G, B=B, F=F,
                          classes=self.classes, smoothing=self.smoothing, p_drop=self.p_drop if drop else 0.0,
                          seed=self.seed_buf, site=self.SITE_HEAD_LOGITS)
        self._labels = labels
        self.saved_valid = train
        return a["logits"], a["loss"]

    def _fusion_fwd(self, mb: int, X: torch.Tensor, train: bool):
        """CrossAttentionBlock over all K fusions of multi-block `mb`; rewrites the CLS rows of the
        receiving streams in place (outs[i] = cat(new_cls, attn[i][:, 1:]), model_cross.py:142)."""
        a, N, C, F, H, T, K, B = self.a, self.N, self.C, self.F, self.H, self.T, self.K, self.B
        s = mb if train else 0
        tag = f"X{mb}"
        f_cls = a["f_cls"][s]
        # save the CLS rows the fusions read (they are overwritten below): one launch over the K donor streams
        ops.gather_rows_f32_indexed(X, f_cls, rows=B, C_=C, src_row_stride=N * C, src_gs=T * C, src_groups=self.cls_src,
                                    dst_row_stride=C, dst_gs=B * C)
        drop = self.drop
        if self.fold:
            HC = H * C
            # query from the normalised CLS row, then q'_h = Wk_h^T q_h against the head-expanded key weight
            ops.ln_fwd(f_cls, self.w(f"{tag}.lnA.w"), self.w(f"{tag}.lnA.b"), a["f_xncls"][s], a["f_mean0"][s], a["f_rstd0"][s],
                       rows_per_group=B, groups=K, C=C)
            self._fwd(a["f_xncls"][s], self.wb(f"{tag}.wq"), a["f_qb"][s], G=K, T=B, N=C, K=C, epi=EPI_BIAS,
                      bias=self.w(f"{tag}.bq"))
            Ekv = a["f_Ekv"][s]
            ops.expand_heads(self.wb(f"{tag}.wkv"), Ekv, groups=2 * K, C_=C, H=H)
            ops.gemm(a["f_qb"][s], Ekv, a["f_qp"][s], M=B, N=HC, K=C, groups=K, b_mn=True, lda=C, ldb=HC, ldo=HC,
                     a_gs=B * C, b_gs=2 * C * HC, out_gs=B * HC)
            ops.xfold_fwd(X, f_cls, a["f_qp"][s], self.w(f"{tag}.lnA.w"), self.w(f"{tag}.lnA.b"), a["f_zhat"][s], a["f_zb"][s],
                          a["f_probs"][s], a["f_mean"][s], a["f_rstd"][s], a["xf_scratch"], K=K, B=B, N=N, C_=C, H=H,
                          cls_src=self.cls_src, tok_src=self.tok_src, scale=self.scale)
            # o = Wv z + bv on the head-expanded value weight (second half of each fusion's Ekv / bkv)
            ops.gemm(a["f_zb"][s], Ekv[:, 1], a["f_xo"][s], M=B, N=C, K=HC, groups=K, lda=HC, ldb=HC, ldo=C, a_gs=B * HC,
                     b_gs=2 * C * HC, out_gs=B * C, epi=EPI_BIAS, bias=self.w(f"{tag}.bkv")[:, C:], bias_gs=2 * C)
        else:
            ops.ln_fusion_fwd(X, f_cls, self.w(f"{tag}.lnA.w"), self.w(f"{tag}.lnA.b"), a["f_xn"][s], a["f_mean"][s],
                              a["f_rstd"][s], B=B, N=N, C_=C, cls_src=self.cls_src, tok_src=self.tok_src)
            # K | V projection of every token, one [T, 2C] GEMM per fusion (grouped)
            self._fwd(a["f_xn"][s], self.wb(f"{tag}.wkv"), a["f_kv"][s], G=K, T=T, N=2 * C, K=C, epi=EPI_BIAS,
                      bias=self.w(f"{tag}.bkv"))
            # query from the CLS row only: rows b*N of xn (row stride N*C)
            self._fwd(a["f_xn"][s], self.wb(f"{tag}.wq"), a["f_q"][s], G=K, T=B, N=C, K=C, epi=EPI_BIAS,
                      bias=self.w(f"{tag}.bq"), lda=N * C, a_gs=T * C)
            ops.xattn_fwd(a["f_q"][s], a["f_kv"][s], a["f_xo"][s], a["f_probs"][s], K=K, B=B, N=N, H=H, scale=self.scale,
                          p_drop=self.p_drop if drop else 0.0, seed=self.seed_buf, site=self.site_fusion(mb, "attn"))
        ops.cast_bf16(a["f_xo"][s], a["f_xob"][s])
        if drop:
            self._fwd(a["f_xob"][s], self.wb(f"{tag}.wp"), a["br_f"], G=K, T=B, N=C, K=C, epi=EPI_BIAS, bias=self.w(f"{tag}.bp"))
            self._dropout(ops.DROP_ADD, f_cls, a["br_f"], a["f_y"][s], self.site_fusion(mb, "proj"))
        else:
            self._fwd(a["f_xob"][s], self.wb(f"{tag}.wp"), a["f_y"][s], G=K, T=B, N=C, K=C, epi=EPI_BIAS_RESID,
                      bias=self.w(f"{tag}.bp"), resid=f_cls)
        ops.ln_fwd(a["f_y"][s], self.w(f"{tag}.lnF.w"), self.w(f"{tag}.lnF.b"), a["f_yn"][s], a["f_meany"][s],
                   a["f_rstdy"][s], rows_per_group=B, groups=K, C=C)
        self._fwd(a["f_yn"][s], self.wb(f"{tag}.w1"), a["f_h"][s], G=K, T=B, N=F, K=C, epi=EPI_BIAS_GELU,
                  bias=self.w(f"{tag}.b1"), aux=a["f_u"][s])
        if drop:
            self._dropout(ops.DROP_BF16, a["f_h"][s], None, a["f_h"][s], self.site_fusion(mb, "gelu"))
            self._fwd(a["f_h"][s], self.wb(f"{tag}.w2"), a["br_f"], G=K, T=B, N=C, K=F, epi=EPI_BIAS, bias=self.w(f"{tag}.b2"))
            self._dropout(ops.DROP_ADD, a["f_y"][s], a["br_f"], a["f_z"][s], self.site_fusion(mb, "fc2"))
        else:
            self._fwd(a["f_h"][s], self.wb(f"{tag}.w2"), a["f_z"][s], G=K, T=B, N=C, K=F, epi=EPI_BIAS_RESID,
                      bias=self.w(f"{tag}.b2"), resid=a["f_y"][s])
        ops.gather_rows_f32_indexed(a["f_z"][s], X, rows=B, C_=C, src_row_stride=C, src_gs=B * C, dst_row_stride=N * C,
                                    dst_gs=T * C, dst_groups=self.cls_src)

    # ------------------------------------------------------------------ backward
    def drop_graphs(self):
        """Forget every captured CUDA graph (current and parked plans); the next steps run eagerly and re-capture."""
        import gc
        self._fwd_graphs.clear()
        self._bwd_graphs.clear()
        for key, (a, _f, _b, fold) in list(self._plans.items()):
            self._plans[key] = (a, {}, {}, fold)
        gc.collect()

    def _next_grad_buffer(self):
        """Gradient buffer for this backward. If a parameter's .grad still aliases the candidate
        (gradient accumulation without zero_grad), switch to another buffer so nothing is clobbered."""
        def aliased(buf):
            lo, hi = buf.data_ptr(), buf.data_ptr() + 4 * buf.numel()
            for p in self.params.values():
                if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                    return True
            return False
        for i, buf in enumerate(self.grad_bufs):
            if not aliased(buf):
                self._grad_idx = i
                return buf
        self.grad_bufs.append(torch.zeros(self.layout.total, dtype=F32, device=self.device))
        self._grad_idx = len(self.grad_bufs) - 1
        return self.grad_bufs[-1]

    def backward(self, loss_scale: float = 1.0, on_range_done=None, loss_scale_dev: Optional[torch.Tensor] = None):
        """Backward of the last training forward. Fills the flat gradient buffer and returns it.
        `on_range_done(tag, start, end)` is called right after the kernels producing the gradients
        of flat range [start, end) have been enqueued (used to overlap the DP all-reduce)."""
        if not self.saved_valid:
            raise _abi.CavitError("backward() without a preceding training forward()")
        self.saved_valid = False
        with torch.cuda.device(self.device):
            self.grad = self._next_grad_buffer()
            out = self._backward_dispatch(loss_scale, on_range_done, loss_scale_dev)
            if self.post_backward is not None:   # e.g. cavit.ddp's whole-buffer gradient all-reduce
                self.post_backward(out)
        return out

    def _backward_dispatch(self, loss_scale, on_range_done, loss_scale_dev):
        # A range hook that only ENQUEUES stream work (cavit.ddp: event record, NCCL all-reduce on a side stream, stream
        # wait) is recorded into the backward graph like the kernels around it (`hook_capturable`); any other hook
        # forces eager launches.
        hook_ok = on_range_done is None or (self.hook_capturable and not self._hook_capture_failed)
        if not (self.use_graphs and ops.PROFILE is None and hook_ok):
            return self._backward_impl(loss_scale, on_range_done, loss_scale_dev)
        # (the fused embedding wgrad re-reads the forward's input volumes: a graph is tied to the buffer it recorded)
        bkey = (self.B, self._grad_idx, self.drop, on_range_done is not None,
                self._img.data_ptr() if (self.embed_fused and self._img is not None) else 0)
        st = self._bwd_graphs.setdefault(bkey, {"runs": 0, "graph": None})
        if st["graph"] is None:
            st["runs"] += 1
            if st["runs"] <= 2:
                return self._backward_impl(loss_scale, on_range_done, loss_scale_dev)
            st["scale"] = torch.ones(1, dtype=F32, device=self.device)
            st["labels"] = self._labels
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            n0 = _abi.launch_count()
            if on_range_done is None:
                with torch.cuda.graph(graph):
                    self._backward_impl(1.0, None, st["scale"])
            else:
                try:   # collectives inside the capture: thread-local capture mode keeps NCCL's watchdog thread out of it
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        self._backward_impl(1.0, on_range_done, st["scale"])
                except Exception as exc:   # this stack cannot capture the hook's work: stay eager from now on (and say so)
                    self._hook_capture_failed = True
                    self.hook_capture_error = repr(exc)
                    del self._bwd_graphs[bkey]
                    torch.cuda.synchronize(self.device)
                    return self._backward_impl(loss_scale, on_range_done, loss_scale_dev)
            st["launches"] = _abi.launch_count() - n0
            st["graph"] = graph
        if st["labels"].data_ptr() != self._labels.data_ptr():
            st["labels"].copy_(self._labels)
        if loss_scale_dev is not None:
            st["scale"].copy_(loss_scale_dev)
            if loss_scale != 1.0:
                st["scale"].mul_(loss_scale)
        else:
            st["scale"].fill_(loss_scale)
        st["graph"].replay()
        self.graph_launches += st["launches"]
        return self.grad

    def _backward_impl(self, loss_scale: float, on_range_done, loss_scale_dev: Optional[torch.Tensor]):
        cfg = self.cfg
        a, G, N, C, F, H, T, K, B = self.a, self.G, self.N, self.C, self.F, self.H, self.T, self.K, self.B
        ranges = {t: (s, e_) for t, s, e_ in self.layout.layer_ranges}

        def done(tag):
            if on_range_done is not None:
                s, e_ = ranges[tag]
                on_range_done(tag, s, e_)

        if self.split:
            from . import engine_fp32
            return engine_fp32.backward_impl(self, loss_scale, done, loss_scale_dev)
        if self.post_norm:
            return self._backward_post(loss_scale, done, loss_scale_dev)
        ws = a["ln_ws"]
        if self.kind == "cnnvit":
            return self._backward_cnnvit(loss_scale, done, loss_scale_dev)
        # ---- loss, heads, final norm
        ops.head_loss_bwd(a["hh"], self.w("head.w2"), self._labels, a["logits"], a["dhh"], self.g("head.w2"),
                          self.g("head.b2"), M=G, B=B, F=F, classes=self.classes, smoothing=self.smoothing,
                          loss_scale=loss_scale, loss_scale_dev=loss_scale_dev, p_drop=self.p_drop if self.drop else 0.0,
                          seed=self.seed_buf, site=self.SITE_HEAD_LOGITS)
        drop = self.drop
        ops.gelu_bwd_bf16(a["dhh"], a["uh"], a["duh"])
        if drop:
            self._dropout(ops.DROP_BF16, a["duh"], None, a["duh"], self.SITE_HEAD_GELU)
        self._dgrad(a["duh"], self.wb("head.w1"), a["dclsn"], G=G, T=B, N=F, K=C)
        self._wgrad(a["duh"], a["clsn"], self.g("head.w1"), G=G, T=B, N=F, K=C)
        self._colsum(a["duh"], self.g("head.b1"), G=G, T=B, N=F)
        dX, dXb = a["dX"], a["dXb"]
        dX.zero_()
        ops.ln_bwd(a["dclsn"], self._x_fin, a["meanc"], a["rstdc"], self.w("fin.ln.w"), dX, self.g("fin.ln.w"),
                   self.g("fin.ln.b"), ws, rows_per_group=B, groups=G, C=C, x_row_stride=N * C, x_gs=T * C,
                   dx_row_stride=N * C, dx_gs=T * C)
        done("head")
        self._backward_layers(dX, dXb, done)
        # ---- embedding: d(pos), d(cls), dW = dY^T unfold(x), db
        if drop:   # through the embedding dropout; the bf16 copy is rebuilt from the masked gradient
            self._dropout(ops.DROP_F32, dX, None, dX, self.SITE_EMBED)
            ops.cast_bf16(dX, dXb)
        ops.embed_param_grads(dX, self.g("pos"), self.g("cls"), M=G, B=B, N=N, C_=C)
        np_seq = self.N - 1
        if self.embed_fused:    # dW = dY^T unfold(x) with the unfold redone by TMA; db = sum of the patch rows of d(pos)
            ops.embed_fused_wgrad(self._img, dXb, self.g("embed.w"), patch_size=cfg.patch_size, C_=C,
                                  sample_major=(self.kind == "vit"))
            ops.embed_bias_grad(self.g("pos"), self.g("embed.b"), N=N, C_=C)
        else:
            ops.compact_patch_rows_bf16(dXb, a["dcomp"], S=G * B, Np=np_seq, C_=C)
            R = self.Mimg * B * self.Np
            self._wgrad(a["dcomp"], a["patches"], self.g("embed.w"), G=1, T=R, N=C, K=self.P)
            self._colsum(a["dcomp"], self.g("embed.b"), G=1, T=R, N=C)
        done("embed")
        return self.grad

    def _backward_layers(self, dX, dXb, done):
        """Backward through the pre-norm blocks L-1 .. 0 (and the fusions between them); on return dX / dXb hold the
        gradient of the token streams after the embedding."""
        cfg = self.cfg
        a, G, N, C, F, H, T, K, B = self.a, self.G, self.N, self.C, self.F, self.H, self.T, self.K, self.B
        ws = a["ln_ws"]
        drop = self.drop
        need_cast = True  # dXb must mirror dX before the first layer's GEMMs
        b2_done = False   # fc2 bias gradient of this layer already produced by the LayerNorm backward above it
        for l in reversed(range(self.L)):
            tag = f"L{l}"
            if self.kind == "cross" and K and (l + 1) % cfg.num_self_blocks == 0:
                self._fusion_bwd(l // cfg.num_self_blocks, dX)
                done(f"X{l // cfg.num_self_blocks}")
                need_cast = True
            if drop:     # gradient of the dropped fc2 branch: mask_fc2 * dX / (1 - p)
                self._dropout(ops.DROP_CAST, dX, None, dXb, self.site_layer(l, "fc2"))
            elif need_cast:
                ops.cast_bf16(dX, dXb)
            need_cast = False
            x_in, x_mid = a["X"][2 * l], a["X"][2 * l + 1]
            # FFN: x_out = x_mid + W2 gelu(W1 LN2(x_mid) + b1) + b2
            self._dgrad(dXb, self.wb(f"{tag}.w2"), a["dbig"], G=G, T=T, N=C, K=F, epi=EPI_GELU_BWD, aux=a["u"][l])
            if drop:
                self._dropout(ops.DROP_BF16, a["dbig"], None, a["dbig"], self.site_layer(l, "gelu"))
            self._wgrad(dXb, a["h"][l], self.g(f"{tag}.w2"), G=G, T=T, N=C, K=F)
            if not b2_done:
                self._colsum(dXb, self.g(f"{tag}.b2"), G=G, T=T, N=C)
            self._dgrad(a["dbig"], self.wb(f"{tag}.w1"), a["dmid"], G=G, T=T, N=F, K=C)
            self._wgrad(a["dbig"], a["xn2"][l], self.g(f"{tag}.w1"), G=G, T=T, N=F, K=C)
            self._colsum(a["dbig"], self.g(f"{tag}.b1"), G=G, T=T, N=F)
            # the column sums of this LayerNorm backward's output are the out-proj bias gradient (no pass over dXb)
            bo_fused = (H != 1) and not drop
            ops.ln_bwd(a["dmid"], x_mid, a["mean2"][l], a["rstd2"][l], self.w(f"{tag}.ln2.w"), dX, self.g(f"{tag}.ln2.w"),
                       self.g(f"{tag}.ln2.b"), ws, rows_per_group=T, groups=G, C=C, dresid=dX,
                       dx_bf16=None if drop else dXb, dcol=self.g(f"{tag}.bo") if bo_fused else None)
            if drop:
                if H != 1:
                    self._dropout(ops.DROP_CAST, dX, None, dXb, self.site_layer(l, "out"))
                else:
                    ops.cast_bf16(dX, dXb)
            # attention: x_mid = x_in + Wo attn(LN1(x_in)) + bo
            if H != 1:
                self._dgrad(dXb, self.wb(f"{tag}.wo"), a["dmid"], G=G, T=T, N=C, K=C)
                self._wgrad(dXb, a["ao"][l], self.g(f"{tag}.wo"), G=G, T=T, N=C, K=C)
                if not bo_fused:
                    self._colsum(dXb, self.g(f"{tag}.bo"), G=G, T=T, N=C)
                dao = a["dmid"]
            else:
                dao = dXb
            ops.attn_bwd(a["qkv"][l], a["ao"][l], dao, a["lse"][l], a["dqkv"], a["delta"], a["dq_acc"], G=G, B=B, N=N,
                         H=H, scale=self.scale)
            self._dgrad(a["dqkv"], self.wb(f"{tag}.wqkv"), a["dmid"], G=G, T=T, N=3 * C, K=C)
            self._wgrad(a["dqkv"], a["xn1"][l], self.g(f"{tag}.wqkv"), G=G, T=T, N=3 * C, K=C)
            if self.qkv_bias:
                self._colsum(a["dqkv"], self.g(f"{tag}.bqkv"), G=G, T=T, N=3 * C)
            # ... and of the layer below's fc2 bias, unless a fusion backward rewrites dX in between
            fusion_next = self.kind == "cross" and K and l % cfg.num_self_blocks == 0
            b2_done = l >= 1 and not drop and not fusion_next
            ops.ln_bwd(a["dmid"], x_in, a["mean1"][l], a["rstd1"][l], self.w(f"{tag}.ln1.w"), dX, self.g(f"{tag}.ln1.w"),
                       self.g(f"{tag}.ln1.b"), ws, rows_per_group=T, groups=G, C=C, dresid=dX,
                       dx_bf16=None if drop else dXb, dcol=self.g(f"L{l - 1}.b2") if b2_done else None)
            done(tag)

    # ------------------------------------------------------------------ ViT (model.py) tail / embedding adjoint
    def _backward_cnnvit(self, loss_scale, done, loss_scale_dev):
        """Backward of the `ViT` core (/root/reference/model.py:253-286): BCE tail, encoder_norm on the CLS rows,
        the pre-norm blocks, the Conv3d(k = stride) patch embedding down to d(stem feature maps)."""
        cfg = self.cfg
        a, N, C, T, B = self.a, self.N, self.C, self.T, self.B
        ws = a["ln_ws"]
        ops.bce_head_bwd(a["clsn32"], self.w("head.w"), self._labels, a["logits"], a["dclsn32"], self.g("head.w"),
                         self.g("head.b"), B=B, C_=C, loss_scale=loss_scale, loss_scale_dev=loss_scale_dev)
        dX, dXb = a["dX"], a["dXb"]
        dX.zero_()
        ops.ln_bwd(a["dclsn32"], self._x_fin, a["meanc"], a["rstdc"], self.w("fin.ln.w"), dX, self.g("fin.ln.w"),
                   self.g("fin.ln.b"), ws, rows_per_group=B, groups=1, C=C, x_row_stride=N * C, x_gs=T * C,
                   dx_row_stride=N * C, dx_gs=T * C)
        done("head")
        self._backward_layers(dX, dXb, done)
        # positional table: row 0 once, rows 1.. summed over the modalities that shared them (model.py:89,103,258)
        Np, M = self.Np, self.Mimg
        ops.embed_param_grads(dX, a["dpos_exp"], self.g("cls"), M=1, B=B, N=N, C_=C)
        gpos = self.g("pos")
        ops.gather_rows_f32(a["dpos_exp"], gpos, rows=Np + 1, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=C,
                            dst_gs=0)
        for m in range(1, M):
            ops.gather_rows_f32(a["dpos_exp"][1 + m * Np:], gpos[1:], rows=Np, C_=C, groups=1, src_row_stride=C, src_gs=0,
                                dst_row_stride=C, dst_gs=0, accumulate=True)
        ops.compact_patch_rows_bf16(dXb, a["dcomp"], S=B, Np=N - 1, C_=C)
        R = M * B * Np
        self._wgrad(a["dcomp"], a["patches"], self.g("embed.w"), G=1, T=R, N=C, K=self.P)
        self._colsum(a["dcomp"], self.g("embed.b"), G=1, T=R, N=C)
        # d(feature maps) = unpatchify(dY W): the stem is trainable, so the path does not end at the embedding
        self._dgrad(a["dcomp"], self.wb("embed.w"), a["drows"], G=1, T=R, N=C, K=self.P)
        ops.conv_patch_rows_bwd(a["drows"], a["dfeat"], M=M, B=B, Cin=cfg.in_channels, dims=cfg.feat_dims, grid=cfg.grid)
        self.dinput = a["dfeat"]
        done("embed")
        return self.grad

    # ------------------------------------------------------------------ ViT3D (modelv2.py): post-norm encoder
    def _plan_post(self, a, B, train):
        dev = self.device
        N, C, F, H, L = self.N, self.C, self.F, self.H, self.L
        T = B * N
        Fh = self.cfg.head_dim_hidden

        def e(shape, dt=F32):
            return torch.empty(shape, dtype=dt, device=dev)

        nL = L if train else 1
        a["x0"], a["xa"], a["xb"] = e((1, T, C)), e((1, T, C)), e((1, T, C))
        a["xin_b"] = [e((1, T, C), BF16) for _ in range(nL + 1 if train else 2)]   # bf16 layer inputs (wgrad operands)
        for nm, shp, dt in [("qkv", (1, T, 3 * C), BF16), ("ao", (1, T, C), BF16), ("lse", (1, B, H, N), F32),
                            ("s1", (1, T, C), F32), ("mean1", (1, T), F32), ("rstd1", (1, T), F32), ("x1b", (1, T, C), BF16),
                            ("h", (1, T, F), BF16), ("s2", (1, T, C), F32), ("mean2", (1, T), F32), ("rstd2", (1, T), F32)]:
            a[nm] = [e(shp, dt) for _ in range(nL)]
        a["zeros"] = torch.zeros(max(3 * C, F), dtype=F32, device=dev)
        a["clsn"], a["meanc"], a["rstdc"] = e((1, B, C), BF16), e((1, B)), e((1, B))
        if not self.has_cls:
            a["pool"] = e((1, B, C))
            a["dcls_dummy"] = e((C,))
            if train:
                a["dpool"] = e((1, B, C))
        a["hh"] = e((1, B, Fh), BF16)
        a["logits"], a["loss"] = e((B, self.classes)), e((1,))
        if train:
            a["dA"], a["dB"], a["dXb"] = e((1, T, C)), e((1, T, C)), e((1, T, C), BF16)
            a["dbig"], a["dmid"], a["dqkv"] = e((1, T, F), BF16), e((1, T, C), BF16), e((1, T, 3 * C), BF16)
            a["delta"], a["dq_acc"] = e((1, B, H, N)), e((1, T, C))
            a["ln_ws"] = ops.ln_bwd_workspace(1, C, dev)
            a["dhh"], a["dclsn"] = e((1, B, Fh), BF16), e((1, B, C), BF16)
            a["dfeat"] = e((B, C, N - int(self.has_cls)))

    def _forward_post(self, feat, labels, train):
        """`ViT3D.forward` from the stem features on (/root/reference/modelv2.py:203-241): token assembly, L post-norm
        nn.TransformerEncoderLayer blocks  x = LN1(x + SA(x)); x = LN2(x + W2 relu(W1 x + b1) + b2), head on the CLS row."""
        a, N, C, F, H, T, B = self.a, self.N, self.C, self.F, self.H, self.T, self.B
        hc = self.has_cls
        ops.tokens_from_channels(feat, self.w("cls") if hc else None, self.w("pos"), a["x0"], B=B, C_=C, S=N - int(hc), has_cls=hc)
        ops.cast_bf16(a["x0"], a["xin_b"][0])
        x_in = a["x0"]
        for l in range(self.L):
            s = l if train else 0
            tag = f"L{l}"
            xin_b = a["xin_b"][l if train else l % 2]
            xout_b = a["xin_b"][l + 1 if train else (l + 1) % 2]
            self._fwd(xin_b, self.wb(f"{tag}.wqkv"), a["qkv"][s], G=1, T=T, N=3 * C, K=C, epi=EPI_BIAS,
                      bias=self.w(f"{tag}.bqkv"))
            ops.attn_fwd(a["qkv"][s], a["ao"][s], a["lse"][s], G=1, B=B, N=N, H=H, scale=self.scale)
            self._fwd(a["ao"][s], self.wb(f"{tag}.wo"), a["s1"][s], G=1, T=T, N=C, K=C, epi=EPI_BIAS_RESID,
                      bias=self.w(f"{tag}.bo"), resid=x_in)
            ops.ln_fwd(a["s1"][s], self.w(f"{tag}.ln1.w"), self.w(f"{tag}.ln1.b"), a["x1b"][s], a["mean1"][s], a["rstd1"][s],
                       rows_per_group=T, groups=1, C=C, eps=self.eps, y_f32=a["xa"])
            self._fwd(a["x1b"][s], self.wb(f"{tag}.w1"), a["h"][s], G=1, T=T, N=F, K=C, epi=EPI_BIAS_RELU,
                      bias=self.w(f"{tag}.b1"))
            self._fwd(a["h"][s], self.wb(f"{tag}.w2"), a["s2"][s], G=1, T=T, N=C, K=F, epi=EPI_BIAS_RESID,
                      bias=self.w(f"{tag}.b2"), resid=a["xa"])
            ops.ln_fwd(a["s2"][s], self.w(f"{tag}.ln2.w"), self.w(f"{tag}.ln2.b"), xout_b, a["mean2"][s], a["rstd2"][s],
                       rows_per_group=T, groups=1, C=C, eps=self.eps, y_f32=a["xb"])
            x_in = a["xb"]
        self._x_fin = x_in
        Fh = self.cfg.head_dim_hidden
        if hc:   # head on the CLS row (row 0 of every sample)
            ops.ln_fwd(x_in, self.w("fin.ln.w"), self.w("fin.ln.b"), a["clsn"], a["meanc"], a["rstdc"], rows_per_group=B,
                       groups=1, C=C, x_row_stride=N * C, x_gs=T * C, eps=1e-5)
        else:    # add_cls_token=False: head on the mean over all tokens (modelv2.py:233-235)
            ops.token_mean_fwd(x_in, a["pool"], B=B, N=N, C_=C)
            ops.ln_fwd(a["pool"], self.w("fin.ln.w"), self.w("fin.ln.b"), a["clsn"], a["meanc"], a["rstdc"], rows_per_group=B,
                       groups=1, C=C, eps=1e-5)
        self._fwd(a["clsn"], self.wb("head.w1"), a["hh"], G=1, T=B, N=Fh, K=C, epi=EPI_BIAS, bias=self.w("head.b1"))
        ops.head_loss_fwd(a["hh"], self.w("head.w2"), self.w("head.b2"), labels, a["logits"], a["loss"], M=1, B=B, F=Fh,
                          classes=self.classes, smoothing=self.smoothing)
        self._labels = labels
        self.saved_valid = train
        return a["logits"], a["loss"]

    def _backward_post(self, loss_scale, done, loss_scale_dev):
        a, N, C, F, H, T, B = self.a, self.N, self.C, self.F, self.H, self.T, self.B
        ws, zeros = a["ln_ws"], a["zeros"]
        Fh = self.cfg.head_dim_hidden
        ops.head_loss_bwd(a["hh"], self.w("head.w2"), self._labels, a["logits"], a["dhh"], self.g("head.w2"),
                          self.g("head.b2"), M=1, B=B, F=Fh, classes=self.classes, smoothing=self.smoothing,
                          loss_scale=loss_scale, loss_scale_dev=loss_scale_dev)
        self._dgrad(a["dhh"], self.wb("head.w1"), a["dclsn"], G=1, T=B, N=Fh, K=C)
        self._wgrad(a["dhh"], a["clsn"], self.g("head.w1"), G=1, T=B, N=Fh, K=C)
        self._colsum(a["dhh"], self.g("head.b1"), G=1, T=B, N=Fh)
        dA, dB, dXb = a["dA"], a["dB"], a["dXb"]
        if self.has_cls:
            dA.zero_()
            ops.ln_bwd(a["dclsn"], self._x_fin, a["meanc"], a["rstdc"], self.w("fin.ln.w"), dA, self.g("fin.ln.w"),
                       self.g("fin.ln.b"), ws, rows_per_group=B, groups=1, C=C, x_row_stride=N * C, x_gs=T * C,
                       dx_row_stride=N * C, dx_gs=T * C)
        else:
            ops.ln_bwd(a["dclsn"], a["pool"], a["meanc"], a["rstdc"], self.w("fin.ln.w"), a["dpool"], self.g("fin.ln.w"),
                       self.g("fin.ln.b"), ws, rows_per_group=B, groups=1, C=C)
            ops.token_mean_bwd(a["dpool"], dA, B=B, N=N, C_=C)
        done("head")
        for l in reversed(range(self.L)):
            tag = f"L{l}"
            # x2 = LN2(s2): dA = d(x2) (fp32) -> dB = d(s2), dXb its bf16 copy, column sums = d(b2)
            ops.ln_bwd(dA, a["s2"][l], a["mean2"][l], a["rstd2"][l], self.w(f"{tag}.ln2.w"), dB, self.g(f"{tag}.ln2.w"),
                       self.g(f"{tag}.ln2.b"), ws, rows_per_group=T, groups=1, C=C, dx_bf16=dXb, dcol=self.g(f"{tag}.b2"))
            self._dgrad(dXb, self.wb(f"{tag}.w2"), a["dbig"], G=1, T=T, N=C, K=F, epi=EPI_RELU_BWD, aux=a["h"][l])
            self._wgrad(dXb, a["h"][l], self.g(f"{tag}.w2"), G=1, T=T, N=C, K=F)
            # d(x1) = d(s2) + d(u) W1  (residual add fused into the dgrad epilogue, fp32)
            self._dgrad(a["dbig"], self.wb(f"{tag}.w1"), dA, G=1, T=T, N=F, K=C, epi=EPI_BIAS_RESID, bias=zeros, resid=dB)
            self._wgrad(a["dbig"], a["x1b"][l], self.g(f"{tag}.w1"), G=1, T=T, N=F, K=C)
            self._colsum(a["dbig"], self.g(f"{tag}.b1"), G=1, T=T, N=F)
            # x1 = LN1(s1): dA = d(x1) -> dB = d(s1), column sums = d(out_proj.bias)
            ops.ln_bwd(dA, a["s1"][l], a["mean1"][l], a["rstd1"][l], self.w(f"{tag}.ln1.w"), dB, self.g(f"{tag}.ln1.w"),
                       self.g(f"{tag}.ln1.b"), ws, rows_per_group=T, groups=1, C=C, dx_bf16=dXb, dcol=self.g(f"{tag}.bo"))
            self._dgrad(dXb, self.wb(f"{tag}.wo"), a["dmid"], G=1, T=T, N=C, K=C)
            self._wgrad(dXb, a["ao"][l], self.g(f"{tag}.wo"), G=1, T=T, N=C, K=C)
            ops.attn_bwd(a["qkv"][l], a["ao"][l], a["dmid"], a["lse"][l], a["dqkv"], a["delta"], a["dq_acc"], G=1, B=B, N=N,
                         H=H, scale=self.scale)
            self._wgrad(a["dqkv"], a["xin_b"][l], self.g(f"{tag}.wqkv"), G=1, T=T, N=3 * C, K=C)
            self._colsum(a["dqkv"], self.g(f"{tag}.bqkv"), G=1, T=T, N=3 * C)
            # d(x_in) = d(s1) + d(qkv) Wqkv
            self._dgrad(a["dqkv"], self.wb(f"{tag}.wqkv"), dA, G=1, T=T, N=3 * C, K=C, epi=EPI_BIAS_RESID, bias=zeros, resid=dB)
            done(tag)
        ops.embed_param_grads(dA, self.g("pos"), self.g("cls") if self.has_cls else a["dcls_dummy"], M=1, B=B, N=N, C_=C)
        ops.tokens_to_channels(dA, a["dfeat"], B=B, C_=C, S=N - int(self.has_cls), has_cls=self.has_cls)
        self.dinput = a["dfeat"]
        done("embed")
        return self.grad

    def _fusion_bwd(self, mb: int, dX: torch.Tensor):
        a, N, C, F, H, T, K, B = self.a, self.N, self.C, self.F, self.H, self.T, self.K, self.B
        tag = f"X{mb}"
        ws = a["ln_ws"]
        s = mb
        d_z = a["d_z"]
        # gradient of the new CLS rows; the old CLS rows of receiving streams get no pass-through gradient
        ops.gather_rows_f32_indexed(dX, d_z, rows=B, C_=C, src_row_stride=N * C, src_gs=T * C, src_groups=self.cls_src,
                                    dst_row_stride=C, dst_gs=B * C, zero_src=True)
        drop = self.drop
        if drop:
            self._dropout(ops.DROP_CAST, d_z, None, a["d_zb"], self.site_fusion(mb, "fc2"))
        else:
            ops.cast_bf16(d_z, a["d_zb"])
        # FFN on the single CLS token
        self._dgrad(a["d_zb"], self.wb(f"{tag}.w2"), a["d_u"], G=K, T=B, N=C, K=F, epi=EPI_GELU_BWD, aux=a["f_u"][s])
        if drop:
            self._dropout(ops.DROP_BF16, a["d_u"], None, a["d_u"], self.site_fusion(mb, "gelu"))
        self._wgrad(a["d_zb"], a["f_h"][s], self.g(f"{tag}.w2"), G=K, T=B, N=C, K=F)
        self._colsum(a["d_zb"], self.g(f"{tag}.b2"), G=K, T=B, N=C)
        self._dgrad(a["d_u"], self.wb(f"{tag}.w1"), a["d_yn"], G=K, T=B, N=F, K=C)
        self._wgrad(a["d_u"], a["f_yn"][s], self.g(f"{tag}.w1"), G=K, T=B, N=F, K=C)
        self._colsum(a["d_u"], self.g(f"{tag}.b1"), G=K, T=B, N=F)
        ops.ln_bwd(a["d_yn"], a["f_y"][s], a["f_meany"][s], a["f_rstdy"][s], self.w(f"{tag}.lnF.w"), a["d_y"],
                   self.g(f"{tag}.lnF.w"), self.g(f"{tag}.lnF.b"), ws, rows_per_group=B, groups=K, C=C, dresid=d_z,
                   dx_bf16=None if drop else a["d_yb"])
        if drop:
            self._dropout(ops.DROP_CAST, a["d_y"], None, a["d_yb"], self.site_fusion(mb, "proj"))
        # y = proj(xattn) + cls_in
        if self.fold:
            self._fusion_bwd_folded(mb, dX)
        else:
            self._dgrad(a["d_yb"], self.wb(f"{tag}.wp"), a["d_xo"], G=K, T=B, N=C, K=C)
            self._wgrad(a["d_yb"], a["f_xob"][s], self.g(f"{tag}.wp"), G=K, T=B, N=C, K=C)
            self._colsum(a["d_yb"], self.g(f"{tag}.bp"), G=K, T=B, N=C)
            ops.xattn_bwd(a["f_q"][s], a["f_kv"][s], a["f_probs"][s], a["d_xo"], a["d_q"], a["d_kv"], K=K, B=B, N=N, H=H,
                          scale=self.scale, p_drop=self.p_drop if drop else 0.0, seed=self.seed_buf,
                          site=self.site_fusion(mb, "attn"))
            # query path (CLS row of xn only)
            ops.cast_bf16(a["d_q"], a["d_qb"])
            self._dgrad(a["d_qb"], self.wb(f"{tag}.wq"), a["d_xncls"], G=K, T=B, N=C, K=C)
            self._wgrad(a["d_qb"], a["f_xn"][s], self.g(f"{tag}.wq"), G=K, T=B, N=C, K=C, ldx=N * C, x_gs=T * C)
            self._colsum(a["d_qb"], self.g(f"{tag}.bq"), G=K, T=B, N=C)
            # key / value path (all tokens)
            self._dgrad(a["d_kv"], self.wb(f"{tag}.wkv"), a["d_xn"], G=K, T=T, N=2 * C, K=C)
            self._wgrad(a["d_kv"], a["f_xn"][s], self.g(f"{tag}.wkv"), G=K, T=T, N=2 * C, K=C)
            self._colsum(a["d_kv"], self.g(f"{tag}.bkv"), G=K, T=T, N=2 * C)
            # LayerNorm of cat(cls_i, patches_j): scatter-add into the stream gradients
            ops.ln_fusion_bwd(a["d_xn"], self._x_for_fusion(mb), a["f_cls"][s], a["f_mean"][s], a["f_rstd"][s],
                              self.w(f"{tag}.lnA.w"), dX, self.g(f"{tag}.lnA.w"), self.g(f"{tag}.lnA.b"), ws, B=B, N=N, C_=C,
                              cls_src=self.cls_src, tok_src=self.tok_src, dy_cls=a["d_xncls"])
        # residual path of the CLS token: y = ... + cls_in
        ops.gather_rows_f32_indexed(a["d_y"], dX, rows=B, C_=C, src_row_stride=C, src_gs=B * C, dst_row_stride=N * C,
                                    dst_gs=T * C, dst_groups=self.cls_src, accumulate=True)

    def _fusion_bwd_folded(self, mb: int, dX: torch.Tensor):
        """Adjoint of the folded forward (see csrc/xfold.cu): everything left of the token streams is a [B, .] problem."""
        a, N, C, H, K, B = self.a, self.N, self.C, self.H, self.K, self.B
        tag, s, ws, HC = f"X{mb}", mb, self.a["ln_ws"], self.H * self.C
        Ekv, dE = a["f_Ekv"][s], a["d_Ekv"]
        gkv, gbkv = self.g(f"{tag}.wkv"), self.g(f"{tag}.bkv").view(K, 2, C)
        self._dgrad(a["d_yb"], self.wb(f"{tag}.wp"), a["d_xob"], G=K, T=B, N=C, K=C)
        self._wgrad(a["d_yb"], a["f_xob"][s], self.g(f"{tag}.wp"), G=K, T=B, N=C, K=C)
        self._colsum(a["d_yb"], self.g(f"{tag}.bp"), G=K, T=B, N=C)
        # value side: gz_h = Wv_h^T do_h, dWv (expanded), dbv = sum_b do, dbk = 0 (scores are shift invariant)
        ops.gemm(a["d_xob"], Ekv[:, 1], a["d_gz"], M=B, N=HC, K=C, groups=K, b_mn=True, lda=C, ldb=HC, ldo=HC, a_gs=B * C,
                 b_gs=2 * C * HC, out_gs=B * HC)
        ops.gemm(a["d_xob"], a["f_zb"][s], dE[:, 1], M=C, N=HC, K=B, groups=K, a_mn=True, b_mn=True, lda=C, ldb=HC, ldo=HC,
                 a_gs=B * C, b_gs=B * HC, out_gs=2 * C * HC)
        self._colsum(a["d_xob"], a["d_bv"], G=K, T=B, N=C)
        ops.gather_rows_f32(a["d_bv"], gbkv[:, 1], rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0, dst_row_stride=2 * C,
                            dst_gs=0)
        a["d_lnA"].zero_()
        ops.gather_rows_f32(a["d_lnA"][0], gbkv[:, 0], rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0,
                            dst_row_stride=2 * C, dst_gs=0)     # dbk = 0
        ops.xfold_bwd(self._x_for_fusion(mb), a["f_cls"][s], a["f_qp"][s], self.w(f"{tag}.lnA.w"), a["f_zhat"][s],
                      a["f_probs"][s], a["f_mean"][s], a["f_rstd"][s], a["d_gz"], a["xf_scratch"], dX, a["d_qp"],
                      a["d_lnA"][0], a["d_lnA"][1], K=K, B=B, N=N, C_=C, H=H, cls_src=self.cls_src, tok_src=self.tok_src,
                      scale=self.scale)
        # key side: dq_h = Wk_h dq'_h, dWk (expanded) = q^T dq'
        ops.cast_bf16(a["d_qp"], a["d_qpb"])
        ops.gemm(a["d_qpb"], Ekv, a["d_qb"], M=B, N=C, K=HC, groups=K, lda=HC, ldb=HC, ldo=C, a_gs=B * HC, b_gs=2 * C * HC,
                 out_gs=B * C)
        ops.gemm(a["f_qb"][s], a["d_qpb"], dE, M=C, N=HC, K=B, groups=K, a_mn=True, b_mn=True, lda=C, ldb=HC, ldo=HC,
                 a_gs=B * C, b_gs=B * HC, out_gs=2 * C * HC)
        ops.fold_heads(dE, gkv, groups=2 * K, C_=C, H=H)
        # query path through the LayerNorm of the CLS rows
        self._dgrad(a["d_qb"], self.wb(f"{tag}.wq"), a["d_xnclsb"], G=K, T=B, N=C, K=C)
        self._wgrad(a["d_qb"], a["f_xncls"][s], self.g(f"{tag}.wq"), G=K, T=B, N=C, K=C)
        self._colsum(a["d_qb"], self.g(f"{tag}.bq"), G=K, T=B, N=C)
        ops.ln_bwd(a["d_xnclsb"], a["f_cls"][s], a["f_mean0"][s], a["f_rstd0"][s], self.w(f"{tag}.lnA.w"), a["d_clsq"],
                   self.g(f"{tag}.lnA.w"), self.g(f"{tag}.lnA.b"), ws, rows_per_group=B, groups=K, C=C)
        ops.gather_rows_f32(a["d_lnA"][0], self.g(f"{tag}.lnA.w"), rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0,
                            dst_row_stride=C, dst_gs=0, accumulate=True)
        ops.gather_rows_f32(a["d_lnA"][1], self.g(f"{tag}.lnA.b"), rows=K, C_=C, groups=1, src_row_stride=C, src_gs=0,
                            dst_row_stride=C, dst_gs=0, accumulate=True)
        ops.gather_rows_f32_indexed(a["d_clsq"], dX, rows=B, C_=C, src_row_stride=C, src_gs=B * C, dst_row_stride=N * C,
                                    dst_gs=self.T * C, dst_groups=self.cls_src, accumulate=True)

    def _x_for_fusion(self, mb: int) -> torch.Tensor:
        # streams as the fusion saw them (patch rows are untouched by the in-place CLS rewrite)
        return self.a["X"][2 * (mb + 1) * self.cfg.num_self_blocks]
