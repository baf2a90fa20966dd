"""Drop-in module classes: same constructor argument (one config attribute bag), same module tree /
``state_dict`` keys / initialisation order, same ``forward(img, labels) -> (logits, loss)`` as the
reference's ``ModelCross`` (/root/reference/model_cross.py:152-212) and ``ModelVIT``
(/root/reference/modelv3.py:90-147) — but ``forward`` runs the sm_100a kernel path of
``cavit.engine`` instead of ATen ops.

The sub-modules below own the parameters under the reference's names (so reference checkpoints load and
`torch.manual_seed(s)` reproduces the reference's random init). The top-level models run the whole-model engine, which
fuses across them; each sub-module's own `forward` stays usable on its own and runs the same C-ABI kernels unfused
(cavit/functional.py) with the reference's semantics — fp32 tensors in and out.
"""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn

from . import _abi
from . import functional as CF
from .engine import Engine

try:  # the reference derives from lightning.LightningModule; keep that when lightning exists
    import lightning as _L
    _Base = _L.LightningModule
except Exception:  # pragma: no cover - lightning is not installed in the build image
    class _Base(nn.Module):
        def log(self, *args, **kwargs):
            return None


def _engine_only(name):
    """forward of a container whose compute only exists inside the fused whole-model path."""
    def forward(self, *a, **k):
        raise _abi.CavitError(
            f"{name} is a parameter container in cavit; call the top-level model's forward, which runs the fused sm_100a path")
    return forward


class PreNorm(nn.Module):
    """/root/reference/model_cross.py:11-17: fn(LayerNorm(x))."""

    def __init__(self, config, fn):
        super().__init__()
        self.norm = nn.LayerNorm(config.hidden_dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(CF.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps), **kwargs)


class FeedForward(nn.Module):
    """/root/reference/model_cross.py:19-31: Linear, GELU, Dropout, Linear, Dropout."""

    def __init__(self, config):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(config.hidden_dim, config.mlp_dim), nn.GELU(), nn.Dropout(config.dropout),
                                 nn.Linear(config.mlp_dim, config.hidden_dim), nn.Dropout(config.dropout))

    def forward(self, x):
        CF._no_dropout(self.training, self.net[2].p, "FeedForward")
        return CF.feed_forward(x, self.net[0].weight, self.net[0].bias, self.net[3].weight, self.net[3].bias)


class Attention(nn.Module):
    """/root/reference/model_cross.py:33-61."""

    def __init__(self, config, dim_head):
        super().__init__()
        inner = dim_head * config.num_heads
        assert inner == config.hidden_dim
        self.heads = config.num_heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(config.hidden_dim, inner * 3, bias=False)
        # the reference drops the output projection when there is a single head spanning hidden_dim
        single = config.num_heads == 1 and dim_head == config.hidden_dim
        self.to_out = nn.Identity() if single else nn.Sequential(nn.Linear(inner, config.hidden_dim),
                                                                 nn.Dropout(config.dropout))

    def forward(self, x):
        if isinstance(self.to_out, nn.Identity):
            return CF.self_attention(x, self.to_qkv.weight, None, None, self.heads)
        CF._no_dropout(self.training, self.to_out[1].p, "Attention")
        return CF.self_attention(x, self.to_qkv.weight, self.to_out[0].weight, self.to_out[0].bias, self.heads)


class SelfAttentionBlock(nn.Module):
    """/root/reference/model_cross.py:64-72."""

    def __init__(self, config):
        super().__init__()
        self.attn = PreNorm(config, Attention(config, dim_head=config.hidden_dim // config.num_heads))
        self.ffn = PreNorm(config, FeedForward(config))

    def forward(self, x):
        x = self.attn(x) + x
        x = self.ffn(x) + x
        return x


class CrossAttention(nn.Module):
    """/root/reference/model_cross.py:74-102: the query is token 0 only."""

    def __init__(self, config):
        super().__init__()
        self.num_heads = config.num_heads
        self.scale = (config.hidden_dim // config.num_heads) ** -0.5
        self.wq = nn.Linear(config.hidden_dim, config.hidden_dim)
        self.wk = nn.Linear(config.hidden_dim, config.hidden_dim)
        self.wv = nn.Linear(config.hidden_dim, config.hidden_dim)
        self.attn_drop = nn.Dropout(config.dropout)
        self.proj = nn.Linear(config.hidden_dim, config.hidden_dim)
        self.proj_drop = nn.Dropout(config.dropout)

    def forward(self, x):
        CF._no_dropout(self.training, self.attn_drop.p, "CrossAttention")
        return CF.cross_attention(x, self.wq.weight, self.wq.bias, self.wk.weight, self.wk.bias, self.wv.weight, self.wv.bias,
                                  self.proj.weight, self.proj.bias, self.num_heads)


class CrossAttentionBlock(nn.Module):
    """/root/reference/model_cross.py:104-114: the residual uses the UN-normalised CLS row."""

    def __init__(self, config, act_layer=nn.GELU):
        super().__init__()
        self.attn = PreNorm(config, CrossAttention(config))
        self.ffn = PreNorm(config, FeedForward(config))

    def forward(self, x):
        y = self.attn(x) + x[:, 0:1]
        y = self.ffn(y) + y
        return y


class MultiScaleBlock(nn.Module):
    """/root/reference/model_cross.py:116-148."""

    def __init__(self, config, act_layer=nn.GELU):
        super().__init__()
        self.attn_order = config.attn_order
        self.blocks = nn.ModuleList([
            nn.Sequential(*[SelfAttentionBlock(config) for _ in range(config.num_self_blocks)])
            for _ in range(config.num_modalities)])
        self.fusion = nn.ModuleList([CrossAttentionBlock(config) for _ in range(len(self.attn_order))])

    def forward(self, x):
        attn = [block(x_) for x_, block in zip(x, self.blocks)]
        outs, cross_count = [], 0
        order = dict(self.attn_order)
        for i in range(len(attn)):
            if str(i) in order:     # all fusions read `attn` (this block's self-attention outputs), never `outs`
                j = int(order[str(i)])
                tmp = torch.cat((attn[i][:, 0:1], attn[j][:, 1:]), dim=1)
                tmp = self.fusion[cross_count](tmp)
                outs.append(torch.cat((tmp, attn[i][:, 1:]), dim=1))
                cross_count += 1
            else:
                outs.append(attn[i])
        return outs


class Transformer(nn.Module):
    """ModelVIT's encoder stack (/root/reference/modelv3.py:69-88): per layer
    [PreNorm(Attention), StochasticDepth(0), PreNorm(FeedForward), StochasticDepth(0)]."""

    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(config.num_layers):
            self.layers.append(nn.ModuleList([
                PreNorm(config, Attention(config, dim_head=config.hidden_dim // config.num_heads)),
                nn.Identity(),  # StochasticDepth(p=0, mode="row") is the identity and owns no parameters
                PreNorm(config, FeedForward(config)),
                nn.Identity(),
            ]))

    def forward(self, x):
        for attn, drop1, ff, drop2 in self.layers:
            x = drop1(attn(x)) + x
            x = drop2(ff(x)) + x
        return x


class _CavitFn(torch.autograd.Function):
    """One autograd node for the whole model: forward and backward are static kernel sequences."""

    @staticmethod
    def forward(ctx, engine, train, drop, img, labels, *params):
        logits, loss = engine.forward(img, labels, train=train, drop=drop)
        ctx.engine = engine
        ctx.train = train
        ctx.mark_non_differentiable(logits_out := logits.clone())
        return logits_out, loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, _dlogits, dloss):
        eng = ctx.engine
        hook = eng.on_range_done
        scale_dev = None
        if dloss is not None:  # consumed on the device: no host synchronisation between fwd and bwd
            scale_dev = dloss.detach().to(device=eng.device, dtype=torch.float32).reshape(1).contiguous()
        flat = eng.backward(loss_scale=1.0, on_range_done=hook, loss_scale_dev=scale_dev)
        grads = []
        for key, p in eng.params.items():
            if p.requires_grad:
                off, shp = eng.layout.slots[key]
                grads.append(flat[off:off + p.numel()].view(shp))
            else:
                grads.append(None)
        return (None, None, None, None, None, *grads)


class _CavitModel(_Base):
    _kind = "cross"

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        eng = self.__dict__.get("_engine_obj")
        if eng is not None:      # values changed in place: the bf16 operand copy must be re-derived
            eng._bf16_version = -1
        return out

    def _post_init(self, config):
        self.config = config
        self.patch_size = config.patch_size
        self.lr = config.lr
        self.weight_decay = config.weight_decay
        self.optim_params = config.optim_params
        self._dropout_p = float(config.dropout)
        self._engine_obj = None
        self._precision = str(getattr(config, "precision", "bf16"))
        self._epoch_metrics = {}
        self.initialize_model()

    def set_precision(self, precision: str):
        """"bf16" (default): bf16 GEMM / attention operands, fp32 accumulation — within ~2e-2 of the reference's fp32 path;
        "fp32": the fp32-tolerance mode (the reference trains in fp32: `L.Trainer` without `precision=`,
        /root/reference/main_mist.py:211-218) — every GEMM operand as a bf16 hi + lo pair (3 tensor-core MMAs per
        product), fp32 attention — within ~1e-3. Returns self."""
        if precision not in ("bf16", "fp32"):
            raise _abi.CavitError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        if precision != self._precision:
            self._precision = precision
            object.__setattr__(self, "_engine_obj", None)    # rebuilt (same parameters) on the next forward
        return self

    @staticmethod
    def init_weights(module):
        if isinstance(module, nn.Linear):
            nn.init.xavier_uniform_(module.weight)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            nn.init.ones_(module.weight)
            nn.init.zeros_(module.bias)

    def initialize_model(self):
        self.apply(type(self).init_weights)
        nn.init.normal_(self.pos_embedding, mean=0.0, std=0.02)
        nn.init.normal_(self.cls_token, mean=0.0, std=0.02)

    # ---------------------------------------------------------------- engine plumbing
    def engine(self) -> Engine:
        dev = self.pos_embedding.device
        if dev.type != "cuda":
            raise _abi.CavitError("cavit models run on a CUDA B200 only: move the model with .cuda() first "
                                  "(there is no CPU / eager fallback)")
        eng = self._engine_obj
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if eng is None or eng.device != dev:
            named = OrderedDict(self.named_parameters())
            eng = Engine(self._kind, self.config, named, dev, precision=self._precision)
            object.__setattr__(self, "_engine_obj", eng)
        return eng

    def forward(self, img, labels):
        eng = self.engine()
        params = list(eng.params.values())
        # grad mode is off inside autograd.Function.forward, so decide here whether to save activations
        train = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        drop = self.training and self._dropout_p > 0.0   # nn.Dropout semantics: active in train() mode only
        return _CavitFn.apply(eng, train, drop, img, labels, *params)

    # ---------------------------------------------------------------- Lightning-style hooks
    def log_stats(self, name, logits, labels):
        """The reference's per-step `log_stats` (/root/reference/model_cross.py:243-255, utils.py:18-62: six torchmetrics
        objects, six `.item()` read-backs and an AUROC sort per step, logged `on_epoch=True, sync_dist=True`). Here a step
        adds its batch to a device-side accumulator with one launch (cavit.metrics.EpochMetrics, no host synchronisation);
        the epoch values — the same batch-size-weighted means, averaged over ranks — are logged under the same keys from
        `on_train_epoch_end` / `on_validation_epoch_end`."""
        from .metrics import EpochMetrics
        em = self._epoch_metrics.get(name)
        if em is None or em.device != logits.device:
            em = self._epoch_metrics[name] = EpochMetrics(logits.device, prefix=name)
        em.update(logits, labels)

    def _log_epoch_stats(self, name):
        em = self._epoch_metrics.get(name)
        if em is None:
            return {}
        vals = em.compute()
        em.reset()
        for k, v in vals.items():
            if not k.endswith("_loss"):     # the loss is logged per step by training_step / validation_step, as in the reference
                self.log(k, v, on_epoch=True, on_step=False, sync_dist=False)    # already reduced over ranks
        return vals

    def on_train_epoch_end(self):
        self._log_epoch_stats("train")

    def on_validation_epoch_end(self):
        self._log_epoch_stats("val")

    def training_step(self, batch, batch_idx):
        x, labels = batch
        logits, loss = self(x, labels)
        self.log("train_loss", loss, on_epoch=True, on_step=False, sync_dist=True)
        self.log_stats("train", logits, labels)
        return loss

    def validation_step(self, batch, batch_idx):
        x, labels = batch
        logits, loss = self(x, labels)
        self.log("val_loss", loss, on_epoch=True, on_step=False, sync_dist=True)
        self.log_stats("val", logits, labels)

    def configure_optimizers(self):
        opt = torch.optim.Adam(self.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=self.optim_params["T_max"],
                                                           eta_min=self.optim_params["eta_min"])
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "interval": "epoch"}}

    def on_test_epoch_start(self):
        self.test_logits, self.test_targets = [], []

    def test_step(self, batch, batch_idx):
        x, labels = batch
        logits, _ = self(x, labels)
        self.test_logits.append(logits.cpu())
        self.test_targets.append(labels.cpu())

    def on_test_epoch_end(self):
        self.test_logits = torch.cat(self.test_logits)
        self.test_targets = torch.cat(self.test_targets)


def _check_divisible(config):
    assert all(config.img_size[i] % config.patch_size[i] == 0 for i in range(len(config.img_size))), \
        'image dimensions must be divisible by the patch size'


class ModelCross(_CavitModel):
    """CrossViT-style multi-sequence MRI classifier (drop-in for the reference's ModelCross)."""
    _kind = "cross"

    def __init__(self, config):
        super().__init__()
        _check_divisible(config)
        D, H, W = config.img_size
        dp, hp, wp = config.patch_size
        num_patches = (D // dp) * (H // hp) * (W // wp)
        self.label_smoothing = config.label_smoothing
        self.num_modalities = config.num_modalities
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, config.hidden_dim))
        self.patch_to_embedding = nn.Linear(dp * hp * wp, config.hidden_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, config.hidden_dim))
        self.dropout = nn.Dropout(config.dropout)
        self.transformer = nn.Sequential(*[MultiScaleBlock(config) for _ in range(config.num_multi_blocks)])
        self.norm = nn.ModuleList([nn.LayerNorm(config.hidden_dim) for _ in range(config.num_modalities)])
        self.mlp_head = nn.ModuleList([
            nn.Sequential(nn.Linear(config.hidden_dim, config.mlp_dim), nn.GELU(), nn.Dropout(config.dropout),
                          nn.Linear(config.mlp_dim, config.num_classes), nn.Dropout(config.dropout))
            for _ in range(config.num_modalities)])
        self._post_init(config)


class ModelVIT(_CavitModel):
    """Plain pre-norm ViT over the concatenation of all sequences' patch tokens (drop-in for the
    reference's ModelVIT)."""
    _kind = "vit"

    def __init__(self, config):
        super().__init__()
        _check_divisible(config)
        D, H, W = config.img_size
        dp, hp, wp = config.patch_size
        num_patches = (D // dp) * (H // hp) * (W // wp) * config.num_modalities
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, config.hidden_dim))
        self.patch_to_embedding = nn.Linear(dp * hp * wp, config.hidden_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, config.hidden_dim))
        self.dropout = nn.Dropout(config.dropout)
        self.transformer = Transformer(config)
        self.to_cls_token = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(config.hidden_dim), nn.Linear(config.hidden_dim, config.mlp_dim),
                                      nn.GELU(), nn.Dropout(config.dropout),
                                      nn.Linear(config.mlp_dim, config.num_classes), nn.Dropout(config.dropout))
        self._post_init(config)
