"""Drop-in classes for the reference's two CNN-stem models: ``ViT`` (/root/reference/model.py:218-286) and
``ViT3D`` (/root/reference/modelv2.py:103-241).

Same constructor signatures, module tree, ``state_dict`` keys and initialisation order (so a seed gives the
reference's weights and reference checkpoints load). What runs where:

* the CNN stems (``CNNEncoder`` / ``CNN3DEncoder``: Conv3d, BatchNorm3d, ReLU, MaxPool3d) are ordinary torch
  modules — cuDNN convolutions are outside this repository's hot path;
* everything after the stem — patch embedding / token assembly, the transformer blocks (biased packed QKV,
  fused attention, LayerNorm, FFN), the final norm, the head and the loss, forward and backward — runs on the
  sm_100a kernels through ``cavit.engine`` (kinds ``cnnvit`` / ``vit3d``); the engine hands d(stem features) back
  to autograd, which finishes the backward through the stem.

The transformer sub-modules below are parameter containers (their ``forward`` raises).
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _abi
from .engine import Engine
from . import functional as CF
from .modules import _Base, _engine_only


class _EncoderFn(torch.autograd.Function):
    """One autograd node from the stem's feature maps to (logits, loss)."""

    @staticmethod
    def forward(ctx, engine, train, feat, labels, *params):
        logits, loss = engine.forward(feat, labels, train=train, drop=False)
        ctx.engine = engine
        ctx.feat_grad = feat.requires_grad
        ctx.mark_non_differentiable(logits_out := logits.clone())
        return logits_out, loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, _dlogits, dloss):
        eng = ctx.engine
        scale_dev = None
        if dloss is not None:
            scale_dev = dloss.detach().to(device=eng.device, dtype=torch.float32).reshape(1).contiguous()
        flat = eng.backward(loss_scale=1.0, on_range_done=eng.on_range_done, loss_scale_dev=scale_dev)
        grads = []
        for key, p in eng.params.items():
            if p.requires_grad:
                off, shp = eng.layout.slots[key]
                grads.append(flat[off:off + p.numel()].view(shp))
            else:
                grads.append(None)
        dfeat = eng.dinput.clone() if ctx.feat_grad else None   # the engine reuses its buffer on the next step
        return (None, None, dfeat, None, *grads)


class _EncoderModel(_Base):
    _kind = ""
    _stem_prefix = ""

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        eng = self.__dict__.get("_engine_obj")
        if eng is not None:      # values changed in place: the bf16 operand copy must be re-derived
            eng._bf16_version = -1
        return out

    def _engine_cfg(self, num_modalities: int) -> SimpleNamespace:
        raise NotImplementedError

    def _anchor(self) -> torch.Tensor:
        raise NotImplementedError

    def engine(self, num_modalities: int) -> Engine:
        dev = self._anchor().device
        if dev.type != "cuda":
            raise _abi.CavitError("cavit models run on a CUDA B200 only: move the model with .cuda() first "
                                  "(there is no CPU / eager fallback)")
        eng = self.__dict__.get("_engine_obj")
        if eng is None or eng.device != dev or eng.Mimg != num_modalities:
            named = OrderedDict((k, p) for k, p in self.named_parameters() if not k.startswith(self._stem_prefix))
            eng = Engine(self._kind, self._engine_cfg(num_modalities), named, dev)
            object.__setattr__(self, "_engine_obj", eng)
        return eng

    def _run(self, feat, labels, num_modalities):
        eng = self.engine(num_modalities)
        params = list(eng.params.values())
        train = torch.is_grad_enabled() and (feat.requires_grad or any(p.requires_grad for p in params))
        return _EncoderFn.apply(eng, train, feat.float(), labels, *params)


# ======================================================================================== ViT (model.py)
class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        mid = mid_channels or out_channels
        convs = [nn.Conv3d(i, o, kernel_size=3, padding=1) for i, o in ((in_channels, mid), (mid, out_channels))]
        self.double_conv = nn.Sequential(convs[0], nn.ReLU(inplace=True), convs[1], nn.ReLU(inplace=True))

    def forward(self, x):
        return self.double_conv(x)


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool3d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        return self.maxpool_conv(x)


class CNNEncoder(nn.Module):
    """Per-modality feature extractor of `ViT` (torch / cuDNN; not on the hot path)."""

    def __init__(self, config, n_channels=1):
        super().__init__()
        self.n_channels = n_channels
        ch = config.encoder_channels
        self.inc = DoubleConv(n_channels, ch[0])
        self.down1 = Down(ch[0], ch[1])
        self.down2 = Down(ch[1], ch[2])

    def forward(self, x):
        return self.down2(self.down1(self.inc(x)))


class Embeddings(nn.Module):
    def __init__(self, config, n_channels=1):
        super().__init__()
        self.cnn_encoder = CNNEncoder(config, n_channels)
        grid = config.patches.grid
        self.patch_embed = nn.Conv3d(config.encoder_channels[2], config.hidden_size, kernel_size=grid, stride=grid)
        f = 2 ** config.down_factor
        num_patches = 1
        for i in range(3):
            num_patches *= config.img_size[i] / (f * grid[i])
        self.class_token = nn.Parameter(torch.zeros(1, 1, config.hidden_size))
        self.positional_embedding = nn.Parameter(torch.randn(1, int(num_patches + 1), config.hidden_size))
        self.dropout = nn.Dropout(config.transformer["dropout_rate"])
    forward = _engine_only("Embeddings")


class Mlp(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.fc1 = nn.Linear(config.hidden_size, config.transformer["mlp_dim"])
        self.fc2 = nn.Linear(config.transformer["mlp_dim"], config.hidden_size)
        self.act_fn = F.gelu
        self.dropout = nn.Dropout(config.transformer["dropout_rate"])

    def forward(self, x):   # /root/reference/model.py:107-121, through the fused GEMM epilogues (cavit/functional.py)
        CF._no_dropout(self.training, self.dropout.p, "Mlp")
        return CF.feed_forward(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)


class MultiHeadAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_attention_heads = config.transformer["num_heads"]
        self.attention_head_size = int(config.hidden_size / self.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.out = nn.Linear(config.hidden_size, config.hidden_size)
        self.attn_dropout = nn.Dropout(config.transformer["attention_dropout_rate"])
        self.proj_dropout = nn.Dropout(config.transformer["attention_dropout_rate"])
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x):   # /root/reference/model.py:123-176: biased q / k / v, scores / sqrt(d), out projection
        CF._no_dropout(self.training, self.attn_dropout.p, "MultiHeadAttention")
        wqkv = torch.cat((self.query.weight, self.key.weight, self.value.weight), 0)
        bqkv = torch.cat((self.query.bias, self.key.bias, self.value.bias), 0)
        return CF.self_attention(x, wqkv, self.out.weight, self.out.bias, self.num_attention_heads, bqkv)


class Block(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.multi_head = MultiHeadAttention(config)
        self.attention_norm = nn.LayerNorm(config.hidden_size, eps=1e-6)
        self.ffn_norm = nn.LayerNorm(config.hidden_size, eps=1e-6)
        self.ffn = Mlp(config)

    def forward(self, x):   # /root/reference/model.py:179-199 (pre-norm, eps 1e-6)
        x = x + self.multi_head(CF.layer_norm(x, self.attention_norm.weight, self.attention_norm.bias, self.attention_norm.eps))
        x = x + self.ffn(CF.layer_norm(x, self.ffn_norm.weight, self.ffn_norm.bias, self.ffn_norm.eps))
        return x


class Encoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.encoder_norm = nn.LayerNorm(config.hidden_size, eps=1e-6)
        self.layers = nn.Sequential(*[copy.deepcopy(Block(config)) for _ in range(config.transformer["num_layers"])])

    def forward(self, x):   # /root/reference/model.py:202-214
        x = self.layers(x)
        return CF.layer_norm(x, self.encoder_norm.weight, self.encoder_norm.bias, self.encoder_norm.eps)


class ViT(_EncoderModel):
    """CNN stem + pre-norm ViT over the concatenated modalities, one logit, BCE-with-logits (drop-in for the
    reference's ``model.ViT``)."""
    _kind = "cnnvit"
    _stem_prefix = "embeddings.cnn_encoder."

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = Embeddings(config)
        self.encoder = Encoder(config)
        self.final = nn.Linear(128, 1)   # the reference hard-codes 128: hidden_size must be 128
        self._init_weights()
        self.loss = nn.BCEWithLogitsLoss()

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def _anchor(self):
        return self.final.weight

    def _engine_cfg(self, num_modalities):
        c = self.config
        if c.hidden_size != self.final.in_features:
            raise _abi.CavitError(f"ViT: hidden_size must be {self.final.in_features} (final = Linear(128, 1))")
        f = 2 ** c.down_factor
        grid = tuple(int(g) for g in c.patches.grid)
        feat = tuple(int(c.img_size[i]) // f for i in range(3))
        if any(feat[i] % grid[i] for i in range(3)):
            raise _abi.CavitError("ViT: stem output must be divisible by patches.grid")
        npm = (feat[0] // grid[0]) * (feat[1] // grid[1]) * (feat[2] // grid[2])
        if npm + 1 != self.embeddings.positional_embedding.shape[1]:
            raise _abi.CavitError("ViT: positional table does not match the patch grid")
        return SimpleNamespace(hidden_dim=c.hidden_size, mlp_dim=c.transformer["mlp_dim"], num_heads=c.transformer["num_heads"],
                               num_layers=c.transformer["num_layers"], num_modalities=num_modalities, num_classes=1,
                               dropout=0.0, label_smoothing=0.0, in_channels=int(c.encoder_channels[2]), feat_dims=feat,
                               grid=grid, patches_per_modality=npm, patch_dim=int(c.encoder_channels[2]) * grid[0] * grid[1] * grid[2])

    def forward(self, x, label=None):
        t = self.config.transformer
        if self.training and (t["dropout_rate"] > 0 or t["attention_dropout_rate"] > 0):
            raise _abi.CavitError("cavit ViT: dropout > 0 in training mode is not supported (use rate 0)")
        B, M = x.shape[0], x.shape[1]
        # one batched stem call over (modality, sample): the stem has no cross-sample op (no BatchNorm), so this equals the
        # reference's per-modality calls (model.py:258)
        feat = self.embeddings.cnn_encoder(x.transpose(0, 1).reshape((M * B,) + tuple(x.shape[2:])))
        tgt = label if label is not None else torch.zeros(B, device=x.device)
        logits, loss = self._run(feat.contiguous(), tgt, M)
        return logits if label is None else (logits, loss)

    def training_step(self, batch, batch_idx):
        x, labels = batch
        _, loss = self(x, labels)
        self.log("train_loss", loss, on_epoch=True, sync_dist=True)
        return loss

    def validation_step(self, batch, batch_idx):
        x, labels = batch
        _, loss = self(x, labels)
        self.log("val_loss", loss, on_epoch=True, sync_dist=True)

    def configure_optimizers(self):
        opt = torch.optim.Adam(self.parameters(), lr=1e-3)
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.1, patience=5)
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "monitor": "val_loss"}}


# ======================================================================================== ViT3D (modelv2.py)
class CNN3DEncoder(nn.Module):
    """Four Conv3d + BatchNorm3d + ReLU stages (two max-pooled, two strided): 16x down-sampling stem of `ViT3D`
    (torch / cuDNN; not on the hot path)."""

    def __init__(self, in_channels: int = 1, hidden_dim: int = 256):
        super().__init__()
        widths = [in_channels, hidden_dim // 8, hidden_dim // 4, hidden_dim // 2, hidden_dim]
        for i, stride in enumerate((1, 1, 2, 2), start=1):
            setattr(self, f"conv{i}", nn.Conv3d(widths[i - 1], widths[i], kernel_size=3, stride=stride, padding=1))
            setattr(self, f"bn{i}", nn.BatchNorm3d(widths[i]))
        self.pool = nn.MaxPool3d(kernel_size=2, stride=2)

    def forward(self, x):
        for i in (1, 2, 3, 4):
            x = F.relu(getattr(self, f"bn{i}")(getattr(self, f"conv{i}")(x)))
            if i <= 2:
                x = self.pool(x)
        return x


class TransformerEncoder(nn.Module):
    """Owns an nn.TransformerEncoder (post-norm, ReLU, 4x FFN, batch_first) as the parameter container."""

    def __init__(self, embed_dim: int, num_heads: int, num_layers: int, dropout: float):
        super().__init__()
        layer = nn.TransformerEncoderLayer(d_model=embed_dim, nhead=num_heads, dim_feedforward=4 * embed_dim,
                                           dropout=dropout, batch_first=True)
        self.transformer = nn.TransformerEncoder(layer, num_layers=num_layers)
    forward = _engine_only("TransformerEncoder")


class ViT3D(_EncoderModel):
    """CNN stem per modality + post-norm Transformer encoder over all modalities' tokens (drop-in for the
    reference's ``modelv2.ViT3D``)."""
    _kind = "vit3d"
    _stem_prefix = "encoder_3d."

    def __init__(self, optimizer_params: dict, lr: float, weight_decay: float, num_modalities: int, config,
                 num_classes: int = 2, add_cls_token: bool = True, pretrained_cnn: bool = False,
                 cnn_out_dim: tuple = (64, 8, 8, 8), label_smoothing: float = 0.0, dropout: float = 0.0,
                 growth_rate: int = 16):
        super().__init__()
        self.lr, self.optimizer_params, self.weight_decay = lr, optimizer_params, weight_decay
        self.label_smoothing = label_smoothing
        if pretrained_cnn:
            raise _abi.CavitError("ViT3D(pretrained_cnn=True) needs MONAI's DenseNet121, which is outside this path")
        self.config, self.num_modalities, self.num_classes, self._dropout_p = config, num_modalities, num_classes, float(dropout)
        self.encoder_3d = CNN3DEncoder(hidden_dim=config.hidden_dim)
        self.add_cls_token = add_cls_token
        self.cls_token = nn.Parameter(torch.zeros(1, 1, config.hidden_dim)) if add_cls_token else None
        D, H, W = config.img_size
        num_tokens = (D // 16) * (H // 16) * (W // 16) * num_modalities
        self.pos_embed = nn.Parameter(torch.zeros(1, num_tokens + int(add_cls_token), config.hidden_dim))
        self.transformer = TransformerEncoder(embed_dim=config.hidden_dim, num_heads=config.transformer.num_heads,
                                              num_layers=config.transformer.num_layers, dropout=dropout)
        self.mlp_head = nn.Sequential(nn.LayerNorm(config.hidden_dim), nn.Linear(config.hidden_dim, config.hidden_dim // 8),
                                      nn.Linear(config.hidden_dim // 8, num_classes))
        self._init_weights()

    def _init_weights(self):
        nn.init.normal_(self.pos_embed, std=0.02)
        if self.cls_token is not None:
            nn.init.normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def _anchor(self):
        return self.pos_embed

    def _engine_cfg(self, num_modalities):
        c = self.config
        hc = int(self.add_cls_token)
        S = (self.pos_embed.shape[1] - hc) // num_modalities
        if S * num_modalities + hc != self.pos_embed.shape[1]:
            raise _abi.CavitError("ViT3D: pos_embed does not match the number of modalities")
        return SimpleNamespace(hidden_dim=c.hidden_dim, mlp_dim=4 * c.hidden_dim, num_heads=c.transformer.num_heads,
                               num_layers=c.transformer.num_layers, num_modalities=num_modalities,
                               num_classes=self.num_classes, dropout=0.0, label_smoothing=float(self.label_smoothing),
                               tokens_per_modality=S, num_tokens=self.pos_embed.shape[1], head_dim_hidden=c.hidden_dim // 8,
                               has_cls=bool(self.add_cls_token))

    def forward(self, x, labels):
        if self.training and self._dropout_p > 0:
            raise _abi.CavitError("cavit ViT3D: dropout > 0 in training mode is not supported (use dropout=0)")
        M = x.shape[1]
        # per-modality stem calls, like the reference: BatchNorm statistics are per call (modelv2.py:203-212)
        feat = torch.cat([self.encoder_3d(x.select(1, m)).flatten(start_dim=2) for m in range(M)], dim=2)
        return self._run(feat.contiguous(), labels, M)

    def training_step(self, batch, batch_idx):
        x, labels = batch
        _, loss = self(x, labels)
        self.log("train_loss", loss, on_epoch=True, on_step=False, sync_dist=True)
        return loss

    def validation_step(self, batch, batch_idx):
        x, labels = batch
        _, loss = self(x, labels)
        self.log("val_loss", loss, on_epoch=True, on_step=False, sync_dist=True)

    def configure_optimizers(self):
        opt = torch.optim.Adam(self.parameters(), lr=self.lr, weight_decay=self.weight_decay)
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=self.optimizer_params["factor"],
                                                           patience=self.optimizer_params["patience"])
        return {"optimizer": opt, "lr_scheduler": {"scheduler": sched, "monitor": self.optimizer_params["type"]}}
