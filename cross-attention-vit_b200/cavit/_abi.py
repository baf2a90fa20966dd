"""ctypes binding of libcavit_sm100a.so (include/cavit.h). There is NO fallback: if the shared
library is missing or the device is not a B200 (sm_100), importing/using the ops raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CAVIT_LIB") or os.path.join(_HERE, "libcavit_sm100a.so")   # CAVIT_LIB: an A/B build (build.py)

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p

ABI_VERSION = 3

EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID, EPI_GELU_BWD, EPI_EMBED, EPI_BIAS_RELU, EPI_RELU_BWD = range(8)


class CavitError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", c_i32), ("N", c_i32), ("K", c_i32), ("groups", c_i32),
        ("a_mn", c_i32), ("b_mn", c_i32), ("epi", c_i32), ("out_fp32", c_i32),
        ("A", c_vp), ("lda", c_i64), ("a_gs", c_i64),
        ("B", c_vp), ("ldb", c_i64), ("b_gs", c_i64),
        ("out", c_vp), ("ldo", c_i64), ("out_gs", c_i64),
        ("bias", c_vp), ("bias_gs", c_i64),
        ("resid", c_vp), ("ldr", c_i64), ("resid_gs", c_i64),
        ("aux", c_vp), ("ldaux", c_i64), ("aux_gs", c_i64),
        ("accumulate", c_i32), ("split_k", c_i32), ("embed_np", c_i32),
        ("A_lo", c_vp), ("B_lo", c_vp),     # fp32-tolerance mode: lo planes of split operands (NULL = plain bf16)
    ]


_SIGS = {
    "cavit_abi_version": (c_i32, []),
    "cavit_last_error": (C.c_char_p, []),
    "cavit_device_ok": (c_i32, [c_i32]),
    "cavit_device_status": (c_i32, [c_i32]),
    "cavit_launch_count": (C.c_longlong, []),
    "cavit_gemm": (c_i32, [C.POINTER(GemmArgs), c_vp]),
    "cavit_ln_fwd": (c_i32, [c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "cavit_ln_bwd_workspace_floats": (C.c_size_t, [c_i32, c_i32]),
    "cavit_ln_bwd": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64,
                             c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cavit_ln_fwd_dual": (c_i32, [c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cavit_ln_bwd_f32": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64,
                                 c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cavit_tokens_from_channels": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_tokens_to_channels": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_conv_patch_rows": (c_i32, [c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "cavit_conv_patch_rows_bwd": (c_i32, [c_vp, c_vp] + [c_i32] * 9 + [c_vp]),
    "cavit_token_mean_fwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "cavit_token_mean_bwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "cavit_bce_head_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp]),
    "cavit_bce_head_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp]),
    "cavit_batch_metrics": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp]),
    "cavit_stage_volumes": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "cavit_adam_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_f32, c_f32, c_f32, c_f32, c_f32, c_i64, c_f32, c_vp]),
    "cavit_ln_fusion_fwd": (c_i32, [c_vp, c_i64, c_vp, c_i32, c_i32, c_i32, c_i32, C.POINTER(c_i32), C.POINTER(c_i32),
                                    c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "cavit_ln_fusion_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32,
                                    C.POINTER(c_i32), C.POINTER(c_i32), c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cavit_attn_fwd": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "cavit_attn_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "cavit_xattn_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_vp, C.c_uint32, c_vp]),
    "cavit_xattn_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_vp,
                                C.c_uint32, c_vp]),
    "cavit_xfold_scratch_floats": (c_i64, [c_i32, c_i32, c_i32, c_i32]),
    "cavit_xfold_tensor_cores": (c_i32, [c_i32]),
    "cavit_xfold_fwd": (c_i32, [c_vp] * 12 + [c_i32] * 5 + [C.POINTER(c_i32), C.POINTER(c_i32), c_f32, c_f32, c_f32, c_vp,
                                C.c_uint32, c_vp]),
    "cavit_xfold_bwd": (c_i32, [c_vp] * 14 + [c_i32] * 5 + [C.POINTER(c_i32), C.POINTER(c_i32), c_f32, c_f32, c_vp,
                                C.c_uint32, c_i32, c_vp]),
    "cavit_expand_heads": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "cavit_fold_heads": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "cavit_dropout": (c_i32, [c_i32, c_vp, c_vp, c_vp, c_i64, c_f32, c_vp, C.c_uint32, c_vp]),
    "cavit_patchify": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_cls_rows": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_embed_param_grads": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_cast_bf16": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "cavit_colsum_bf16": (c_i32, [c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp]),
    "cavit_gather_rows_f32": (c_i32, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_gather_rows_f32_indexed": (c_i32, [c_vp, c_i64, c_i64, C.POINTER(c_i32), c_vp, c_i64, c_i64, C.POINTER(c_i32),
                                              c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_add_bf16_f32": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "cavit_gelu_bwd_bf16": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "cavit_compact_patch_rows_bf16": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "cavit_head_loss_fwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_vp,
                                    C.c_uint32, c_vp]),
    "cavit_head_loss_bwd": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32,
                                    c_f32, c_f32, c_vp, C.c_uint32, c_vp]),
    # ---- K-EMBED: TMA-staged patch unfold fused with the embedding GEMM (forward) and its weight gradient
    "cavit_embed_fused_supported": (c_i32, [c_i32] * 9),
    "cavit_embed_fused_wgrad_supported": (c_i32, [c_i32] * 9),
    "cavit_embed_fused_fwd": (c_i32, [c_vp] * 5 + [c_i32] * 10 + [c_vp]),
    "cavit_embed_fused_wgrad": (c_i32, [c_vp] * 3 + [c_i32] * 10 + [c_vp]),
    "cavit_embed_bias_grad": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp]),
    # ---- fp32-tolerance mode (split bf16 hi + lo operands, fp32 attention / GELU / head)
    "cavit_cast_split": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "cavit_gelu_split": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "cavit_gelu_bwd_split": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "cavit_ln_fwd_split": (c_i32, [c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cavit_ln_bwd_split": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i64,
                                   c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "cavit_colsum_split": (c_i32, [c_vp, c_vp, c_i64, c_i64, c_i32, c_i32, c_i32, c_vp, c_i64, c_vp]),
    "cavit_patchify_split": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "cavit_attn_fwd_f32": (c_i32, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "cavit_attn_bwd_f32": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "cavit_head_loss_fwd_f32": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "cavit_head_loss_bwd_f32": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32,
                                        c_f32, c_vp]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib():
    """Load the shared library (once). Raises CavitError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CavitError(
                f"{LIB_PATH} not found: build it with `python cross-attention-vit_b200/build.py` "
                "(cavit has no CPU / PyTorch fallback path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.cavit_abi_version() != ABI_VERSION:
            raise CavitError("libcavit_sm100a.so ABI version mismatch")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().cavit_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "cavit"):
    if rc != 0:
        raise CavitError(f"{what} failed (code {rc}): {last_error()}")


def require_device(index: int = 0):
    import torch
    if not torch.cuda.is_available():
        raise CavitError("cavit needs a CUDA device (B200, sm_100a); none is visible and there is no fallback")
    if not lib().cavit_device_ok(index):
        raise CavitError(f"cuda:{index} is not compute capability 10.x; cavit is built for sm_100a only")


def device_status(reset: bool = True) -> int:
    return int(lib().cavit_device_status(1 if reset else 0))


def launch_count() -> int:
    return int(lib().cavit_launch_count())
