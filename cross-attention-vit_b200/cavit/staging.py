"""Real-volume input staging: NIfTI files -> the fp32 batch tensor the models take.

The reference builds every sample on the host (/root/reference/dataset_ucsf.py:81-89,121-134,149-158): nibabel
inflates the file and scales the stored int16 voxels to float64, MONAI crops / pads to ``img_size`` with -1, the result is
copied to a float32 tensor, stacked, pinned and sent over PCIe as fp32. Here the host does only what has to happen
there — inflate the file and parse the 348-byte header (``read_nifti``) — and ``VolumeStager`` ships the STORED bytes
of the centre-crop window (2 bytes per voxel for the UCSF volumes) plus one 32-byte descriptor per volume; a single CUDA launch
(``cavit_stage_volumes``, csrc/stage.cu) applies scaling, centre crop / symmetric pad and the axis transposition for the
whole batch. The augmentations of the training chain (dataset_ucsf.py:93-113) are host-side MONAI transforms and stay
outside this path.

There is no CPU path: staging needs the CUDA library and a B200, like everything else in this package.
"""
from __future__ import annotations

import gzip
import os
import struct
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _abi, ops

# NIfTI-1 datatype code -> (numpy dtype, CAVIT_VOX_* of include/cavit.h)
_NIFTI_TYPES = {2: (np.uint8, 0), 4: (np.int16, 1), 8: (np.int32, 2), 16: (np.float32, 3), 64: (np.float64, 4),
                256: (np.int8, 5), 512: (np.uint16, 6), 768: (np.uint32, 7)}
_VOX_CODE = {np.dtype(t).str[1:]: c for t, c in _NIFTI_TYPES.values()}
DESC_DTYPE = np.dtype([("byte_offset", "<i8"), ("dims", "<i4", (3,)), ("dtype", "<i4"), ("slope", "<f4"), ("inter", "<f4")])
assert DESC_DTYPE.itemsize == 32


@dataclass
class RawVolume:
    """Stored voxels of one volume in file order (axis 0 fastest), native byte order, with the header's scaling."""
    data: np.ndarray                 # 1-D, one of the _NIFTI_TYPES dtypes
    dims: Tuple[int, int, int]
    slope: float = 1.0               # after the validity rule below: (1, 0) = take stored values
    inter: float = 0.0

    def __post_init__(self):
        self.data = np.ascontiguousarray(self.data).reshape(-1)
        if self.data.dtype.str[1:] not in _VOX_CODE or not self.data.dtype.isnative:
            raise _abi.CavitError(f"RawVolume: unsupported voxel type {self.data.dtype}")
        if len(self.dims) != 3 or min(self.dims) < 1 or int(np.prod(self.dims)) != self.data.size:
            raise _abi.CavitError(f"RawVolume: dims {self.dims} do not match {self.data.size} stored voxels")
        self.slope, self.inter = _valid_scaling(self.slope, self.inter)


def _valid_scaling(slope: float, inter: float) -> Tuple[float, float]:
    """nibabel's rule for scl_slope / scl_inter: slope 0, NaN or inf = unscaled; a valid slope needs a finite intercept."""
    slope, inter = float(np.float32(slope)), float(np.float32(inter))
    if slope == 0 or not np.isfinite(slope):
        return 1.0, 0.0
    if not np.isfinite(inter):
        raise _abi.CavitError("NIfTI header: valid scl_slope but invalid scl_inter")
    return slope, inter


def read_nifti(path: str) -> RawVolume:
    """Parse a single-file NIfTI-1 image (.nii / .nii.gz, either byte order, header extensions skipped through
    vox_offset) WITHOUT scaling or reordering its voxels."""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:2] == b"\x1f\x8b":
        blob = gzip.decompress(blob)
    if len(blob) < 352:
        raise _abi.CavitError(f"{path}: too short for a NIfTI-1 header")
    for e in ("<", ">"):
        if struct.unpack_from(e + "i", blob, 0)[0] == 348:
            break
    else:
        raise _abi.CavitError(f"{path}: sizeof_hdr is not 348 (NIfTI-2 / Analyze files are not supported)")
    if blob[344:348] != b"n+1\0":
        raise _abi.CavitError(f"{path}: not a single-file NIfTI-1 image (magic {blob[344:348]!r})")
    dim = struct.unpack_from(e + "8h", blob, 40)
    if dim[0] < 3 or dim[0] > 7 or any(d != 1 for d in dim[4:1 + dim[0]]) or min(dim[1:4]) < 1:
        raise _abi.CavitError(f"{path}: expected one 3-D volume, header dim = {dim}")
    code = struct.unpack_from(e + "h", blob, 70)[0]
    if code not in _NIFTI_TYPES:
        raise _abi.CavitError(f"{path}: unsupported NIfTI datatype code {code}")
    vox_offset, slope, inter = struct.unpack_from(e + "fff", blob, 108)
    dt = np.dtype(_NIFTI_TYPES[code][0])
    n = dim[1] * dim[2] * dim[3]
    off = int(vox_offset)
    if off < 352 or off + n * dt.itemsize > len(blob):
        raise _abi.CavitError(f"{path}: voxel data ({n} x {dt}) at offset {off} does not fit the file")
    data = np.frombuffer(blob, dtype=dt.newbyteorder(e), count=n, offset=off)
    if not data.dtype.isnative:
        data = data.astype(dt)
    return RawVolume(data, (dim[1], dim[2], dim[3]), slope, inter)


def plan_batch(vols: Sequence[RawVolume], img_size: Sequence[int]):
    """Layout of one batch's transfer: ``V`` descriptors, then each volume's centre-crop window (16-byte aligned).

    Only the stored voxels inside the window cross PCIe (for 240 x 240 x 155 files and a 128 x 128 x 64 target that is 12 %
    of the file); the descriptor carries the WINDOW's extents, for which the kernel's own crop is the identity and its pad
    widths are those of the full volume. Returns (descriptors, per-volume [(start, length)] * 3, total bytes)."""
    V = len(vols)
    wins = [[(s // 2 - t // 2 if s > t else 0, min(s, t)) for s, t in zip(v.dims, img_size)] for v in vols]
    desc = np.zeros(V, dtype=DESC_DTYPE)
    base = V * DESC_DTYPE.itemsize
    off = base
    for i, (v, w) in enumerate(zip(vols, wins)):
        off = (off + 15) & ~15
        desc[i] = (off - base, [n for _, n in w], _VOX_CODE[v.data.dtype.str[1:]], v.slope, v.inter)
        off += w[0][1] * w[1][1] * w[2][1] * v.data.itemsize
    return desc, wins, off


def pack_batch(vols: Sequence[RawVolume], desc: np.ndarray, wins, hv: np.ndarray, pool=None) -> None:
    """Write the descriptors and the crop windows (file order kept: axis 0 fastest) into the byte buffer ``hv``."""
    V = len(vols)
    base = V * DESC_DTYPE.itemsize
    hv[:base] = desc.view(np.uint8)

    def pack(i):
        v, w = vols[i], wins[i]
        ext = tuple(n for _, n in w)
        o = base + int(desc[i]["byte_offset"])
        dst = hv[o:o + int(np.prod(ext)) * v.data.itemsize].view(v.data.dtype).reshape(ext, order="F")
        np.copyto(dst, v.data.reshape(v.dims, order="F")[tuple(slice(a, a + n) for a, n in w)])

    if pool is not None and V > 1:     # numpy's strided copy releases the GIL
        list(pool.map(pack, range(V)))
    else:
        for i in range(V):
            pack(i)


class VolumeStager:
    """Builds the ``[B, M, 1, D, H, W]`` fp32 device batch from stored volumes with one H2D copy of the stored bytes and
    one kernel launch. ``stage`` may be called every step: the pinned and device buffers are reused and grown on demand."""

    def __init__(self, img_size: Sequence[int], device, pad_value: float = -1.0, workers: Optional[int] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _abi.CavitError("VolumeStager needs a CUDA device (there is no CPU path)")
        _abi.require_device(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.img_size = tuple(int(s) for s in img_size)
        if len(self.img_size) != 3 or min(self.img_size) < 1:
            raise _abi.CavitError(f"img_size must be three positive extents, got {img_size}")
        self.pad_value = float(pad_value)
        workers = min(8, os.cpu_count() or 1) if workers is None else int(workers)
        self._pool = ThreadPoolExecutor(max_workers=workers) if workers > 1 else None   # packs the crop windows
        self._host: Optional[torch.Tensor] = None      # pinned uint8: descriptors, then the stored bytes
        self._dev: Optional[torch.Tensor] = None
        self._copied: Optional[torch.cuda.Event] = None
        self._raw_dev: Optional[torch.Tensor] = None   # device view of the last batch's stored bytes (tools, tests)
        self.h2d_bytes = 0                              # of the last call

    def _buffers(self, nbytes: int):
        if self._host is None or self._host.numel() < nbytes:
            cap = max(nbytes, 1 << 20)
            self._host = torch.empty(cap, dtype=torch.uint8).pin_memory()
            self._dev = torch.empty(cap, dtype=torch.uint8, device=self.device)
            self._copied = None
        elif self._copied is not None:
            self._copied.synchronize()                  # the previous batch has left the pinned buffer
        return self._host, self._dev

    def stage(self, samples: Sequence[Sequence[RawVolume]], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = len(samples)
        if B == 0:
            raise _abi.CavitError("VolumeStager.stage: empty batch")
        M = len(samples[0])
        if M == 0 or any(len(s) != M for s in samples):
            raise _abi.CavitError("VolumeStager.stage: every sample needs the same, non-zero number of volumes")
        vols = [v for s in samples for v in s]
        V = len(vols)
        desc, wins, off = plan_batch(vols, self.img_size)
        base = V * DESC_DTYPE.itemsize
        host, dev = self._buffers(off)
        pack_batch(vols, desc, wins, host.numpy(), self._pool)
        D, H, W = self.img_size
        if out is None:
            out = torch.empty(B, M, 1, D, H, W, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != (B, M, 1, D, H, W) or out.dtype != torch.float32 or not out.is_contiguous() \
                or out.device != self.device:
            raise _abi.CavitError("VolumeStager.stage: out must be a contiguous fp32 [B, M, 1, D, H, W] tensor on the device")
        with torch.cuda.device(self.device):
            dev[:off].copy_(host[:off], non_blocking=True)
            self._copied = torch.cuda.Event()
            self._copied.record()
            ops.stage_volumes(dev[base:], dev, out, volumes=V, D=D, H=H, W=W, pad_value=self.pad_value)
        self._raw_dev = dev[base:]
        self.h2d_bytes = off
        return out
