"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the reference's deterministic input chain.

``BrainDataset`` without augmentation (/root/reference/dataset_ucsf.py:81-89, 121-134, 149-158) applies, per volume,

    LoadImaged(reader='nibabelreader') -> EnsureChannelFirstd(no_channel) -> ResizeWithPadOrCropd(img_size, -1)
    -> ToTensord -> .to(torch.float)

PARITY UNPINNED: the arithmetic lives in two third-party packages that are absent from this image and unpinned by the
reference (`requirements.txt` names monai and nibabel without versions), and the reference holds no test or golden
vector for it. What is restated here is their published behaviour:

* nibabel `ArrayProxy` read scaling (`apply_read_scaling`): stored array (file order = Fortran order over dim[1..3])
  times `scl_slope`, plus `scl_inter`, as two float64 numpy operations, skipped when (slope, inter) == (1, 0); a slope
  that is 0 or not finite means "no scaling" (`Nifti1Header.get_slope_inter`), a finite slope with a non-finite
  intercept is an error;
* MONAI `ResizeWithPadOrCrop`: `CenterSpatialCrop(roi_size)` (start = size // 2 - roi // 2, only where size > roi) and then
  `SpatialPad(method="symmetric")` (before = (roi - size) // 2, the odd voxel after) with `constant_values`;
* the result as a C-contiguous float32 array `[1, D, H, W]`.

The reader is checked against the reference's own UCSF-PDGM files where /root/reference is present
(tests/test_staging_cpu.py); the crop / pad rule must be re-verified wherever MONAI exists.
"""
from __future__ import annotations

import gzip
import struct
from typing import Sequence, Tuple

import numpy as np

NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                768: np.uint32}


def read_scaling(stored: np.ndarray, slope: float, inter: float) -> np.ndarray:
    """nibabel's read scaling of a stored array (any shape); returns float64 unless nothing is applied."""
    slope, inter = float(np.float32(slope)), float(np.float32(inter))
    if slope == 0 or not np.isfinite(slope):
        return stored
    if not np.isfinite(inter):
        raise ValueError("valid slope but invalid intercept")
    if (slope, inter) == (1.0, 0.0):
        return stored
    out = stored.astype(np.float64)
    if slope != 1.0:
        out = out * np.float64(slope)
    if inter != 0.0:
        out = out + np.float64(inter)
    return out


def resize_with_pad_or_crop(arr: np.ndarray, size: Sequence[int], pad_value: float) -> np.ndarray:
    """Centre crop then symmetric constant pad of a 3-D array to ``size``."""
    sl = []
    for s, t in zip(arr.shape, size):
        start = s // 2 - t // 2 if s > t else 0
        sl.append(slice(start, start + min(s, t)))
    arr = arr[tuple(sl)]
    widths = []
    for s, t in zip(arr.shape, size):
        w = max(t - s, 0)
        widths.append((w // 2, w - w // 2))
    return np.pad(arr, widths, mode="constant", constant_values=pad_value)


def stage_volume(stored_flat: np.ndarray, dims: Tuple[int, int, int], slope: float, inter: float,
                 img_size: Sequence[int], pad_value: float = -1.0) -> np.ndarray:
    """One volume: stored voxels in file order -> float32 [1, D, H, W]."""
    arr = np.asarray(stored_flat).reshape(tuple(dims), order="F")
    arr = read_scaling(arr, slope, inter).astype(np.float32)     # LoadImage(dtype=np.float32), before the resize
    arr = resize_with_pad_or_crop(arr, img_size, pad_value)
    return np.ascontiguousarray(arr)[None]


def stage_batch(samples, img_size: Sequence[int], pad_value: float = -1.0) -> np.ndarray:
    """``samples[b][m] = (stored_flat, dims, slope, inter)`` -> float32 [B, M, 1, D, H, W]
    (``torch.stack(data)`` per sample, dataset_ucsf.py:158, then the DataLoader's batch stack)."""
    return np.stack([np.stack([stage_volume(*v, img_size, pad_value) for v in s]) for s in samples])


# ------------------------------------------------------------------------------------------ synthetic files for tests
def write_nifti(path: str, arr: np.ndarray, slope: float = 1.0, inter: float = 0.0, *, big_endian: bool = False,
                extension_bytes: int = 0) -> None:
    """Write a single-file NIfTI-1 (.nii or .nii.gz) holding ``arr`` (3-D, index [i, j, k]) as its stored array."""
    code = {v: k for k, v in NIFTI_DTYPES.items()}[arr.dtype.type]
    e = ">" if big_endian else "<"
    vox_offset = 352 + extension_bytes
    hdr = bytearray(348)
    struct.pack_into(e + "i", hdr, 0, 348)
    struct.pack_into(e + "8h", hdr, 40, 3, *arr.shape, 1, 1, 1, 1)
    struct.pack_into(e + "hh", hdr, 70, code, arr.dtype.itemsize * 8)
    struct.pack_into(e + "8f", hdr, 76, 1, 1, 1, 1, 1, 1, 1, 1)
    struct.pack_into(e + "fff", hdr, 108, float(vox_offset), slope, inter)
    hdr[344:348] = b"n+1\0"
    body = bytes(hdr) + (b"\1\0\0\0" if extension_bytes else b"\0\0\0\0") + bytes(extension_bytes)
    data = np.asfortranarray(arr).astype(arr.dtype.newbyteorder(e)).tobytes(order="F")
    with (gzip.open(path, "wb") if path.endswith(".gz") else open(path, "wb")) as f:
        f.write(body + data)
