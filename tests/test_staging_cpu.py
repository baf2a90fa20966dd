"""CPU: NIfTI reader and the numpy restatement of the reference's deterministic input chain
(/root/reference/dataset_ucsf.py:81-89,121-134,149-158; oracle/staging.py — parity unpinned, see its header)."""
import glob
import os

import numpy as np
import pytest

from oracle import staging as O


def test_crop_and_pad_known_answers():
    """Hand-computed windows of MONAI's centre crop / symmetric pad rule."""
    a = np.arange(7, dtype=np.float32).reshape(7, 1, 1)
    # 7 -> 4: start = 7 // 2 - 4 // 2 = 1
    assert O.resize_with_pad_or_crop(a, (4, 1, 1), -1).ravel().tolist() == [1, 2, 3, 4]
    # 7 -> 3: start = 3 - 1 = 2
    assert O.resize_with_pad_or_crop(a, (3, 1, 1), -1).ravel().tolist() == [2, 3, 4]
    # 7 -> 10: 3 pad voxels, 1 before and 2 after
    assert O.resize_with_pad_or_crop(a, (10, 1, 1), -1).ravel().tolist() == [-1, 0, 1, 2, 3, 4, 5, 6, -1, -1]
    # mixed: crop one axis, pad another, keep the third
    b = np.arange(2 * 5 * 3, dtype=np.float32).reshape(2, 5, 3)
    r = O.resize_with_pad_or_crop(b, (3, 2, 3), -1)
    assert r.shape == (3, 2, 3)
    assert np.all(r[0] == b[0, 1:3]) and np.all(r[1] == b[1, 1:3]) and np.all(r[2] == -1)


def test_read_scaling_rules():
    s = np.array([-3, 0, 7], dtype=np.int16)
    assert O.read_scaling(s, 0.0, 5.0) is s                  # slope 0: unscaled, intercept ignored
    assert O.read_scaling(s, float("nan"), 5.0) is s
    assert O.read_scaling(s, 1.0, 0.0) is s
    out = O.read_scaling(s, 0.5, 2.0)
    assert out.dtype == np.float64 and out.tolist() == [0.5, 2.0, 5.5]
    with pytest.raises(ValueError):
        O.read_scaling(s, 2.0, float("inf"))
    # header fields are float32: the float64 arithmetic uses the float32-rounded slope
    slope = 0.17052756249904633
    assert O.read_scaling(s, slope, 0.0)[2] == 7.0 * float(np.float32(slope))


@pytest.mark.parametrize("dtype", [np.int16, np.uint8, np.float32, np.uint16, np.int32, np.float64, np.int8, np.uint32])
@pytest.mark.parametrize("big_endian,gz,ext", [(False, True, 0), (True, False, 0), (False, False, 2896)])
def test_reader_round_trip(tmp_path, dtype, big_endian, gz, ext):
    from cavit.staging import read_nifti
    rng = np.random.default_rng(5)
    info = np.iinfo(dtype) if np.issubdtype(dtype, np.integer) else None
    arr = (rng.integers(info.min, info.max, size=(5, 4, 3), endpoint=True).astype(dtype) if info
           else rng.standard_normal((5, 4, 3)).astype(dtype))
    path = str(tmp_path / ("v.nii.gz" if gz else "v.nii"))
    O.write_nifti(path, arr, slope=0.25, inter=-3.0, big_endian=big_endian, extension_bytes=ext)
    vol = read_nifti(path)
    assert vol.dims == (5, 4, 3) and vol.data.dtype == np.dtype(dtype) and (vol.slope, vol.inter) == (0.25, -3.0)
    assert np.array_equal(vol.data.reshape(vol.dims, order="F"), arr)


def test_reader_rejects_what_it_cannot_stage(tmp_path):
    from cavit import CavitError
    from cavit.staging import RawVolume, read_nifti
    p = str(tmp_path / "short.nii")
    open(p, "wb").write(b"\0" * 100)
    with pytest.raises(CavitError):
        read_nifti(p)
    arr = np.zeros((2, 2, 2), np.int16)
    good = str(tmp_path / "g.nii")
    O.write_nifti(good, arr)
    blob = bytearray(open(good, "rb").read())
    bad = str(tmp_path / "b.nii")
    open(bad, "wb").write(bytes(blob[:-4]))                     # truncated voxel data
    with pytest.raises(CavitError):
        read_nifti(bad)
    blob[344:348] = b"ni1\0"                                    # header / image pair
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(CavitError):
        read_nifti(bad)
    with pytest.raises(CavitError):
        RawVolume(np.zeros(7, np.int16), (2, 2, 2))
    with pytest.raises(CavitError):
        RawVolume(np.zeros(8, np.int16), (2, 2, 2), slope=2.0, inter=float("nan"))
    assert RawVolume(np.zeros(8, np.int16), (2, 2, 2), slope=0.0, inter=9.0).slope == 1.0   # invalid slope: unscaled


def test_stager_has_no_cpu_path():
    import torch
    from cavit import CavitError
    from cavit.staging import VolumeStager
    with pytest.raises(CavitError):
        VolumeStager((8, 8, 8), torch.device("cpu"))


def test_reader_on_the_reference_volumes():
    """The reference ships UCSF-PDGM volumes (int16, 240 x 240 x 155, header extensions: vox_offset 3248). Container only."""
    files = sorted(glob.glob("/root/reference/ucsf-data/*/*_T1c.nii.gz"))[:2] + \
        sorted(glob.glob("/root/reference/ucsf-data/*/*_tumor_segmentation.nii.gz"))[:1]
    if not files:
        pytest.skip("reference data not present")
    from cavit.staging import read_nifti
    for f in files:
        v = read_nifti(f)
        assert v.dims == (240, 240, 155)
        seg = "segmentation" in os.path.basename(f)
        assert v.data.dtype == (np.uint8 if seg else np.int16)
        assert (v.slope, v.inter) == (1.0, 0.0) if seg else v.slope > 0
        out = O.stage_volume(v.data, v.dims, v.slope, v.inter, (112, 256, 155))
        assert out.shape == (1, 112, 256, 155) and out.dtype == np.float32
        assert np.all(out[0, :, :8] == -1) and np.all(out[0, :, -8:] == -1)       # 240 -> 256: 8 pad rows each side
        stored = v.data.reshape(v.dims, order="F")[120 - 56:120 + 56]            # 240 -> 112: start 64
        assert np.array_equal(out[0, :, 8:-8], O.read_scaling(stored, v.slope, v.inter).astype(np.float32))
        if not seg:
            assert out.min() >= -1.0 and out.max() > 100.0                       # intensities in scanner units


def test_transfer_layout_carries_exactly_the_crop_windows():
    """Host logic of VolumeStager (plan_batch / pack_batch): descriptors + 16-byte aligned crop windows in file order.
    Decoding the packed windows with the oracle must give what the oracle gives for the whole volumes."""
    from concurrent.futures import ThreadPoolExecutor
    from cavit.staging import DESC_DTYPE, RawVolume, pack_batch, plan_batch
    rng = np.random.default_rng(21)
    img_size = (12, 20, 7)
    specs = [((30, 9, 7), np.int16, 0.5, 3.0), ((12, 20, 7), np.uint8, 1.0, 0.0), ((5, 41, 16), np.float32, 0.0, 1.0),
             ((13, 21, 8), np.float64, 2.0, -1.0), ((1, 1, 1), np.int32, 1.0, 2.0)]
    vols = []
    for dims, dt, s, i in specs:
        arr = (rng.integers(-100, 100, size=dims) if np.issubdtype(dt, np.integer) and dt != np.uint8
               else rng.integers(0, 200, size=dims)).astype(dt)
        vols.append(RawVolume(np.asfortranarray(arr).ravel(order="F"), dims, s, i))
    desc, wins, total = plan_batch(vols, img_size)
    base = len(vols) * DESC_DTYPE.itemsize
    assert [tuple(d["dims"]) for d in desc] == [(12, 9, 7), (12, 20, 7), (5, 20, 7), (12, 20, 7), (1, 1, 1)]
    assert wins[0][0] == (30 // 2 - 12 // 2, 12) and wins[2][1] == (41 // 2 - 20 // 2, 20) and wins[3][2] == (8 // 2 - 7 // 2, 7)
    assert all(int(d["byte_offset"]) % 16 == 0 for d in desc)
    for pool in (None, ThreadPoolExecutor(max_workers=3)):
        hv = np.full(total + 8, 0xAB, dtype=np.uint8)
        pack_batch(vols, desc, wins, hv, pool)
        assert np.all(hv[total:] == 0xAB)                         # nothing written past the planned size
        got = np.frombuffer(hv[:base].tobytes(), dtype=DESC_DTYPE)
        assert np.array_equal(got["dims"], desc["dims"]) and np.array_equal(got["byte_offset"], desc["byte_offset"])
        for v, d in zip(vols, desc):
            n = int(np.prod(d["dims"]))
            o = base + int(d["byte_offset"])
            win = np.frombuffer(hv[o:o + n * v.data.itemsize].tobytes(), dtype=v.data.dtype)
            a = O.stage_volume(win, tuple(int(x) for x in d["dims"]), float(d["slope"]), float(d["inter"]), img_size)
            b = O.stage_volume(v.data, v.dims, v.slope, v.inter, img_size)
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
