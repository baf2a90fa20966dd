"""GPU: cavit_stage_volumes through the C ABI against the numpy restatement of the reference's input chain
(oracle/staging.py; /root/reference/dataset_ucsf.py:81-89,121-134,149-158). Bit-exact: integer index work plus two
float64 operations and one rounding to fp32."""
import numpy as np
import pytest
import torch

from oracle import staging as O

pytestmark = pytest.mark.gpu


def _raw(rng, dtype, dims):
    if np.issubdtype(dtype, np.integer):
        info = np.iinfo(dtype)
        lo, hi = max(info.min, -30000), min(info.max, 30000)
        arr = rng.integers(lo, hi, size=dims, endpoint=True).astype(dtype)
    else:
        arr = (rng.standard_normal(dims) * 1000).astype(dtype)
    return np.asfortranarray(arr).ravel(order="F")


def _check(samples, img_size, pad=-1.0, stager=None):
    from cavit.staging import RawVolume, VolumeStager
    st = stager or VolumeStager(img_size, "cuda:0", pad_value=pad)
    got = st.stage([[RawVolume(*v) for v in s] for s in samples])
    torch.cuda.synchronize()
    want = O.stage_batch(samples, img_size, pad)
    assert tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    return st


def _check_abi(samples, img_size, pad=-1.0):
    """The entry point itself with WHOLE stored volumes (VolumeStager ships only the crop window, so the kernel's own
    centre crop is exercised here)."""
    from cavit import ops
    from cavit.staging import DESC_DTYPE, _VOX_CODE
    vols = [v for s in samples for v in s]
    desc = np.zeros(len(vols), dtype=DESC_DTYPE)
    blob, off = [], 0
    for i, (data, dims, slope, inter) in enumerate(vols):
        fill = (-off) % 16
        blob.append(np.zeros(fill, np.uint8))
        off += fill
        if not (slope != 0 and np.isfinite(slope)):
            slope, inter = 1.0, 0.0
        desc[i] = (off, dims, _VOX_CODE[data.dtype.str[1:]], slope, inter)
        blob.append(data.view(np.uint8))
        off += data.nbytes
    raw = torch.from_numpy(np.concatenate(blob)).cuda()
    dsc = torch.from_numpy(desc.view(np.uint8).copy()).cuda()
    D, H, W = img_size
    out = torch.empty(len(samples), len(samples[0]), 1, D, H, W, device="cuda:0")
    ops.stage_volumes(raw, dsc, out, volumes=len(vols), D=D, H=H, W=W, pad_value=pad)
    torch.cuda.synchronize()
    want = O.stage_batch(samples, img_size, pad)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("img_size", [(16, 24, 8), (33, 31, 40), (7, 70, 3), (64, 64, 1), (1, 1, 1), (5, 2, 65),
                                      (130, 3, 33), (300, 40, 1)])
def test_ragged_volumes_crop_pad_and_transpose(img_size):
    """Every volume of the batch has its own extents: larger, smaller and equal to the target, odd and even, per axis."""
    rng = np.random.default_rng(1)
    dims = [(16, 24, 8), (17, 23, 9), (40, 5, 8), (3, 50, 41), (1, 1, 1), (335, 70, 2)]
    samples = [[(_raw(rng, np.int16, d), d, 0.170527562, 5587.847168) for d in dims[:3]],
               [(_raw(rng, np.int16, d), d, 0.043116443, 1412.8396) for d in dims[3:]]]
    _check(samples, img_size)
    _check_abi(samples, img_size, pad=0.5)


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.int32, np.float32, np.float64, np.int8, np.uint16, np.uint32])
def test_every_stored_type_and_scaling_rule(dtype):
    rng = np.random.default_rng(2)
    d = (19, 12, 21)
    scalings = [(1.0, 0.0), (0.0, 7.0), (float("nan"), 1.0), (1.0, -12.5), (8.828195063870226e-08, 0.0026368231046944857),
                (3.0, 0.0)]
    samples = [[(_raw(rng, dtype, d), d, s, i) for s, i in scalings]]
    _check(samples, (16, 16, 16))
    _check_abi(samples, (16, 16, 16))


def test_mixed_types_in_one_batch_and_buffer_reuse():
    rng = np.random.default_rng(3)
    a, b = (20, 20, 20), (9, 31, 14)
    s1 = [[(_raw(rng, np.int16, a), a, 0.5, 1.0), (_raw(rng, np.uint8, b), b, 1.0, 0.0)],
          [(_raw(rng, np.float64, b), b, 2.0, -1.0), (_raw(rng, np.float32, a), a, 0.0, 0.0)]]
    st = _check(s1, (24, 24, 16), pad=-1.0)
    s2 = [[(_raw(rng, np.uint16, b), b, 0.25, 0.0), (_raw(rng, np.int32, a), a, 1.0, 3.0)]]
    _check(s2, (24, 24, 16), stager=st)                      # smaller batch through the same buffers
    big = (70, 65, 40)
    s3 = [[(_raw(rng, np.int16, big), big, 0.1, 0.0)] * 2] * 3
    _check(s3, (24, 24, 16), stager=st)                      # larger than the first allocation: buffers grow


def test_full_size_ucsf_shape_into_model_input(tmp_path):
    """240 x 240 x 155 int16 files (the reference's data, /root/reference/dataset_ucsf.py:149-158) -> cfg1's and cfg3's
    img_size; through read_nifti and a caller-provided output tensor."""
    from cavit.staging import VolumeStager, read_nifti
    rng = np.random.default_rng(4)
    dims = (240, 240, 155)
    vols, raws = [], []
    for m in range(2):
        arr = rng.integers(-32768, 32767, size=dims, endpoint=True).astype(np.int16)
        p = str(tmp_path / f"m{m}.nii.gz")
        O.write_nifti(p, arr, slope=0.05 * (m + 1), inter=1645.089355, extension_bytes=2896)
        vols.append(read_nifti(p))
        raws.append((vols[-1].data, vols[-1].dims, vols[-1].slope, vols[-1].inter))
    for img_size in [(128, 128, 64), (240, 240, 160)]:
        st = VolumeStager(img_size, "cuda:0")
        out = torch.full((1, 2, 1) + img_size, 7.0, device="cuda:0")
        assert st.stage([vols], out=out) is out
        torch.cuda.synchronize()
        want = O.stage_batch([raws], img_size)
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
        win = [min(a, b) for a, b in zip(dims, img_size)]
        assert st.h2d_bytes < 2 * int(np.prod(win)) * 2 + 4096        # stored int16 bytes of the crop window only


def test_bad_arguments_fail_loudly():
    from cavit import CavitError
    from cavit.staging import RawVolume, VolumeStager
    st = VolumeStager((4, 4, 4), "cuda:0")
    v = RawVolume(np.zeros(8, np.int16), (2, 2, 2))
    with pytest.raises(CavitError):
        st.stage([])
    with pytest.raises(CavitError):
        st.stage([[v, v], [v]])
    with pytest.raises(CavitError):
        st.stage([[v]], out=torch.empty(1, 1, 1, 4, 4, 5, device="cuda:0"))


@pytest.mark.parametrize("name", ["cross_ring4", "cross_heads3"])
def test_files_to_logits_and_gradients_match_oracle(tmp_path, name):
    """The whole input side of a training step from FILES: .nii.gz -> read_nifti -> VolumeStager -> ModelCross forward +
    backward, against the oracle chain (oracle/staging.py -> oracle/functional.py, fp64). Volumes are cropped along one axis
    and padded along another; the staged batch is bit-exact, logits / gradients within the model tolerances."""
    from oracle import functional as OF
    from oracle.cases import build_case
    from cavit.modules import ModelCross
    from cavit.staging import VolumeStager, read_nifti
    kind, cfg, state, img0, labels = build_case(name)
    B, M = img0.shape[:2]
    D, H, W = cfg.img_size
    rng = np.random.default_rng(11)
    samples, raws = [], []
    for b in range(B):
        vs, rs = [], []
        for m in range(M):
            dims = (D + 5 + m, max(H - 6, 1), W)                  # crop i, pad j, keep k
            arr = rng.integers(-3000, 3000, size=dims).astype(np.int16)
            p = str(tmp_path / f"s{b}_m{m}.nii.gz")
            O.write_nifti(p, arr, slope=1.0 / 1024, inter=0.25 * m, extension_bytes=2896)
            v = read_nifti(p)
            vs.append(v)
            rs.append((v.data, v.dims, v.slope, v.inter))
        samples.append(vs)
        raws.append(rs)
    x = VolumeStager(cfg.img_size, "cuda:0").stage(samples)
    want = torch.from_numpy(O.stage_batch(raws, cfg.img_size))
    assert torch.equal(x.cpu(), want)
    model = ModelCross(cfg)
    model.load_state_dict(state)
    model = model.cuda().train()
    logits, loss = model(x, labels.cuda())
    loss.backward()
    ref_logits, ref_loss, ref_grads = OF.forward_backward(state, want, labels, cfg, kind, torch.float64)
    err = float((logits.detach().double().cpu() - ref_logits).norm() / ref_logits.norm())
    assert err < 2e-2, err
    num = sum(float((p.grad.double().cpu() - ref_grads[k]).norm()) ** 2 for k, p in model.named_parameters())
    den = sum(float(g.norm()) ** 2 for g in ref_grads.values())
    assert (num / den) ** 0.5 < 3e-2, (num / den) ** 0.5
