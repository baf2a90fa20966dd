"""Staging kernel rate (tools; GPU): B x M stored int16 volumes 240 x 240 x 155 -> fp32 [B, M, 1, D, H, W], against the
host-side numpy chain of the reference's formulation (oracle/staging.py) for the same batch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cross-attention-vit_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cavit.staging import RawVolume, VolumeStager  # noqa: E402
from oracle import staging as O  # noqa: E402

B, M, dims = 8, 4, (240, 240, 155)
rng = np.random.default_rng(0)
base = rng.integers(-3000, 3000, size=int(np.prod(dims))).astype(np.int16)
for img_size, (slope, inter) in [((128, 128, 64), (0.05, 1645.09)), ((240, 240, 160), (0.05, 1645.09)),
                                 ((240, 240, 160), (1.0, 0.0))]:
    samples = [[RawVolume(np.roll(base, 17 * (b * M + m)), dims, slope, inter) for m in range(M)] for b in range(B)]
    st = VolumeStager(img_size, "cuda:0")
    out = st.stage(samples)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        st.stage(samples, out=out)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / 5
    # kernel alone (bytes already on the device): CUDA events on the current stream
    from cavit import ops
    V = B * M
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D, H, W = img_size
    e0.record()
    for _ in range(20):
        ops.stage_volumes(st._raw_dev, st._dev, out, volumes=V, D=D, H=H, W=W, pad_value=-1.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    win = [min(s, t) for s, t in zip(dims, img_size)]
    alg = V * (2 * np.prod(win) + 4 * np.prod(img_size))
    t0 = time.perf_counter()
    O.stage_batch([[(v.data, v.dims, v.slope, v.inter) for v in s] for s in samples[:2]], img_size)
    host = (time.perf_counter() - t0) * B / 2
    print(f"img_size {img_size} scaling {(slope, inter)}: kernel {ms:.3f} ms = {alg / ms / 1e6:.0f} GB/s algorithmic; "
          f"stage() incl. pack + H2D of {st.h2d_bytes / 1e6:.0f} MB stored bytes {e2e * 1e3:.1f} ms "
          f"({V / e2e:.0f} volumes/s); host numpy chain {host * 1e3:.0f} ms for the batch ({V / host:.0f} volumes/s) "
          f"+ fp32 H2D of {V * np.prod(img_size) * 4 / 1e6:.0f} MB")
