"""2-rank NCCL data parallelism on real GPUs: all-reduced shard gradients == the 1-GPU global-batch gradients, for the
graph-captured overlapped slab all-reduce ("overlap") and the single post-backward all-reduce ("post"). Needs >= 2 GPUs
(skipped on the 1-GPU box; run with `gpurun --gpus 2`). The slab bookkeeping itself is covered on CPU with gloo
(tests/test_ddp_cpu.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("mode", ["overlap", "post"])
def test_two_rank_nccl_gradients_match_global_batch(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611" if mode == "overlap" else "29612", os.path.join(HERE, "ddp_nccl_worker.py"), mode]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    print(r.stdout[-4000:], r.stderr[-4000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
