"""Multi-GPU: the sharded step over NCCL equals the single-GPU step.  Needs >= 2 GPUs on
the box (skipped otherwise); launched the way the driver launches bench.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_total", [1_000_003, 20_000_000])
def test_sharded_step_equals_single_gpu(n_total):
    import torch

    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    g = 2 if g < 4 else (4 if g < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={g}", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_check.py"), str(n_total)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0
