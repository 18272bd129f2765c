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


def test_ops_run_on_the_device_of_their_tensors():
    """``Aspire(device="cuda:1")`` without a ``torch.cuda.set_device``: every wrapper launches on the GPU that holds
    its tensors (stream, workspaces and kernels of that device), and gives what cuda:0 gives."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    from aspire_b200 import ops
    from aspire_b200 import targets as tg
    from aspire_b200._lib import AspireB200Error

    assert torch.cuda.current_device() == 0
    results = []
    for dev in (torch.device("cuda:0"), torch.device("cuda:1")):
        n, dims = 200_000, 32
        prob = tg.rosenbrock_problem(dims, device=dev)
        x = ops.philox_normal(5, 0, n, dims, dev, mean=prob["q"].mean, sigma=prob["q"].sigma)
        assert x.device == dev
        ll, lp, lq = prob["like"].log_prob(x), prob["prior"].log_prob(x), prob["q"].log_prob(x)
        d, stats = ops.reweight(ll, lp, lq)
        st = ops.beta_search(d, stats, 0.0, 0.5, 1e-6)
        dbeta = st.beta_min
        p, _, tiles = ops.smc_weights(d, dbeta, dbeta * st.d_max, st.s1_min)
        cdf = ops.cdf_exact(p, tile_stats=tiles)
        idx = ops.search_multinomial_sorted(cdf, 3, 1, n)
        mu, sigma = ops.moments_to_mu_sigma(ops.weighted_moments(x, idx=idx), dims, floor=1e-150)
        prm = ops.make_pcn_params(dbeta, 0.4, mu, sigma, 3, seed=7, stream0=2)
        out = ops.pcn_sweep(x, ll, lp, lq, idx, prob["like"], prob["prior"], prob["q"], prm)
        assert all(t.device == dev for t in out)
        results.append((st.beta_min, idx.cpu().numpy(), out[0].cpu().numpy(), out[4].cpu().numpy()))
        assert torch.cuda.current_device() == 0
    assert results[0][0] == results[1][0]
    np.testing.assert_array_equal(results[0][1], results[1][1])
    np.testing.assert_array_equal(results[0][2], results[1][2])
    np.testing.assert_array_equal(results[0][3], results[1][3])
    a = torch.zeros(8, dtype=torch.float64, device="cuda:0")
    b = torch.zeros(8, dtype=torch.float64, device="cuda:1")
    with pytest.raises(AspireB200Error):
        ops.reweight(a, a, b)
