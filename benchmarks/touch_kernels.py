#!/usr/bin/env python
"""Launch every secondary kernel once at a profiling-friendly size (for `ncu --set full`): the three-level
bisection pass, the streaming systematic search, the semi-fused move, the tpCN scales, device PCG64."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from aspire_b200 import ops  # noqa: E402
from aspire_b200 import targets as tg  # noqa: E402

dev = torch.device("cuda")
n, dims = int(os.environ.get("N", 20_000_000)), 32
pg = tg.rosenbrock_problem(dims, device=dev)
x = ops.philox_normal(5, 0, n, dims, dev, mean=pg["q"].mean, sigma=pg["q"].sigma)
ll, lp, lq = pg["like"].log_prob(x), pg["prior"].log_prob(x), pg["q"].log_prob(x)
for _ in range(2):
    d, stats = ops.reweight(ll, lp, lq)
    state = ops.beta_search(d, stats, 0.0, 0.5, 1e-6)
    beta = state.beta_min
    st = stats.cpu().numpy()
    p, result, tiles = ops.smc_weights(d, beta, beta * st[1], state.s1_min)
    cdf = ops.cdf_exact(p, tile_stats=tiles)
    idx_s = ops.search_systematic(cdf, 0.37, n)
    rng = np.random.default_rng(1)
    s = rng.bit_generator.state["state"]
    u = ops.pcg64_uniform(s["state"], s["inc"], n, dev)
    idx = ops.search_multinomial(cdf, u)
    mu, sigma = ops.moments_to_mu_sigma(ops.weighted_moments(x[: 1 << 20]), dims)
    prm = ops.make_pcn_params(beta, 0.4, mu, sigma, 1, seed=3, stream0=1)
    xp, ll2, lp2, dr = ops.pcn_propose_eval(x, pg["like"], pg["prior"], prm, 0)
    acc = torch.zeros(1, dtype=torch.int64, device=dev)
    ops.pcn_accept_scalar(x, ll, lp, lq, xp, ll2, lp2, pg["q"].log_prob(xp), dr, prm, 0, acc)
    prm_t = ops.make_pcn_params(beta, 0.4, mu, sigma, 1, seed=3, stream0=1, nu=6.0)
    ops.tpcn_scales(x, prm_t, 0)
torch.cuda.synchronize()
print("ok", float(acc[0]) / n)
