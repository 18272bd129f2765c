#!/usr/bin/env python
"""A learned coupling flow as the SMC proposal on a multi-modal target, with the whole mutation in one kernel.

    python examples/flow_proposal.py

16-D mixture of two Gaussians under a broad Gaussian prior.  The flow is fitted on the device to draws from a crude
envelope (``Aspire.fit``: forward / backward / Adam kernels), then ``sample_posterior(sampler="smc")`` tempers from the
flow to the posterior.  Likelihood and prior are built-in densities (``aspire_b200.targets``) and the step is plain pCN,
so every mutation runs in ``asp_pcn_sweep_flow``: resampling gather, proposal, mixture likelihood, log q(x') through the
flow's coupling layers on the tensor cores and the accept test, particle by particle, in one launch.  With Python
callables instead of built-in densities the same run goes through the per-move kernels (see examples/quickstart.py).
"""
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))  # run from a checkout

import numpy as np
import torch

from aspire_b200 import Aspire, Samples
from aspire_b200 import targets as tg
from aspire_b200.rng import PhiloxRNG

DIMS = 16


def main(n_samples: int = 200_000):
    dev = torch.device("cuda")
    means = np.zeros((2, DIMS))
    means[0, :4], means[1, :4] = 1.5, -1.5
    like = tg.GaussianMixtureDensity(means, 0.5, device=dev)
    prior = tg.GaussianDensity(0.0, 4.0, DIMS, device=dev)
    aspire = Aspire(log_likelihood=like, log_prior=prior, dims=DIMS, parameters=[f"x{i}" for i in range(DIMS)],
                    flow_class="coupling", rng=PhiloxRNG(1), device="cuda")
    init = Samples(PhiloxRNG(2).normal(20_000, DIMS, dev) * 2.0)
    history = aspire.fit(init, n_epochs=20, batch_size=500, lr=2e-3)
    posterior, smc = aspire.sample_posterior(sampler="smc", n_samples=n_samples, return_history=True, rng=PhiloxRNG(3),
                                             store_sample_history=False,
                                             sampler_kwargs=dict(n_steps=10, step_fn="pcn"))
    # evidence: the likelihood is a normalised mixture, so Z = mean_k N(m_k; 0, (0.5^2 + 4^2) I)
    var = 0.25 + 16.0
    log_z = math.log(0.5) + np.logaddexp(*(-0.5 * (m @ m) / var for m in means)) - 0.5 * DIMS * math.log(2 * math.pi * var)
    x = posterior.x
    left = float((x[:, 0] < 0).double().mean())
    print(f"flow fit: loss {history.training_loss[0]:.2f} -> {history.training_loss[-1]:.2f}")
    print(f"{len(smc.beta)} tempering steps, acceptance {np.mean(smc.mcmc_acceptance):.2f}")
    print(f"log evidence {posterior.log_evidence:.3f} +/- {posterior.log_evidence_error:.3f} (analytic {log_z:.3f})")
    print(f"mass of the mode at -1.5: {left:.3f} (0.5)")
    return posterior, log_z, left


if __name__ == "__main__":
    main()
