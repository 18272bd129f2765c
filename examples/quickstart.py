#!/usr/bin/env python
"""The reference's README quickstart (README.md:42-77) on aspire_b200: same calls, torch tensors on the GPU.

    python examples/quickstart.py

The callables receive a ``Samples`` object whose ``x`` is a CUDA tensor (the reference passes arrays of the
namespace ``xp``; here ``xp`` is torch).  With Python callables the pCN moves run unfused (one proposal kernel,
the callables, one accept kernel per move); pass ``aspire_b200.targets`` densities instead to get the fused sweep.
"""
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))  # run from a checkout

import torch

from aspire_b200 import Aspire, Samples


def log_likelihood(samples):
    return -0.5 * torch.sum(samples.x**2, dim=-1)


def log_prior(samples):
    return -0.5 * torch.sum(samples.x**2, dim=-1)


def main():
    init = Samples(torch.randn(2_000, 4, dtype=torch.float64, device="cuda"))
    aspire = Aspire(log_likelihood=log_likelihood, log_prior=log_prior, dims=4,
                    parameters=[f"x{i}" for i in range(4)])
    aspire.fit(init, n_epochs=20)
    posterior = aspire.sample_posterior(sampler="smc", n_samples=500, sampler_kwargs=dict(n_steps=100))
    print(f"log evidence {posterior.log_evidence:.3f} +/- {posterior.log_evidence_error:.3f} "
          f"(analytic {2 * math.log(math.pi):.3f})")
    print("posterior mean", posterior.x.mean(dim=0).tolist())
    print("posterior var ", posterior.x.var(dim=0).tolist(), "(analytic 0.5)")


if __name__ == "__main__":
    main()
