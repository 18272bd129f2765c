#!/usr/bin/env python
"""Bounded and periodic parameters on aspire_b200 - the calls a user of the reference makes
(`Aspire(prior_bounds=..., periodic_parameters=...)`, `sample_posterior(preconditioning=...)`).

    python examples/bounded_parameters.py

A 3-parameter toy: an amplitude in [0, 1] (Gaussian likelihood squeezed against the upper bound), a phase on
[0, 2 pi) (von Mises likelihood whose mode sits next to the seam) and a free offset.  With `prior_bounds` given the
proposal flow lives on the logit image of the bounded amplitude (no proposal mass outside the bounds), and
`preconditioning_kwargs` make the pCN moves run in the logit + periodic + standardised space, so every particle
stays inside the bounds and the phase wraps around.  Callables receive a `Samples` whose `x` is a torch tensor.
"""
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))  # run from a checkout

import torch

from aspire_b200 import Aspire, Samples

TWO_PI = 2 * math.pi
PARAMETERS = ["amplitude", "phase", "offset"]
PRIOR_BOUNDS = {"amplitude": [0.0, 1.0], "phase": [0.0, TWO_PI], "offset": [-math.inf, math.inf]}


def log_likelihood(samples):
    a, phi, c = samples.x[:, 0], samples.x[:, 1], samples.x[:, 2]
    return -0.5 * ((a - 0.95) / 0.1) ** 2 + 2.0 * torch.cos(phi - 0.3) - 0.5 * c**2


def log_prior(samples):
    a, phi, c = samples.x[:, 0], samples.x[:, 1], samples.x[:, 2]
    inside = (a > 0) & (a < 1) & (phi >= 0) & (phi < TWO_PI)
    lp = -math.log(TWO_PI) - 0.5 * (c / 3.0) ** 2 - math.log(3.0 * math.sqrt(TWO_PI))
    return torch.where(inside, lp, torch.full_like(lp, -math.inf))


def main(device="cuda", n_samples=20_000):
    u = torch.rand(4_000, 3, dtype=torch.float64, device=device)
    init = Samples(torch.stack([u[:, 0], u[:, 1] * TWO_PI, (u[:, 2] - 0.5) * 6.0], dim=1))
    aspire = Aspire(log_likelihood=log_likelihood, log_prior=log_prior, dims=3, parameters=PARAMETERS,
                    prior_bounds=PRIOR_BOUNDS, periodic_parameters=["phase"], bounded_to_unbounded=True,
                    bounded_transform="logit", device=device)
    aspire.fit(init)
    posterior, history = aspire.sample_posterior(
        n_samples=n_samples, sampler="smc", return_history=True, store_sample_history=False,
        preconditioning="standard", preconditioning_kwargs=dict(bounded_to_unbounded=True, affine_transform=True),
        sampler_kwargs=dict(n_steps=15))
    x = posterior.x
    phase = math.atan2(float(torch.sin(x[:, 1]).mean()), float(torch.cos(x[:, 1]).mean()))
    print(f"{len(history.beta)} tempering steps, mean acceptance {sum(history.mcmc_acceptance) / len(history.beta):.2f}")
    print(f"log evidence {posterior.log_evidence:.3f} +/- {posterior.log_evidence_error:.3f}")
    print(f"amplitude in [{float(x[:, 0].min()):.4f}, {float(x[:, 0].max()):.4f}], mean {float(x[:, 0].mean()):.3f}")
    print(f"phase in [{float(x[:, 1].min()):.4f}, {float(x[:, 1].max()):.4f}), circular mean {phase:.3f} (mode 0.3)")
    print(f"offset mean {float(x[:, 2].mean()):.3f}, std {float(x[:, 2].std()):.3f} (analytic 0, {3 / math.sqrt(10):.3f})")
    return posterior, history


if __name__ == "__main__":
    main()
