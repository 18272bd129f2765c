"""Proposal flows behind the reference's ``Flow`` plug-in contract
(``src/aspire/flows/base.py:11-98``: ``log_prob``, ``sample``,
``sample_and_log_prob``, ``fit``, class attribute ``xp``).

``DiagGaussianFlow`` is the analytic diagonal-Gaussian proposal of the benchmark
configs ("q = diagonal Gaussian fitted to 2,000 draws"); its density
is a built-in ``asp_density_t`` so the fused pCN sweep can evaluate ``log_q``
in-kernel.  Any other object with the same methods works through the unfused path.
"""

from __future__ import annotations

import numpy as np
import torch

from . import ops
from .history import FlowHistory
from .rng import as_uniform_source
from .targets import GaussianDensity


class Flow:
    """Minimal mirror of the reference ``Flow`` base class."""

    xp = torch

    def __init__(self, dims: int, device=None, data_transform=None, dtype=None):
        self.dims = int(dims)
        self.device = torch.device(device if device is not None else "cuda")
        self.data_transform = data_transform
        self.dtype = torch.float64

    def log_prob(self, x):
        raise NotImplementedError

    def sample(self, n_samples):
        return self.sample_and_log_prob(n_samples)[0]

    def sample_and_log_prob(self, n_samples):
        raise NotImplementedError

    def fit(self, samples, **kwargs) -> FlowHistory:
        raise NotImplementedError


class DiagGaussianFlow(Flow):
    """q(x) = N(mean, diag(sigma^2)); ``fit`` matches moments (population std)."""

    def __init__(self, dims: int, device=None, data_transform=None, dtype=None, mean=None, sigma=None, rng=None):
        super().__init__(dims, device, data_transform, dtype)
        self.rng = as_uniform_source(rng)
        self.density = None
        if mean is not None or sigma is not None:
            self._set(np.zeros(dims) if mean is None else mean, np.ones(dims) if sigma is None else sigma)

    def _set(self, mean, sigma):
        self.density = GaussianDensity(mean, sigma, self.dims, device=self.device)

    def fit(self, samples, **kwargs) -> FlowHistory:
        x = samples.x if hasattr(samples, "x") else samples
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device)
        if not x.is_contiguous():
            x = x.contiguous()
        m = ops.weighted_moments(x)
        mu, sigma = ops.moments_to_mu_sigma(m, self.dims)
        self._set(mu.cpu().numpy(), sigma.cpu().numpy())
        return FlowHistory()

    def _require(self):
        if self.density is None:
            raise RuntimeError("flow has not been fitted")

    def log_prob(self, x):
        self._require()
        return self.density.log_prob(torch.as_tensor(x, dtype=torch.float64, device=self.device))

    def sample_and_log_prob(self, n_samples):
        self._require()
        x = self.rng.normal(n_samples, self.dims, self.device, mean=self.density.mean, sigma=self.density.sigma)
        return x, self.density.log_prob(x)
