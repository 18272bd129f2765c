"""Built-in log-densities (benchmark targets) as objects the samplers understand.

The reference takes arbitrary Python callables ``log_likelihood(samples)`` /
``log_prior(samples)`` (``src/aspire/samplers/base.py:81-91``).  These classes
are such callables - ``density(samples)`` returns a ``[N]`` CUDA tensor computed
by ``asp_logdensity`` - and additionally carry the ``asp_density_t`` descriptor
that lets the SMC sampler run its fused pCN sweep without leaving the device.
Problem shapes follow SURVEY 8d (C1-C4).
"""

from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

LOG_2PI = math.log(2.0 * math.pi)


def _dev_tensor(a, device):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), dtype=torch.float64).to(device)


class Density:
    """Base: a log-density with an ``asp_density_t`` descriptor."""

    kind: int = -1

    def __init__(self, dims: int, device=None):
        self.dims = int(dims)
        self.device = torch.device(device if device is not None else _lib.DEFAULT_DEVICE)
        self._struct = None

    # -- descriptor ---------------------------------------------------------
    def _fill(self, s: _lib.Density) -> None:
        raise NotImplementedError

    def struct(self) -> _lib.Density:
        if self._struct is None:
            s = _lib.Density()
            s.kind = self.kind
            s.dims = self.dims
            s.n_comp = 0
            s.p0 = s.p1 = s.p2 = None
            s.s0 = s.s1 = s.s2 = 0.0
            self._fill(s)
            self._struct = s
        return self._struct

    def to(self, device):
        raise NotImplementedError

    def describe(self) -> dict:
        """Plain-numpy description (same keys as ``oracle/densities.py``)."""
        raise NotImplementedError

    # -- callable, reference style -----------------------------------------
    def log_prob(self, x: torch.Tensor) -> torch.Tensor:
        from . import ops

        return ops.logdensity(self, x)

    def __call__(self, samples) -> torch.Tensor:
        x = samples.x if hasattr(samples, "x") else samples
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device)
        return self.log_prob(x)


class GaussianDensity(Density):
    """``const - 1/2 sum ((x - mean) / sigma)^2``; normalised unless told otherwise."""

    kind = _lib.DENSITY_GAUSSIAN

    def __init__(self, mean, sigma, dims=None, normalised=True, device=None):
        mean = np.asarray(mean, dtype=np.float64)
        sigma = np.asarray(sigma, dtype=np.float64)
        dims = int(dims if dims is not None else max(mean.size, sigma.size))
        super().__init__(dims, device)
        self.mean_np = np.broadcast_to(mean, (dims,)).copy()
        self.sigma_np = np.broadcast_to(sigma, (dims,)).copy()
        self.normalised = bool(normalised)
        self.const = float(-np.sum(np.log(self.sigma_np)) - 0.5 * dims * LOG_2PI) if normalised else 0.0
        self.mean = _dev_tensor(self.mean_np, self.device)
        self.sigma = _dev_tensor(self.sigma_np, self.device)
        self.inv_sigma = _dev_tensor(1.0 / self.sigma_np, self.device)

    def _fill(self, s):
        s.p0 = self.mean.data_ptr()
        s.p1 = self.inv_sigma.data_ptr()
        s.s0 = self.const

    def to(self, device):
        return GaussianDensity(self.mean_np, self.sigma_np, self.dims, self.normalised, device)

    def describe(self):
        return dict(kind=self.kind, mean=self.mean_np, sigma=self.sigma_np, inv_sigma=1.0 / self.sigma_np,
                    const=self.const)


class RosenbrockDensity(Density):
    """``-sum_{j<D-1} [a (x_{j+1} - x_j^2)^2 + (b - x_j)^2] / scale``"""

    kind = _lib.DENSITY_ROSENBROCK

    def __init__(self, dims, a=100.0, b=1.0, scale=20.0, device=None):
        super().__init__(dims, device)
        self.a, self.b, self.scale = float(a), float(b), float(scale)

    def _fill(self, s):
        s.s0, s.s1, s.s2 = self.a, self.b, self.scale

    def to(self, device):
        return RosenbrockDensity(self.dims, self.a, self.b, self.scale, device)

    def describe(self):
        return dict(kind=self.kind, a=self.a, b=self.b, scale=self.scale)


class GaussianMixtureDensity(Density):
    """Isotropic components: ``logsumexp_k [log w_k + log N(x; m_k, sigma_k^2 I)]``."""

    kind = _lib.DENSITY_GMM

    def __init__(self, means, sigma, weights=None, device=None):
        means = np.asarray(means, dtype=np.float64)
        n_comp, dims = means.shape
        if n_comp > 32:
            raise ValueError("at most 32 mixture components")
        super().__init__(dims, device)
        self.means_np = means.copy()
        self.sigma_np = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (n_comp,)).copy()
        w = np.full(n_comp, 1.0 / n_comp) if weights is None else np.asarray(weights, dtype=np.float64)
        self.weights_np = w / w.sum()
        self.logc_np = np.log(self.weights_np) - dims * np.log(self.sigma_np) - 0.5 * dims * LOG_2PI
        self.n_comp = n_comp
        self.means = _dev_tensor(self.means_np, self.device)
        self.inv_sigma = _dev_tensor(1.0 / self.sigma_np, self.device)
        self.logc = _dev_tensor(self.logc_np, self.device)

    def _fill(self, s):
        s.n_comp = self.n_comp
        s.p0 = self.means.data_ptr()
        s.p1 = self.inv_sigma.data_ptr()
        s.p2 = self.logc.data_ptr()

    def to(self, device):
        return GaussianMixtureDensity(self.means_np, self.sigma_np, self.weights_np, device)

    def describe(self):
        return dict(kind=self.kind, means=self.means_np, inv_sigma=1.0 / self.sigma_np, logc=self.logc_np)


class UniformDensity(Density):
    """Constant inside the open box ``(lo, hi)``, ``-inf`` outside."""

    kind = _lib.DENSITY_UNIFORM

    def __init__(self, lo, hi, dims=None, normalised=True, device=None):
        lo = np.asarray(lo, dtype=np.float64)
        hi = np.asarray(hi, dtype=np.float64)
        dims = int(dims if dims is not None else max(lo.size, hi.size))
        super().__init__(dims, device)
        self.lo_np = np.broadcast_to(lo, (dims,)).copy()
        self.hi_np = np.broadcast_to(hi, (dims,)).copy()
        self.normalised = bool(normalised)
        self.const = float(-np.sum(np.log(self.hi_np - self.lo_np))) if normalised else 0.0
        self.lo = _dev_tensor(self.lo_np, self.device)
        self.hi = _dev_tensor(self.hi_np, self.device)

    def _fill(self, s):
        s.p0 = self.lo.data_ptr()
        s.p1 = self.hi.data_ptr()
        s.s0 = self.const

    def to(self, device):
        return UniformDensity(self.lo_np, self.hi_np, self.dims, self.normalised, device)

    def describe(self):
        return dict(kind=self.kind, lo=self.lo_np, hi=self.hi_np, const=self.const)


# ---- the benchmark problems of SURVEY 8d ----------------------------------
def gaussian_problem(dims=16, device=None):
    """C2: prior N(0, I), likelihood N(0.5, 0.7^2 I), proposal N(0.3, 1.2^2 I)."""
    return dict(dims=dims, like=GaussianDensity(0.5, 0.7, dims, device=device),
                prior=GaussianDensity(0.0, 1.0, dims, device=device),
                q=GaussianDensity(0.3, 1.2, dims, device=device))


def quickstart_problem(dims=4, device=None):
    """C1 (reference README.md:42-77): both callables are the unnormalised -x.x/2."""
    return dict(dims=dims, like=GaussianDensity(0.0, 1.0, dims, normalised=False, device=device),
                prior=GaussianDensity(0.0, 1.0, dims, normalised=False, device=device),
                q=GaussianDensity(0.0, 1.0, dims, device=device))


def rosenbrock_problem(dims=32, q_mean=0.5, q_sigma=1.5, device=None):
    """C3: Rosenbrock/20 likelihood, prior N(0, 3^2 I), diagonal-Gaussian proposal."""
    return dict(dims=dims, like=RosenbrockDensity(dims, device=device),
                prior=GaussianDensity(0.0, 3.0, dims, device=device),
                q=GaussianDensity(q_mean, q_sigma, dims, device=device))


def gmm_problem(dims=64, n_comp=8, seed=20251104, device=None):
    """C4: 8 isotropic components, means +-2 on a seeded sign pattern, sigma 0.7."""
    rng = np.random.default_rng(seed)
    means = 2.0 * (2.0 * rng.integers(0, 2, size=(n_comp, dims)) - 1.0)
    return dict(dims=dims, like=GaussianMixtureDensity(means, 0.7, device=device),
                prior=GaussianDensity(0.0, 4.0, dims, device=device),
                q=GaussianDensity(0.0, 2.5, dims, device=device))
