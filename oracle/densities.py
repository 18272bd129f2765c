"""numpy spec of the built-in synthetic log-densities (benchmark targets).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference has no
built-in densities: users pass Python callables (``src/aspire/samplers/base.py:81-91``).
These follow the reference's own fixtures where one exists - the Gaussian of
``README.md:47-52`` / ``tests/integration_tests/conftest.py:84-108`` and the
mixture of ``examples/smc_example.py:40-58`` - plus the Rosenbrock of SURVEY 8d.
Parameterisation = the C-ABI's ``asp_density_t`` (include/aspire_b200.h).
"""

from __future__ import annotations

import math

import numpy as np

KIND_GAUSSIAN = 0
KIND_ROSENBROCK = 1
KIND_GMM = 2
KIND_UNIFORM = 3

LOG_2PI = math.log(2.0 * math.pi)


def gaussian(x, mean, inv_sigma, const):
    """``const - 0.5 * sum_j ((x_j - mean_j) * inv_sigma_j)**2``"""
    z = (x - mean) * inv_sigma
    return const - 0.5 * np.sum(z * z, axis=-1)


def gaussian_const(sigma, normalised=True):
    sigma = np.asarray(sigma, dtype=np.float64)
    if not normalised:
        return 0.0
    return float(-np.sum(np.log(sigma)) - 0.5 * sigma.size * LOG_2PI)


def rosenbrock(x, a=100.0, b=1.0, scale=20.0):
    """``-sum_{j<D-1} [a (x_{j+1} - x_j^2)^2 + (b - x_j)^2] / scale``"""
    x0 = x[..., :-1]
    x1 = x[..., 1:]
    t = x1 - x0 * x0
    s = b - x0
    return -np.sum(a * t * t + s * s, axis=-1) / scale


def gmm(x, means, inv_sigma, logc):
    """``logsumexp_k [ logc_k - 0.5 * inv_sigma_k^2 * |x - m_k|^2 ]`` with
    isotropic components; ``logc_k = log w_k - D log sigma_k - D/2 log 2pi``."""
    d2 = np.sum((x[:, None, :] - means[None, :, :]) ** 2, axis=-1)  # [N,K]
    t = logc[None, :] - 0.5 * (inv_sigma[None, :] ** 2) * d2
    m = np.max(t, axis=-1)
    return m + np.log(np.sum(np.exp(t - m[:, None]), axis=-1))


def gmm_logc(weights, sigma, dims):
    weights = np.asarray(weights, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    return np.log(weights) - dims * np.log(sigma) - 0.5 * dims * LOG_2PI


def uniform_box(x, lo, hi, const):
    """``const`` inside the open box, ``-inf`` outside."""
    inside = np.all((x > lo) & (x < hi), axis=-1)
    return np.where(inside, const, -np.inf)


def evaluate(desc: dict, x: np.ndarray) -> np.ndarray:
    """Evaluate a density given as the dict form of ``asp_density_t``."""
    kind = desc["kind"]
    if kind == KIND_GAUSSIAN:
        return gaussian(x, desc["mean"], desc["inv_sigma"], desc["const"])
    if kind == KIND_ROSENBROCK:
        return rosenbrock(x, desc["a"], desc["b"], desc["scale"])
    if kind == KIND_GMM:
        return gmm(x, desc["means"], desc["inv_sigma"], desc["logc"])
    if kind == KIND_UNIFORM:
        return uniform_box(x, desc["lo"], desc["hi"], desc["const"])
    if kind == "callable":  # a proposal that only offers log_prob (a flow): desc["fn"](x) -> [n]
        return np.asarray(desc["fn"](x), dtype=np.float64)
    raise ValueError(f"unknown density kind {kind}")
