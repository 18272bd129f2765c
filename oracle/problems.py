"""Synthetic benchmark problems of SURVEY 8d in numpy (test infrastructure only).

Each problem is a dict ``{"like": desc, "prior": desc, "q": desc, "dims": D}``
whose entries are the dict form of the C-ABI's ``asp_density_t``
(see ``oracle/densities.py``).  ``aspire_b200.targets`` builds the same
problems for the product; ``tests/test_targets.py`` checks they agree.
"""

from __future__ import annotations

import numpy as np

from . import densities as dn


def _gauss(mean, sigma, dims, normalised=True):
    mean = np.broadcast_to(np.asarray(mean, dtype=np.float64), (dims,)).copy()
    sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), (dims,)).copy()
    return dict(kind=dn.KIND_GAUSSIAN, mean=mean, inv_sigma=1.0 / sigma, sigma=sigma,
                const=dn.gaussian_const(sigma, normalised))


def gaussian_problem(dims=16):
    """C2: prior N(0, I), likelihood N(0.5, 0.7^2 I), proposal N(0.3, 1.2^2 I)."""
    return dict(dims=dims, like=_gauss(0.5, 0.7, dims), prior=_gauss(0.0, 1.0, dims),
                q=_gauss(0.3, 1.2, dims))


def quickstart_problem(dims=4):
    """C1 (README.md:42-77): both callables are the unnormalised -x.x/2."""
    return dict(dims=dims, like=_gauss(0.0, 1.0, dims, normalised=False),
                prior=_gauss(0.0, 1.0, dims, normalised=False), q=_gauss(0.0, 1.0, dims))


def rosenbrock_problem(dims=32, q_mean=None, q_sigma=None):
    """C3: Rosenbrock/20 likelihood, prior N(0, 3^2 I), diagonal-Gaussian proposal."""
    q_mean = 0.5 if q_mean is None else q_mean
    q_sigma = 1.5 if q_sigma is None else q_sigma
    return dict(dims=dims, like=dict(kind=dn.KIND_ROSENBROCK, a=100.0, b=1.0, scale=20.0),
                prior=_gauss(0.0, 3.0, dims), q=_gauss(q_mean, q_sigma, dims))


def gmm_problem(dims=64, n_comp=8, seed=20251104):
    """C4: 8 isotropic components, means +-2 on a seeded sign pattern, sigma 0.7."""
    rng = np.random.default_rng(seed)
    means = 2.0 * (2.0 * rng.integers(0, 2, size=(n_comp, dims)) - 1.0)
    sigma = np.full(n_comp, 0.7)
    w = np.full(n_comp, 1.0 / n_comp)
    like = dict(kind=dn.KIND_GMM, means=means, inv_sigma=1.0 / sigma,
                logc=dn.gmm_logc(w, sigma, dims))
    return dict(dims=dims, like=like, prior=_gauss(0.0, 4.0, dims), q=_gauss(0.0, 2.5, dims))


def box_problem(dims=4):
    """Bounded parameters (scope row 8f-4): unnormalised Gaussian likelihood inside the unit box, flat prior on the
    box, diagonal-Gaussian proposal centred in it (its draws outside the box are dropped by the samplers)."""
    mu, sigma = np.linspace(0.25, 0.8, dims), np.linspace(0.12, 0.3, dims)
    prior = dict(kind=dn.KIND_UNIFORM, lo=np.zeros(dims), hi=np.ones(dims), const=0.0)
    return dict(dims=dims, like=_gauss(mu, sigma, dims, normalised=False), prior=prior, q=_gauss(0.5, 0.22, dims))


def draw(problem, n, seed):
    """Draw ``n`` particles from the diagonal-Gaussian proposal with a numpy
    Generator and evaluate the three log-densities."""
    rng = np.random.default_rng(seed)
    q = problem["q"]
    x = q["mean"] + q["sigma"] * rng.standard_normal((n, problem["dims"]))
    log_q = dn.evaluate(q, x)
    log_p = dn.evaluate(problem["prior"], x)
    log_l = dn.evaluate(problem["like"], x)
    return x, log_l, log_p, log_q
