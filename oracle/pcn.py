"""pCN mutation sweep - numpy spec of ``asp_pcn_sweep`` (include/aspire_b200.h).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED.  The reference delegates the step arithmetic to the
third-party package ``minipcn`` (``minipcn[array-api]>=0.2.0a3``,
``pyproject.toml:43-46``; imported lazily at
``src/aspire/samplers/smc/minipcn.py:89``), which is neither under
``/root/reference`` nor installed in this image, and no reference test checks a
pCN numeric result (``tests/test_samplers/test_mcmc/test_checkpointing.py:88-109``
only fakes the call signature).  What IS pinned to the reference:

* the tempered target, ``src/aspire/samplers/smc/base.py:507-519`` +
  ``src/aspire/samples.py:1217-1219``:
  ``log pi_beta(x) = (1-beta) log_q(x) + beta (logL(x) + logP(x))``, NaN -> -inf;
* the wrapper contract, ``src/aspire/samplers/smc/minipcn.py:97-134``: run
  ``n_steps`` moves, keep the final ensemble, record the mean acceptance,
  return log_q / log_prior / log_likelihood at the final state.

The step itself is the standard preconditioned Crank-Nicolson proposal
(Cotter et al. 2013) with a diagonal Gaussian reference measure N(mu, sigma^2)
fitted to the ensemble:

    x'      = mu + sqrt(1 - rho^2) (x - mu) + rho sigma xi,    xi ~ N(0, I)
    log a   = [log pi(x') - log pi(x)] + [r(x') - r(x)],
              r(y) = 1/2 sum_j ((y_j - mu_j) / sigma_j)^2
    accept  iff log u < log a,   u ~ U(0, 1]

which leaves pi invariant (the proposal is reversible w.r.t. N(mu, sigma^2)).
"""

from __future__ import annotations

import numpy as np

from . import densities, philox


def tempered_target(log_l, log_p, log_q, beta):
    """src/aspire/samples.py:1217-1219 + NaN guard of smc/base.py:518."""
    t = (1.0 - beta) * log_q + beta * (log_l + log_p)
    return np.where(np.isnan(t), -np.inf, t)


def ref_quad(x, mu, inv_sigma):
    z = (x - mu) * inv_sigma
    return 0.5 * np.sum(z * z, axis=-1)


def propose(x, mu, sigma, rho, xi):
    a = np.sqrt(1.0 - rho * rho)
    return mu + (a * (x - mu) + (rho * sigma) * xi)


def sweep(x, log_l, log_p, log_q, *, like, prior, q, beta, mu, sigma, rho, n_moves,
          xi=None, log_u=None, seed=None, stream0=0, index0=0, index=None):
    """Run ``n_moves`` pCN moves.  Noise is either supplied (``xi`` [M,N,D],
    ``log_u`` [M,N]; parity mode) or generated with Philox (``seed``; move t
    uses stream ``stream0 + t``, particle i uses element index ``index0 + i``, or
    ``index[i]`` when per-particle output slots are given).
    Returns (x, log_l, log_p, log_q, accepts_per_move)."""
    x = np.array(x, dtype=np.float64)
    log_l = np.array(log_l, dtype=np.float64)
    log_p = np.array(log_p, dtype=np.float64)
    log_q = np.array(log_q, dtype=np.float64)
    n, d = x.shape
    inv_sigma = 1.0 / sigma
    accepts = np.zeros(n_moves, dtype=np.int64)
    index = (np.arange(index0, index0 + n, dtype=np.uint64) if index is None
             else np.asarray(index, dtype=np.uint64))
    for t in range(n_moves):
        if xi is not None:
            z = xi[t]
            lu = log_u[t]
        else:
            z = philox.normals_f32(seed, stream0 + t, index, d).astype(np.float64)
            lu = philox.accept_log_uniform(seed, stream0, t, index)
        xp = propose(x, mu, sigma, rho, z)
        lq2 = densities.evaluate(q, xp)
        lp2 = densities.evaluate(prior, xp)
        ll2 = densities.evaluate(like, xp)
        cur = tempered_target(log_l, log_p, log_q, beta)
        new = tempered_target(ll2, lp2, lq2, beta)
        with np.errstate(invalid="ignore"):
            log_a = (new - cur) + (ref_quad(xp, mu, inv_sigma) - ref_quad(x, mu, inv_sigma))
        acc = lu < log_a
        accepts[t] = int(acc.sum())
        x[acc] = xp[acc]
        log_l[acc] = ll2[acc]
        log_p[acc] = lp2[acc]
        log_q[acc] = lq2[acc]
    return x, log_l, log_p, log_q, accepts


def adapt_rho(rho: float, acceptance: float, target: float = 0.234,
              rho_min: float = 1e-3, rho_max: float = 1.0) -> float:
    """Between-sweep step-size adaptation (ours): multiplicative
    Robbins-Monro step in log rho towards the target acceptance."""
    return float(min(max(rho * np.exp(acceptance - target), rho_min), rho_max))


def weighted_moments(x, w):
    """Reference-measure fit: importance-weighted mean and std of the
    ensemble BEFORE resampling (weights = the resampling weights)."""
    sw = np.sum(w)
    mu = np.sum(w[:, None] * x, axis=0) / sw
    var = np.sum(w[:, None] * (x * x), axis=0) / sw - mu * mu   # one-pass form, as the kernel
    return mu, np.sqrt(np.maximum(var, 0.0))
