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
              rho_min: float = 1e-3, rho_max: float = 1.0 - 1e-6) -> float:
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


# --------------------------------------------------------------------------------------
# tpCN: the Student-t preconditioned Crank-Nicolson step (the reference's default
# ``step_fn="tpcn"``, ``src/aspire/samplers/smc/minipcn.py:48``).  PARITY UNPINNED like pCN
# above: minipcn is absent.  Restated from the published algorithm (Karamanis et al. 2022,
# pocoMC's t-preconditioned CN) with a DIAGONAL scale matrix, consistent with our pCN:
#
#   r2(x)  = sum_j ((x_j - mu_j) / sigma_j)^2
#   s      ~ InvGamma((D + nu)/2, (nu + r2(x))/2)  =  (nu + r2(x)) / (2 g),  g ~ Gamma((D + nu)/2, 1)
#   x'     = mu + sqrt(1 - rho^2) (x - mu) + rho sqrt(s) sigma xi,            xi ~ N(0, I)
#   log a  = [log pi(x') - log pi(x)] + (nu + D)/2 [log1p(r2(x')/nu) - log1p(r2(x)/nu)]
#   accept iff log u < log a
#
# (reversible with respect to the multivariate Student-t t_nu(mu, diag sigma^2)).
# --------------------------------------------------------------------------------------
GAMMA_BLOCK0 = 0xFFFFFF00  # Philox blocks GAMMA_BLOCK0 + it: iteration `it` of a move's gamma draw


def tpcn_scale(x, mu, sigma, nu, g):
    """sqrt(s) of the tpCN proposal for gamma variates ``g ~ Gamma((D + nu)/2, 1)``."""
    z = (x - mu) / sigma
    r2 = np.sum(z * z, axis=-1)
    return np.sqrt((nu + r2) / (2.0 * g))


def gamma_mt_f32(seed: int, stream: int, index, a: float, max_iter: int = 64) -> np.ndarray:
    """Gamma(a, 1), a >= 1, by Marsaglia & Tsang (2000) in fp32 with counter-based draws: iteration ``it``
    of element ``index`` uses Philox block GAMMA_BLOCK0 + it of ``stream``: words (0, 1) -> the first
    Box-Muller normal z, word 2 -> u = fma(r, 2^-32, 2^-33); v = (1 + c z)^3, accepted when v > 0 and
    ln u < z^2/2 + d - d v + d ln v  (d = a - 1/3, c = 1/sqrt(9 d)); returns d v."""
    index = np.asarray(index, dtype=np.uint64)
    d = np.float32(a - 1.0 / 3.0)
    c = np.float32(1.0) / np.sqrt(np.float32(9.0) * d)
    out = np.zeros(len(index), dtype=np.float64)
    todo = np.ones(len(index), dtype=bool)
    c0, c1 = philox._split(index)
    k0, k1 = philox._key(seed)
    for it in range(max_iter):
        if not todo.any():
            break
        r = philox.philox4x32(c0, c1, np.uint32(stream), np.uint32((GAMMA_BLOCK0 + it) & 0xFFFFFFFF), k0, k1,
                              rounds=philox.NOISE_ROUNDS)
        z, _ = philox.box_muller_f32(r[0], r[1])
        u = philox._u01_f32(r[2])
        t = (np.float32(1.0) + c * z).astype(np.float32)
        v = (t * t * t).astype(np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            ln_u = (np.float32(0.6931471805599453) * np.log2(u)).astype(np.float32)
            ln_v = (np.float32(0.6931471805599453) * np.log2(np.maximum(v, np.float32(1e-30)))).astype(np.float32)
            ok = (t > 0) & (ln_u < np.float32(0.5) * z * z + d - d * v + d * ln_v)
        take = todo & ok
        out[take] = (d * v).astype(np.float32)[take].astype(np.float64)
        todo &= ~ok
    out[todo] = float(d)  # never reached in practice (acceptance > 95 % per iteration)
    return out


def tpcn_sweep(x, log_l, log_p, log_q, *, like, prior, q, beta, mu, sigma, nu, rho, n_moves,
               xi=None, log_u=None, gamma=None, seed=None, stream0=0, index0=0, index=None):
    """``n_moves`` tpCN moves; noise supplied (``xi`` [M,N,D], ``log_u`` [M,N], ``gamma`` [M,N]) or from
    Philox as in :func:`sweep` plus :func:`gamma_mt_f32`.  -> (x, log_l, log_p, log_q, accepts_per_move)."""
    x = np.array(x, dtype=np.float64)
    log_l, log_p, log_q = (np.array(a, dtype=np.float64) for a in (log_l, log_p, log_q))
    n, d = x.shape
    a_c = np.sqrt(1.0 - rho * rho)
    accepts = np.zeros(n_moves, dtype=np.int64)
    index = (np.arange(index0, index0 + n, dtype=np.uint64) if index is None
             else np.asarray(index, dtype=np.uint64))
    half = 0.5 * (nu + d)
    for t in range(n_moves):
        if xi is not None:
            z, lu, g = xi[t], log_u[t], gamma[t]
        else:
            z = philox.normals_f32(seed, stream0 + t, index, d).astype(np.float64)
            lu = philox.accept_log_uniform(seed, stream0, t, index)
            g = gamma_mt_f32(seed, stream0 + t, index, half)
        sc = tpcn_scale(x, mu, sigma, nu, g)
        xp = mu + (a_c * (x - mu) + ((rho * sigma) * sc[:, None]) * z)
        lq2, lp2, ll2 = densities.evaluate(q, xp), densities.evaluate(prior, xp), densities.evaluate(like, xp)
        cur = tempered_target(log_l, log_p, log_q, beta)
        new = tempered_target(ll2, lp2, lq2, beta)
        r2 = 2.0 * ref_quad(x, mu, 1.0 / sigma)
        r2p = 2.0 * ref_quad(xp, mu, 1.0 / sigma)
        with np.errstate(invalid="ignore"):
            log_a = (new - cur) + half * (np.log1p(r2p / nu) - np.log1p(r2 / nu))
        acc = lu < log_a
        accepts[t] = int(acc.sum())
        x[acc], log_l[acc], log_p[acc], log_q[acc] = xp[acc], ll2[acc], lp2[acc], lq2[acc]
    return x, log_l, log_p, log_q, accepts


def fit_student_t_diag(x, n_iter: int = 30, nu0: float = 10.0, nu_min: float = 1.0, nu_max: float = 1e6):
    """EM fit of a multivariate Student-t with diagonal scale (Liu & Rubin 1995, ECME flavour):
    -> (mu [D], sigma [D], nu).  E step  w_i = (nu + D) / (nu + r2_i);  M step  mu = sum w x / sum w,
    sigma^2 = sum w (x - mu)^2 / N;  nu solves
    -psi(nu/2) + log(nu/2) + 1 + mean(log w - w) + psi((nu + D)/2) - log((nu + D)/2) = 0  (bisection in log nu)."""
    from scipy.special import digamma

    x = np.asarray(x, dtype=np.float64)
    n, d = x.shape
    mu, sigma, nu = x.mean(axis=0), x.std(axis=0), float(nu0)
    for _ in range(n_iter):
        z = (x - mu) / sigma
        r2 = np.sum(z * z, axis=1)
        w = (nu + d) / (nu + r2)
        mu = (w[:, None] * x).sum(axis=0) / w.sum()
        sigma = np.sqrt((w[:, None] * (x - mu) ** 2).sum(axis=0) / n)
        const = 1.0 + np.mean(np.log(w) - w) + digamma(0.5 * (nu + d)) - np.log(0.5 * (nu + d))

        def f(v):
            return -digamma(0.5 * v) + np.log(0.5 * v) + const

        lo, hi = nu_min, nu_max
        if f(hi) > 0:      # Gaussian-like data: the likelihood keeps growing with nu
            nu = hi
        elif f(lo) < 0:
            nu = lo
        else:
            for _ in range(60):
                mid = np.sqrt(lo * hi)
                if f(mid) > 0:
                    lo = mid
                else:
                    hi = mid
            nu = float(np.sqrt(lo * hi))
    return mu, sigma, nu
