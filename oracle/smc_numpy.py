"""numpy restatement of the reference's own SMC / importance arithmetic.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Every function cites
the reference lines (relative to ``/root/reference``) it follows.  PINNED:
``tests/test_oracle_vs_reference.py`` runs it against the unmodified reference
imported through ``oracle/ref_shim`` (dev container only) and
``tests/test_oracle_golden.py`` against the vectors committed in
``tests/golden`` (everywhere).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


class BetaScheduleError(RuntimeError):
    """src/aspire/samplers/smc/base.py:26"""


# --------------------------------------------------------------------------
# src/aspire/utils.py
# --------------------------------------------------------------------------
def logsumexp(x: np.ndarray) -> float:
    """src/aspire/utils.py:248-255 - ``c + log(sum(exp(x - c)))``, c = global max."""
    c = x.max()
    return c + np.log(np.sum(np.exp(x - c)))


def effective_sample_size(log_w: np.ndarray) -> float:
    """src/aspire/utils.py:510-512 - ``exp(2 lse(lw) - lse(2 lw))``."""
    return np.exp(np.asarray(logsumexp(log_w) * 2 - logsumexp(log_w * 2)))


# --------------------------------------------------------------------------
# src/aspire/samples.py  (SMCSamples)
# --------------------------------------------------------------------------
def log_p_t(log_l, log_p, log_q, beta):
    """src/aspire/samples.py:1217-1219"""
    return (1 - beta) * log_q + beta * (log_l + log_p)


def unnormalized_log_weights(log_l, log_p, log_q, beta0, beta):
    """src/aspire/samples.py:1221-1224"""
    return (beta0 - beta) * log_q + (beta - beta0) * (log_l + log_p)


def log_evidence_ratio(log_l, log_p, log_q, beta0, beta):
    """src/aspire/samples.py:1226-1228"""
    lw = unnormalized_log_weights(log_l, log_p, log_q, beta0, beta)
    return logsumexp(lw) - math.log(len(lw))


def log_evidence_ratio_variance(log_l, log_p, log_q, beta0, beta):
    """src/aspire/samples.py:1230-1242 (delta method, population variance)."""
    lw = unnormalized_log_weights(log_l, log_p, log_q, beta0, beta)
    m = np.max(lw)
    u = np.exp(lw - m)
    mean_w = np.mean(u)
    var_w = np.var(u)
    return var_w / (len(lw) * (mean_w**2)) if mean_w != 0 else np.nan


def log_weights(log_l, log_p, log_q, beta0, beta):
    """src/aspire/samples.py:1244-1249 (NaN guard, then ADDS the normaliser -
    a constant shift, reproduced as is)."""
    lw = unnormalized_log_weights(log_l, log_p, log_q, beta0, beta)
    if np.isnan(lw).any():
        raise ValueError(f"Log weights contain NaN values for beta={beta}")
    ler = logsumexp(lw) - math.log(len(lw))
    return lw + ler


def resample_weights(log_l, log_p, log_q, beta0, beta):
    """The ``p`` handed to ``rng.choice`` - src/aspire/samples.py:1274-1277."""
    lw = log_weights(log_l, log_p, log_q, beta0, beta)
    return np.exp(lw - logsumexp(lw))


def choice_cdf(p: np.ndarray) -> np.ndarray:
    """numpy ``Generator.choice(..., p=p)``: ``cdf = p.cumsum(); cdf /= cdf[-1]``
    (sequential fp64 adds, IEEE division).  Called at src/aspire/samples.py:1278."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf


def choice_indices(p: np.ndarray, u: np.ndarray) -> np.ndarray:
    """numpy ``Generator.choice`` given its uniforms:
    ``cdf.searchsorted(u, side='right')`` as int64."""
    return choice_cdf(p).searchsorted(u, side="right").astype(np.int64)


def systematic_positions(n_out: int, u0: float) -> np.ndarray:
    """Systematic resampling positions ``(i + u0) / n`` (extension; the
    reference has only the multinomial resampler).  One fp64 add and
    one IEEE division per position, no fused multiply-add."""
    return (np.arange(n_out, dtype=np.float64) + np.float64(u0)) / np.float64(n_out)


def systematic_indices(p: np.ndarray, u0: float, n_out: int | None = None) -> np.ndarray:
    n_out = len(p) if n_out is None else n_out
    idx = choice_cdf(p).searchsorted(systematic_positions(n_out, u0), side="right")
    return np.minimum(idx, len(p) - 1).astype(np.int64)


def resample(x, log_l, log_p, log_q, beta0, beta, rng, n_samples=None):
    """src/aspire/samples.py:1251-1287.  Returns (idx, x, log_l, log_p, log_q);
    idx is None for the same-beta shortcut (the reference returns ``self``)."""
    n = len(x)
    if n_samples is None:
        n_samples = n
    if beta == beta0:
        if n_samples == n:
            return None, x, log_l, log_p, log_q
        lw = np.zeros(n)
    else:
        lw = log_weights(log_l, log_p, log_q, beta0, beta)
    w = np.exp(lw - logsumexp(lw))
    idx = rng.choice(n, size=n_samples, replace=True, p=w)
    return idx, x[idx], log_l[idx], log_p[idx], log_q[idx]


# --------------------------------------------------------------------------
# src/aspire/samples.py  (Samples.compute_weights)
# --------------------------------------------------------------------------
def compute_weights(log_l, log_p, log_q) -> dict:
    """src/aspire/samples.py:457-475 (weights are the UNSHIFTED exp(log_w))."""
    log_w = log_l + log_p - log_q
    n = len(log_w)
    log_evidence = logsumexp(log_w) - math.log(n)
    weights = np.exp(log_w)
    evidence = np.exp(log_evidence)
    evidence_error = np.sqrt(np.sum((weights - evidence) ** 2) / (n * (n - 1)))
    log_evidence_error = np.abs(evidence_error / evidence)
    lw = log_w - np.max(log_w)
    ess = np.exp(np.asarray(logsumexp(lw) * 2 - logsumexp(lw * 2)))
    return dict(
        log_w=log_w,
        weights=weights,
        log_evidence=float(log_evidence),
        evidence=float(evidence),
        evidence_error=float(evidence_error),
        log_evidence_error=float(log_evidence_error),
        effective_sample_size=float(ess),
    )


# --------------------------------------------------------------------------
# src/aspire/samplers/smc/base.py
# --------------------------------------------------------------------------
def current_target_efficiency(target_efficiency, rate, beta):
    """src/aspire/samplers/smc/base.py:114-121"""
    if isinstance(target_efficiency, tuple):
        lo, hi = target_efficiency
        return lo + (hi - lo) * (beta**rate)
    return target_efficiency


def determine_beta(
    log_l,
    log_p,
    log_q,
    beta,
    *,
    adaptive=True,
    beta_step=float("nan"),
    min_beta_step=0.0,
    max_beta_step=1.0,
    beta_tolerance=1e-8,
    target_efficiency=0.5,
    target_efficiency_rate=1.0,
    adaptive_min_beta_step=False,
    trace: list | None = None,
):
    """src/aspire/samplers/smc/base.py:123-213.  ``trace`` (optional) collects
    every ``(beta_try, eff)`` the bisection evaluated."""
    n = len(log_l)
    if not adaptive:
        beta += beta_step
        if beta >= 1.0:
            beta = 1.0
        return beta, min_beta_step
    beta_prev = beta
    beta_min = beta_prev
    beta_max = 1.0
    eff_beta_max = effective_sample_size(log_weights(log_l, log_p, log_q, beta_prev, beta_max)) / n
    if trace is not None:
        trace.append((beta_max, float(eff_beta_max)))
    current_eff = current_target_efficiency(target_efficiency, target_efficiency_rate, beta_prev)
    if eff_beta_max >= current_eff:
        beta_min = 1.0
    target_eff = current_eff
    while beta_max - beta_min > beta_tolerance:
        beta_try = 0.5 * (beta_max + beta_min)
        eff = effective_sample_size(log_weights(log_l, log_p, log_q, beta_prev, beta_try)) / n
        if trace is not None:
            trace.append((beta_try, float(eff)))
        if eff >= target_eff:
            beta_min = beta_try
        else:
            beta_max = beta_try
    beta_star = beta_min
    if adaptive_min_beta_step:
        min_beta_step = min_beta_step * (1 - beta_prev) / (1 - beta_star)
    beta = max(beta_star, beta_prev + min_beta_step)
    beta = min(beta, beta_prev + max_beta_step, 1.0)
    if beta == beta_prev:
        raise BetaScheduleError(f"Beta did not increase from previous value {beta:.6g}.")
    return beta, min_beta_step


@dataclass
class StepRecord:
    """What one trip of the tempering loop appends to ``SMCHistory``
    (src/aspire/samplers/smc/base.py:412-440, src/aspire/history.py:71-82)."""

    beta: float
    eff_target: float
    ess: float
    ess_target: float
    log_norm_ratio: float
    log_norm_ratio_var: float
    idx: np.ndarray | None = field(default=None, repr=False)


def smc_step_reweight(
    log_l, log_p, log_q, beta0, **beta_kwargs
) -> tuple[StepRecord, float]:
    """One trip of the hot loop up to (not including) ``resample``:
    src/aspire/samplers/smc/base.py:404-440."""
    beta, min_step = determine_beta(log_l, log_p, log_q, beta0, **beta_kwargs)
    te = beta_kwargs.get("target_efficiency", 0.5)
    rate = beta_kwargs.get("target_efficiency_rate", 1.0)
    ess = effective_sample_size(log_weights(log_l, log_p, log_q, beta0, beta))
    ess_target = effective_sample_size(log_weights(log_l, log_p, log_q, beta0, 1.0))
    rec = StepRecord(
        beta=float(beta),
        eff_target=float(current_target_efficiency(te, rate, beta)),
        ess=float(ess),
        ess_target=float(ess_target),
        log_norm_ratio=float(log_evidence_ratio(log_l, log_p, log_q, beta0, beta)),
        log_norm_ratio_var=float(log_evidence_ratio_variance(log_l, log_p, log_q, beta0, beta)),
    )
    return rec, min_step


# --------------------------------------------------------------------------
# src/aspire/samplers/mcmc.py
# --------------------------------------------------------------------------
def filter_initial(log_q, log_p, log_l):
    """src/aspire/samplers/mcmc.py:68-91 - non-finite log_q is an error,
    particles with non-finite prior or likelihood are dropped."""
    if not np.isfinite(log_q).all():
        raise ValueError("Proposal returned non-finite log probabilities.")
    return np.isfinite(log_p) & np.isfinite(log_l)
