"""The CPU tempering step of ``oracle/cpu_step.py`` on torch CPU tensors with all host threads.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference runs on any array-API backend
(``xp``); with ``xp = torch`` on the CPU its elementwise passes, reductions, ``searchsorted`` and gathers
use torch's intra-op thread pool, which is how a user of the reference gets more than one core.  This is
the same loop body (``src/aspire/samplers/smc/base.py:401-454``: determine_beta, ESS at beta and at 1,
evidence ratio / variance, resample, mutate) written with torch CPU ops so that bench.py can report a
multi-threaded host baseline next to the single-threaded numpy one.  It is checked against the numpy
restatement in ``tests/test_oracle_cpu_torch.py``; built-in targets: Gaussian and Rosenbrock (config C3).
"""

from __future__ import annotations

import math

import numpy as np
import torch

from . import densities as dn


def _lse(x):
    c = torch.max(x)
    return float(c + torch.log(torch.sum(torch.exp(x - c))))


def _ess(lw):
    return math.exp(2.0 * _lse(lw) - _lse(2.0 * lw))


def _uw(ll, lp, lq, beta0, beta):
    return (beta0 - beta) * lq + (beta - beta0) * (ll + lp)


def _log_weights(ll, lp, lq, beta0, beta):
    uw = _uw(ll, lp, lq, beta0, beta)
    if bool(torch.isnan(uw).any()):
        raise ValueError(f"Log weights contain NaN values for beta={beta}")
    return uw + (_lse(uw) - math.log(len(uw)))


def determine_beta(ll, lp, lq, beta0, target_efficiency=0.5, beta_tolerance=1e-6):
    n = len(ll)
    beta_min, beta_max = beta0, 1.0
    if _ess(_log_weights(ll, lp, lq, beta0, 1.0)) / n >= target_efficiency:
        beta_min = 1.0
    while beta_max - beta_min > beta_tolerance:
        beta_try = 0.5 * (beta_max + beta_min)
        if _ess(_log_weights(ll, lp, lq, beta0, beta_try)) / n >= target_efficiency:
            beta_min = beta_try
        else:
            beta_max = beta_try
    return min(beta_min, 1.0)


def _density(desc, x):
    if desc["kind"] == dn.KIND_GAUSSIAN:
        z = (x - desc["_mean"]) * desc["_inv_sigma"]
        return desc["const"] - 0.5 * torch.sum(z * z, dim=-1)
    if desc["kind"] == dn.KIND_ROSENBROCK:
        x0, x1 = x[..., :-1], x[..., 1:]
        t, s = x1 - x0 * x0, desc["b"] - x0
        return -torch.sum(desc["a"] * t * t + s * s, dim=-1) / desc["scale"]
    raise NotImplementedError("cpu_step_torch: Gaussian and Rosenbrock targets only")


def _prep(desc):
    d = dict(desc)
    if d["kind"] == dn.KIND_GAUSSIAN:
        d["_mean"] = torch.as_tensor(np.asarray(d["mean"], dtype=np.float64))
        d["_inv_sigma"] = torch.as_tensor(np.asarray(d["inv_sigma"], dtype=np.float64))
    return d


def smc_step(x, log_l, log_p, log_q, problem, beta0=0.0, n_moves=10, rho=0.5, target_efficiency=0.5,
             beta_tolerance=1e-6, seed=0, threads=None):
    """Same signature and result layout as ``cpu_step.smc_step``; inputs are numpy arrays."""
    if threads:
        torch.set_num_threads(int(threads))
    rng = np.random.default_rng(seed)
    x, ll, lp, lq = (torch.as_tensor(np.ascontiguousarray(a)) for a in (x, log_l, log_p, log_q))
    like, prior, q = _prep(problem["like"]), _prep(problem["prior"]), _prep(problem["q"])
    n, d = x.shape
    beta = determine_beta(ll, lp, lq, beta0, target_efficiency, beta_tolerance)
    lw = _log_weights(ll, lp, lq, beta0, beta)
    ess, ess_target = _ess(lw), _ess(_log_weights(ll, lp, lq, beta0, 1.0))
    uw = _uw(ll, lp, lq, beta0, beta)
    ler = _lse(uw) - math.log(n)
    u_ = torch.exp(uw - torch.max(uw))
    ler_var = float(torch.var(u_, unbiased=False) / (n * torch.mean(u_) ** 2))
    # resample: numpy Generator.choice semantics (sequential fp64 cumsum, searchsorted 'right')
    p = torch.exp(lw - _lse(lw))
    cdf = torch.cumsum(p, dim=0)
    cdf = cdf / cdf[-1]
    idx = torch.searchsorted(cdf, torch.as_tensor(rng.random(n)), right=True).clamp_(max=n - 1)
    x, ll, lp, lq = x[idx], ll[idx], lp[idx], lq[idx]
    mu, sigma = x.mean(dim=0), x.std(dim=0, unbiased=False)
    inv_sigma = 1.0 / sigma
    a = math.sqrt(1.0 - rho * rho)
    accepts = []
    for _ in range(n_moves):
        xi = torch.as_tensor(rng.standard_normal((n, d)))
        log_u = torch.as_tensor(np.log(1.0 - rng.random(n)))
        xp = mu + (a * (x - mu) + (rho * sigma) * xi)
        lq2, lp2, ll2 = _density(q, xp), _density(prior, xp), _density(like, xp)
        cur = (1.0 - beta) * lq + beta * (ll + lp)
        new = (1.0 - beta) * lq2 + beta * (ll2 + lp2)
        cur = torch.where(torch.isnan(cur), torch.full_like(cur, -math.inf), cur)
        new = torch.where(torch.isnan(new), torch.full_like(new, -math.inf), new)
        z0, z1 = (x - mu) * inv_sigma, (xp - mu) * inv_sigma
        log_a = (new - cur) + (0.5 * torch.sum(z1 * z1, dim=-1) - 0.5 * torch.sum(z0 * z0, dim=-1))
        acc = log_u < log_a
        accepts.append(int(acc.sum()))
        x = torch.where(acc[:, None], xp, x)
        ll, lp, lq = torch.where(acc, ll2, ll), torch.where(acc, lp2, lp), torch.where(acc, lq2, lq)
    rec = dict(beta=beta, ess=ess, ess_target=ess_target, log_norm_ratio=ler, log_norm_ratio_var=ler_var)
    return rec, (x.numpy(), ll.numpy(), lp.numpy(), lq.numpy(), np.asarray(accepts))
