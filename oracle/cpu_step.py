"""One full SMC tempering step on the CPU - the timed baseline of bench.py.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  This is what the
reference executes per trip of its hot loop (``src/aspire/samplers/smc/base.py:401-454``)
restated in numpy: determine_beta (1 + ~20 ESS evaluations), ESS at beta and at 1,
log-evidence ratio and variance, ``resample`` (numpy ``Generator.choice`` + four
fancy-index gathers) and the mutation.  The mutation arithmetic is the restated pCN of
``oracle/pcn.py`` ("restated, not reference code": ``minipcn`` is absent), everything
else follows the reference line by line (``oracle/smc_numpy.py``).
"""

from __future__ import annotations

import numpy as np

from . import pcn
from . import smc_numpy as sn


def smc_step(x, log_l, log_p, log_q, problem, beta0=0.0, n_moves=10, rho=0.5, target_efficiency=0.5,
             beta_tolerance=1e-6, seed=0):
    rng = np.random.default_rng(seed)
    rec, _ = sn.smc_step_reweight(log_l, log_p, log_q, beta0, target_efficiency=target_efficiency,
                                  beta_tolerance=beta_tolerance)
    _, xr, llr, lpr, lqr = sn.resample(x, log_l, log_p, log_q, beta0, rec.beta, rng)
    mu, sigma = xr.mean(axis=0), xr.std(axis=0)
    n, d = xr.shape
    xi = rng.standard_normal((n_moves, n, d))
    log_u = np.log(1.0 - rng.random((n_moves, n)))
    out = pcn.sweep(xr, llr, lpr, lqr, like=problem["like"], prior=problem["prior"], q=problem["q"], beta=rec.beta,
                    mu=mu, sigma=sigma, rho=rho, n_moves=n_moves, xi=xi, log_u=log_u)
    return rec, out
