"""One SMC tempering step with the REFERENCE's own classes on the CPU (bench.py's ``cpu_baseline`` / ``--impl
reference`` arm when ``/root/reference`` is importable through ``oracle/ref_shim.py``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``SMCSampler.determine_beta`` (src/aspire/samplers/smc/base.py:
123-213), the ESS / evidence bookkeeping of the loop body (:419-440) and ``SMCSamples.resample`` (src/aspire/samples.py:
1251-1287) are the reference's unmodified code; the mutation is the restated pCN of ``oracle/pcn.py`` (``minipcn`` is
absent from the reference tree and from this image).  On the GPU box ``/root/reference`` does not exist and the numpy /
torch-CPU restatement (``oracle/cpu_step.py``, ``oracle/cpu_step_torch.py``) is timed instead (kind "port").
"""

from __future__ import annotations

import numpy as np

from . import densities as dn
from . import pcn, ref_shim


def available() -> bool:
    return ref_shim.reference_available()


class _Flow:
    def __init__(self, q, rng):
        self.q, self.rng = q, rng

    def sample_and_log_prob(self, n):
        x = self.q["mean"] + self.q["sigma"] * self.rng.standard_normal((n, len(self.q["mean"])))
        return x, dn.evaluate(self.q, x)

    def log_prob(self, x):
        return dn.evaluate(self.q, np.asarray(x))


def smc_step(x, log_l, log_p, log_q, problem, beta0=0.0, n_moves=10, rho=0.5, target_efficiency=0.5,
             beta_tolerance=1e-6, seed=0):
    ref_shim.import_reference()
    import array_api_compat.numpy as xp
    from aspire.samplers.smc.base import SMCSampler
    from aspire.samples import SMCSamples
    from aspire.utils import effective_sample_size

    rng = np.random.default_rng(seed)

    class _S(SMCSampler):
        def mutate(self, particles, beta, n_steps=None):  # pragma: no cover
            return particles

    s = _S(lambda s_: dn.evaluate(problem["like"], s_.x), lambda s_: dn.evaluate(problem["prior"], s_.x),
           problem["dims"], _Flow(problem["q"], rng), xp, dtype="float64", rng=rng)
    s.adaptive, s.adaptive_min_beta_step = True, False
    s.target_efficiency, s.target_efficiency_rate = target_efficiency, 1.0
    smc = SMCSamples(x=x, log_likelihood=log_l, log_prior=log_p, log_q=log_q, beta=beta0, xp=xp, dtype="float64")
    beta, _ = s.determine_beta(smc, beta0, np.nan, 0.0, 1.0, beta_tolerance)
    ess = effective_sample_size(smc.log_weights(beta))            # smc/base.py:419
    ess_target = effective_sample_size(smc.log_weights(1.0))      # :429-431
    ratio = smc.log_evidence_ratio(beta)                          # :433-440
    var = smc.log_evidence_ratio_variance(beta)
    new = smc.resample(beta, rng=rng)
    xr = np.asarray(new.x)
    mu, sigma = xr.mean(axis=0), xr.std(axis=0)
    n, d = xr.shape
    xi = rng.standard_normal((n_moves, n, d))
    log_u = np.log(1.0 - rng.random((n_moves, n)))
    out = pcn.sweep(xr, np.asarray(new.log_likelihood), np.asarray(new.log_prior), np.asarray(new.log_q),
                    like=problem["like"], prior=problem["prior"], q=problem["q"], beta=float(beta), mu=mu, sigma=sigma,
                    rho=rho, n_moves=n_moves, xi=xi, log_u=log_u)
    return dict(beta=float(beta), ess=float(ess), ess_target=float(ess_target), log_norm_ratio=float(ratio),
                log_norm_ratio_var=float(var)), out
