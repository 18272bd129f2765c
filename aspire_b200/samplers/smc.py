"""``sampler="smc"`` - the tempering loop on the B200 kernels.

Mirrors, name for name, ``SMCSampler`` (``src/aspire/samplers/smc/base.py:30-488``)
and ``MiniPCNSMC`` (``src/aspire/samplers/smc/minipcn.py:14-135``) of the
reference: same constructor, same ``sample`` keywords and defaults, same
exceptions, same ``SMCHistory`` fields, same order of operations per trip of
the loop (determine_beta -> history -> resample -> mutate).  The arithmetic is
the C-ABI's: ``asp_reweight`` + device-driven bisection (``asp_beta_search_*``)
+ ``asp_smc_weights`` + the exact CDF / ancestor search + ``asp_pcn_sweep``.

Two mutation paths:
* fused   - likelihood / prior are built-in ``aspire_b200.targets`` densities and
            the proposal flow exposes ``.density``: the resampling gather and
            all pCN moves run in one kernel (particles read once, written once);
* unfused - arbitrary Python callables: ``asp_pcn_propose`` / callables /
            ``asp_pcn_accept`` per move.
"""

from __future__ import annotations

import logging
import math
from typing import Any, Callable

import numpy as np
import torch

from .. import ops
from ..history import SMCHistory
from ..rng import as_uniform_source
from ..samples import Samples, SMCSamples
from ..targets import Density
from .base import Sampler, track_calls

logger = logging.getLogger(__name__)

DEFAULT_BETA_TOLERANCE = 1e-8  # smc/base.py:23
SAMPLE_HISTORY_MAX_ELEMENTS = 20_000_000  # above this the per-step host copies are refused (SURVEY hard part 9)


class BetaScheduleError(RuntimeError):
    """smc/base.py:26"""


# Bounds of the between-sweep step-size adaptation (oracle/pcn.py::adapt_rho).  The upper bound stops just short of
# the independence proposal rho = 1: with a high acceptance the adaptation saturates after a step or two, and at
# rho < 1 every mutation keeps its fastest kernel (the particle-per-thread sweep carries sqrt(1 - rho^2) (x - mu) as its
# state and cannot represent rho = 1; an explicit rho = 1 is still honoured, through the general kernels).  The step
# x' = mu + 1.4e-3 (x - mu) + 0.999999 sigma xi is a valid pCN step like any other.
RHO_MIN, RHO_MAX = 1e-3, 1.0 - 1e-6


class SMCSampler(Sampler):
    default_beta_tolerance = DEFAULT_BETA_TOLERANCE
    default_step_fn = "pcn"  # the base class runs plain pCN (the north-star kernel); MiniPCNSMC: the reference's "tpcn"

    def __init__(
        self,
        log_likelihood: Callable,
        log_prior: Callable,
        dims: int,
        prior_flow,
        xp: Callable = None,
        dtype: Any | str | None = None,
        parameters: list[str] | None = None,
        rng=None,
        preconditioning_transform: Callable | None = None,
    ):
        super().__init__(log_likelihood, log_prior, dims, prior_flow, xp, dtype, parameters, preconditioning_transform)
        self.rng = as_uniform_source(rng)
        self._adaptive_target_efficiency = False
        self.sampler_kwargs: dict | None = None
        self._weights_cache = None

    # ---- target efficiency (smc/base.py:80-121) ---------------------------
    @property
    def target_efficiency(self):
        return self._target_efficiency

    @target_efficiency.setter
    def target_efficiency(self, value):
        if isinstance(value, float):
            if not (0 < value < 1):
                raise ValueError("target_efficiency must be in (0, 1)")
            self._target_efficiency = value
            self._adaptive_target_efficiency = False
        elif len(value) != 2:
            raise ValueError("target_efficiency must be a float or tuple of two floats")
        else:
            value = tuple(map(float, value))
            if not (0 < value[0] < value[1] < 1):
                raise ValueError("target_efficiency tuple must be in (0, 1) and increasing")
            self._target_efficiency = value
            self._adaptive_target_efficiency = True

    def current_target_efficiency(self, beta: float) -> float:
        if self._adaptive_target_efficiency:
            lo, hi = self._target_efficiency
            return lo + (hi - lo) * (beta**self.target_efficiency_rate)
        return self._target_efficiency

    # ---- fused-path detection ---------------------------------------------
    @property
    def fused(self) -> bool:
        return (isinstance(self._log_likelihood, Density) and isinstance(self._log_prior, Density)
                and isinstance(getattr(self.prior_flow, "density", None), Density) and self.dims <= 128)

    # ---- initial particles (samplers/mcmc.py:49-110) ------------------------
    def draw_initial_samples(self, n_samples: int) -> Samples:
        n_drawn, samples = 0, None
        while n_drawn < n_samples:
            x, log_q = self.prior_flow.sample_and_log_prob(n_samples)
            log_q = torch.as_tensor(log_q, dtype=torch.float64, device=self.device)
            finite = bool(torch.isfinite(log_q).all())
            self._check_flow()  # a kernel time-out is reported as such, not as a bad proposal
            if not finite:
                raise ValueError(
                    "Proposal returned non-finite log probabilities. aspire assumes the proposal is a valid, "
                    "normalized probability distribution and should therefore only return samples with finite "
                    "log probabilities.")
            new = Samples(x, log_q=log_q, parameters=self.parameters, device=self.device)
            new.log_prior = self.log_prior(new)
            new.log_likelihood = self.log_likelihood(new)
            valid = torch.isfinite(new.log_prior) & torch.isfinite(new.log_likelihood)
            n_valid = int(valid.sum())
            if n_valid == len(new):
                piece = new
            elif n_valid > 0:
                piece = new[valid]
            else:
                piece = None
            if piece is not None:
                samples = piece if samples is None else Samples.concatenate([samples, piece])
                n_drawn += n_valid
        if n_drawn > n_samples:
            samples = samples[:n_samples]
        return samples

    # ---- beta (smc/base.py:123-213) ------------------------------------------
    def determine_beta(self, samples: SMCSamples, beta: float, beta_step: float, min_beta_step: float,
                       max_beta_step: float = 1.0, beta_tolerance: float = DEFAULT_BETA_TOLERANCE):
        d, stats, st = samples._slope()
        n = len(samples)
        if st[2] > 0 or st[3] > 0:
            raise ValueError(f"Log weights contain NaN values for beta={beta}")
        beta_prev = beta
        sums = None
        ess_target = None
        n_evals = 0
        if not self.adaptive:
            beta += beta_step
            if beta >= 1.0:
                beta = 1.0
        else:
            target_eff = self.current_target_efficiency(beta_prev)
            state = ops.beta_search(d, stats, beta_prev, target_eff, beta_tolerance, n_total=n)
            n_evals = state.n_evals
            beta_star = state.beta_min
            ess_target = state.eff_beta_max * n
            if beta_star <= beta_prev + beta_tolerance and beta_prev < 1.0:
                logger.warning(
                    "Adaptive beta search could not find a beta above %.6g that satisfies the target efficiency "
                    "%.3f within tolerance %.1e; beta may remain unchanged. Consider decreasing beta_tolerance or "
                    "target_efficiency.", beta_prev, target_eff, beta_tolerance)
            if self.adaptive_min_beta_step and beta_star < 1.0:
                # smc/base.py:198-201; the reference divides by zero here when the search lands on beta = 1
                min_beta_step = min_beta_step * (1 - beta_prev) / (1 - beta_star)
            beta = max(beta_star, beta_prev + min_beta_step)
            beta = min(beta, beta_prev + max_beta_step, 1.0)
            if beta == beta_prev:
                raise BetaScheduleError(
                    f"Beta did not increase from previous value {beta:.6g}. Adaptive beta search may have failed to "
                    f"find a suitable beta. Consider adjusting beta_tolerance ({beta_tolerance}), min_beta_step "
                    f"({min_beta_step}) or target_efficiency ({target_eff}) (values may be adaptive).")
            if state.min_valid and beta == beta_star:
                sums = (state.s1_min, state.s2_min)
        dbeta = beta - beta_prev
        if sums is None or ess_target is None:
            s = ops.ess_sums(d, [dbeta, 1.0 - beta_prev], stats).cpu().numpy()
            sums = (float(s[0, 0]), float(s[0, 1])) if sums is None else sums
            ess_target = float(s[1, 0] ** 2 / s[1, 1]) if ess_target is None else ess_target
        c = dbeta * (st[1] if dbeta >= 0 else st[0])
        self._weights_cache = dict(beta_prev=beta_prev, beta=beta, d=d, dbeta=dbeta, c=c, s1=sums[0], s2=sums[1],
                                   ess_target=ess_target, n_evals=n_evals)
        return beta, min_beta_step

    # ---- the loop (smc/base.py:215-488, smc/minipcn.py:22-67) ---------------
    @track_calls
    def sample(
        self,
        n_samples: int,
        n_steps: int = None,
        min_beta_step: float | None = None,
        max_beta_step: float | None = None,
        max_n_steps: int | None = None,
        adaptive: bool = True,
        target_efficiency: float = 0.5,
        target_efficiency_rate: float = 1.0,
        n_final_samples: int | None = None,
        sampler_kwargs: dict | None = None,
        rng=None,
        checkpoint_callback=None,
        checkpoint_every: int | None = None,
        checkpoint_file_path: str | None = None,
        resume_from=None,
        beta_tolerance: float | None = None,
        store_sample_history: bool = True,
        resampling: str = "multinomial",
    ):
        if beta_tolerance is None:
            beta_tolerance = self.default_beta_tolerance
        self._configure_kernel(sampler_kwargs)
        if rng is not None:
            self.rng = as_uniform_source(rng)
        self.resampling = resampling

        resumed = resume_from is not None
        if resumed:  # smc/base.py:301-314
            samples, beta, iterations = self.restore_from_checkpoint(resume_from)
            logger.info(f"Resumed SMC sampling at iteration {iterations} with beta={beta:.4f}")
        else:
            samples = self.draw_initial_samples(n_samples)
            samples = SMCSamples.from_samples(samples, beta=0.0)
            beta, iterations = 0.0, 0
            self.history = SMCHistory()
        if store_sample_history and samples.x.numel() > SAMPLE_HISTORY_MAX_ELEMENTS:
            logger.warning("store_sample_history=True ignored: %d x %d particles per step exceed the %d-element limit of "
                           "the per-step host copies (history.sample_history stays empty)", samples.x.shape[0],
                           samples.x.shape[1], SAMPLE_HISTORY_MAX_ELEMENTS)
            store_sample_history = False
        if store_sample_history:
            self.history.sample_history.append(samples.to_numpy())

        if bool(torch.isnan(samples.log_q).any()):
            raise ValueError("Log proposal contains NaN values")
        if bool(torch.isnan(samples.log_prior).any()):
            raise ValueError("Log prior contains NaN values")
        if bool(torch.isnan(samples.log_likelihood).any()):
            raise ValueError("Log likelihood contains NaN values")

        n_final_steps = self.sampler_kwargs.pop("n_final_steps", None)
        self.target_efficiency = target_efficiency
        self.target_efficiency_rate = target_efficiency_rate
        if n_steps is not None:
            beta_step = 1 / n_steps
        elif not adaptive:
            raise ValueError("Either n_steps or adaptive=True must be set")
        else:
            beta_step = math.nan
        self.adaptive = adaptive
        if min_beta_step is None:
            if max_n_steps is None:
                min_beta_step, self.adaptive_min_beta_step = 0.0, False
            else:
                min_beta_step, self.adaptive_min_beta_step = 1 / max_n_steps, True
        else:
            self.adaptive_min_beta_step = False
        if max_beta_step is not None:
            if max_beta_step <= 0 or max_beta_step >= 1:
                raise ValueError("max_beta_step must be in (0, 1)")
            self.max_beta_step = max_beta_step
        else:
            self.max_beta_step = 1.0

        iterations = iterations or 0
        if checkpoint_callback is None and checkpoint_every is not None:  # smc/base.py:366-371
            checkpoint_callback = self.default_file_checkpoint_callback(checkpoint_file_path)
        if checkpoint_callback is not None and checkpoint_every is None:
            checkpoint_every = 1
        run_smc_loop = True
        if resumed:
            last_beta = self.history.beta[-1] if self.history.beta else beta
            if last_beta >= 1.0:
                run_smc_loop = False
                logger.info(f"Checkpoint beta {last_beta:.4f} indicates the SMC loop already completed")

        while run_smc_loop:
            iterations += 1
            samples, beta, min_beta_step = self.tempering_step(samples, beta, beta_step, min_beta_step,
                                                               beta_tolerance, iterations)
            if store_sample_history:
                self.history.sample_history.append(samples.to_numpy())
            if checkpoint_callback is not None and checkpoint_every > 0 and iterations % checkpoint_every == 0:
                checkpoint_callback(self.build_checkpoint_state(samples, iterations, beta))
            if beta == 1.0 or (max_n_steps is not None and iterations >= max_n_steps):
                break

        if n_final_samples is not None and len(samples.x) != n_final_samples:
            logger.info(f"Generating {n_final_samples} final samples")
            idx = samples.resample_indices(1.0, n_samples=n_final_samples, rng=self.rng, method=self.resampling)
            samples = self.mutate(samples, 1.0, n_steps=n_final_steps, idx=idx)

        samples.log_evidence = float(np.sum(np.asarray(self.history.log_norm_ratio)))
        samples.log_evidence_error = float(np.sqrt(np.sum(np.asarray(self.history.log_norm_ratio_var))))
        if checkpoint_callback is not None:  # maybe_checkpoint(force=True), smc/base.py:476: the final state is
            checkpoint_callback(self.build_checkpoint_state(samples, iterations, beta))  # written whatever checkpoint_every
        final = samples.to_standard_samples()
        logger.info(f"Log evidence: {final.log_evidence:.2f} +/- {final.log_evidence_error:.2f}")
        return final

    def tempering_step(self, samples: SMCSamples, beta: float, beta_step: float, min_beta_step: float,
                       beta_tolerance: float, iterations: int = 0, out=None, mark=None):
        """One trip of the hot loop (smc/base.py:401-454): determine_beta, history,
        evidence bookkeeping, resample, mutate.  ``out`` optionally supplies the
        output buffers of the sweep, ``mark(name)`` is called between phases
        (bench.py records CUDA events there).  -> (samples, beta, min_beta_step)"""
        mark = mark or (lambda name: None)
        n = len(samples)
        beta, min_beta_step = self.determine_beta(samples, beta, beta_step, min_beta_step,
                                                  max_beta_step=self.max_beta_step, beta_tolerance=beta_tolerance)
        mark("beta_search")
        w = self._weights_cache
        self.history.eff_target.append(float(self.current_target_efficiency(beta)))
        logger.info(f"it {iterations} - beta: {beta}")
        self.history.beta.append(float(beta))
        ess = w["s1"] ** 2 / w["s2"]
        eff = ess / n
        if eff < 0.1:
            logger.warning(f"it {iterations} - Low sample efficiency: {eff:.2f}")
        self.history.ess.append(float(ess))
        self.history.ess_target.append(float(w["ess_target"]))
        self.history.n_beta_evals.append(int(w["n_evals"]))
        # evidence bookkeeping + resampling weights in one pass over d (samples.py:1226-1242, 1277)
        p, result, tiles = ops.smc_weights(w["d"], w["dbeta"], w["c"], w["s1"])
        log_ratio = w["c"] + math.log(w["s1"]) - math.log(n)
        mean_u = w["s1"] / n
        self.history.log_norm_ratio.append(float(log_ratio))
        # resample: exact CDF + ancestor search (samples.py:1251-1287 / numpy Generator.choice)
        cdf = ops.cdf_exact(p, tile_stats=tiles)
        if self.resampling == "systematic":
            idx = ops.search_systematic(cdf, self.rng.scalar_uniform(), n)
        else:
            idx = self.rng.multinomial_ancestors(cdf, n)
        del p, cdf
        mark("resample")
        samples = self.mutate(samples, beta, idx=idx, out=out, mark=mark)
        var_sum = float(result[0])
        self.history.log_norm_ratio_var.append(var_sum / n / (n * mean_u * mean_u) if mean_u != 0 else math.nan)
        logger.info(f"it {iterations} - ESS: {ess:.1f} ({eff:.2f} efficiency), log evidence ratio {log_ratio:.2f}")
        return samples, beta, min_beta_step

    # ---- checkpoint / resume (smc/base.py:521-562) ----------------------------------
    def build_checkpoint_state(self, samples, iteration: int, beta: float) -> dict:
        return super().build_checkpoint_state(samples, iteration, meta={"beta": beta})

    def _rng_state(self):
        gen = getattr(self.rng, "rng", None)
        if gen is not None and hasattr(gen, "bit_generator"):
            return gen.bit_generator.state
        if hasattr(self.rng, "seed"):
            return {"bit_generator": "Philox4x32-10 (aspire_b200)", "seed": self.rng.seed, "stream": self.rng._stream}
        return None

    def _checkpoint_extra_state(self) -> dict:
        import copy

        return {"history": copy.deepcopy(self.history), "rng_state": self._rng_state(),
                "sampler_kwargs": dict(getattr(self, "sampler_kwargs", None) or {}), "rho": getattr(self, "rho", None)}

    def restore_from_checkpoint(self, source):
        """-> (SMCSamples on the device, beta, iteration); restores history, rng state and the pCN step."""
        saved, state = super().restore_from_checkpoint(source)
        meta = state.get("meta", {}) if isinstance(state, dict) else {}
        beta = meta.get("beta") if isinstance(meta, dict) else None
        if beta is None:
            beta = state.get("beta", 0.0)
        self.history = state.get("history") or SMCHistory()
        rs = state.get("rng_state")
        if rs is not None:
            gen = getattr(self.rng, "rng", None)
            if gen is not None and hasattr(gen, "bit_generator") and "state" in rs:
                gen.bit_generator.state = rs
            elif hasattr(self.rng, "seed") and "seed" in rs:
                self.rng.seed, self.rng._stream = int(rs["seed"]), int(rs["stream"])
        if state.get("rho") is not None:
            self.rho = float(state["rho"])
        get = (lambda k: saved.get(k)) if isinstance(saved, dict) else (lambda k: getattr(saved, k, None))
        samples = SMCSamples(x=get("x"), log_likelihood=get("log_likelihood"), log_prior=get("log_prior"),
                             log_q=get("log_q"), beta=float(beta), parameters=self.parameters, device=self.device)
        return samples, float(beta), int(state.get("iteration") or 0)

    # ---- kernel configuration (smc/minipcn.py:44-50) ---------------------------
    def _configure_kernel(self, sampler_kwargs):
        kw = dict(sampler_kwargs or {})
        kw.setdefault("n_steps", 5 * self.dims)
        kw.setdefault("target_acceptance_rate", 0.234)
        kw.setdefault("step_fn", self.default_step_fn)
        kw.setdefault("verbose", False)
        kw.setdefault("rho", 0.5)
        kw.setdefault("adapt_rho", True)
        kw.setdefault("moment_particles", 1 << 22)
        if kw["step_fn"] not in ("pcn", "tpcn"):
            raise ValueError(f"step_fn={kw['step_fn']!r}: expected 'pcn' or 'tpcn'")
        # both run the fused sweep with built-in densities; 'tpcn' (the reference's default, smc/minipcn.py:48)
        # first fits a Student-t reference measure by EM; other dims / callables: unfused moves
        self.step_fn = kw["step_fn"]
        self.sampler_kwargs = kw
        self.rho = float(kw["rho"])

    # ---- tempered target for the unfused path (smc/base.py:507-519) ----------
    def log_prob(self, x, beta=None):
        s = SMCSamples(x, beta=beta, device=self.device, parameters=self.parameters)
        s.log_q = torch.as_tensor(self.prior_flow.log_prob(s.x), dtype=torch.float64, device=self.device)
        s.log_prior = self.log_prior(s)
        s.log_likelihood = self.log_likelihood(s)
        lp = s.log_p_t(beta)
        return torch.where(torch.isnan(lp), torch.full_like(lp, -math.inf), lp)

    # ---- mutation (smc/minipcn.py:69-135) --------------------------------------
    def _pcn_params(self, beta, mu, sigma, n_moves, n, nu, tpcn):
        """Noise of one mutation: host draws in a fixed order (parity mode) or Philox streams."""
        dims = self.dims
        if self.rng.is_host:
            xi = torch.from_numpy(self.rng.rng.standard_normal((n_moves, n, dims))).to(self.device)
            log_u = torch.from_numpy(np.log(1.0 - self.rng.rng.random((n_moves, n)))).to(self.device)
            gam = (torch.from_numpy(self.rng.rng.gamma(0.5 * (dims + nu), 1.0, (n_moves, n))).to(self.device)
                   if tpcn else None)
            return ops.make_pcn_params(beta, self.rho, mu, sigma, n_moves, xi=xi, log_u=log_u, nu=nu, gamma=gam)
        return ops.make_pcn_params(beta, self.rho, mu, sigma, n_moves, seed=self.rng.seed,
                                   stream0=self.rng.next_stream(n_moves), index0=getattr(self, "index0", 0),
                                   total_accepts_only=False, nu=nu)

    def _flow_sweep_ok(self, particles, n_moves: int, tpcn: bool) -> bool:
        """Whether the mutation can run in ``asp_pcn_sweep_flow``: pCN, built-in Gaussian prior and Gaussian /
        mixture likelihood, an unpadded ``CouplingFlow`` (16 / 32 / 64 dimensions) as ``prior_flow``."""
        from ..flows import CouplingFlow

        f = self.prior_flow
        return (not tpcn and self.sampler_kwargs.get("flow_sweep", True) and type(f) is CouplingFlow and not f._pad
                and isinstance(self._log_likelihood, Density) and isinstance(self._log_prior, Density)
                and n_moves <= 256 and particles.x.data_ptr() % 32 == 0  # (rho = 1, the independence proposal, included)
                and ops.pcn_sweep_flow_supported(self.dims, self._log_likelihood, self._log_prior))

    def _mutate_preconditioned(self, particles: SMCSamples, beta: float, n_moves: int, idx, mark) -> SMCSamples:
        """Mutation in the space of the preconditioning transform (smc/minipcn.py:102-131): z = T.fit(x), pCN / tpCN
        moves on z against ``log_prob(z) = log_p_t(T^-1(z)) + log|det J|`` (smc/base.py:507-519), x = T^-1(z_last),
        the three logs recomputed at x.  The Jacobian is folded into the log prior and log proposal handed to the
        accept kernel: (1-b)(lq + J) + b(ll + lp + J) = (1-b) lq + b(ll + lp) + J."""
        T = self.preconditioning_transform
        n = len(particles) if idx is None else idx.numel()
        dims = self.dims
        if idx is not None:
            x, ll, lp, lq = ops.gather_rows(particles.x, idx, particles.log_likelihood, particles.log_prior,
                                            particles.log_q)
        else:
            x, ll, lp, lq = (t.clone() for t in (particles.x, particles.log_likelihood, particles.log_prior,
                                                 particles.log_q))
        z = T.fit(x).contiguous()
        _, lj = T.inverse(z)
        lpj, lqj = (lp + lj).contiguous(), (lq + lj).contiguous()
        m_use = min(n, int(self.sampler_kwargs.get("moment_particles", 1 << 22)))
        mu, sigma = ops.moments_to_mu_sigma(ops.weighted_moments(z[:m_use]), dims)
        sigma = torch.clamp_min(sigma, 1e-150)
        tpcn = getattr(self, "step_fn", "pcn") == "tpcn"
        nu = 0.0
        if tpcn:
            mu, sigma, nu = fit_student_t_diag(z[:min(m_use, 1 << 20)].contiguous(), mu, sigma)
            self.history.__dict__.setdefault("tpcn_nu", []).append(nu)
        if mark:
            mark("moments")
        prm = self._pcn_params(beta, mu, sigma, n_moves, n, nu, tpcn)
        accepts = torch.zeros(max(n_moves, 1), dtype=torch.int64, device=self.device)
        for t in range(n_moves):
            if tpcn:
                ops.tpcn_scales(z, prm, t)
            zp = ops.pcn_propose(z, prm, t)
            xp, lj2 = T.inverse(zp)
            prop = SMCSamples(xp, beta=beta, device=self.device, parameters=self.parameters)
            lq2 = torch.as_tensor(self.prior_flow.log_prob(xp), dtype=torch.float64, device=self.device).contiguous()
            prop.log_q = lq2
            lp2 = self.log_prior(prop)
            prop.log_prior = lp2
            ll2 = self.log_likelihood(prop)
            ops.pcn_accept(z, ll, lpj, lqj, zp, ll2, (lp2 + lj2).contiguous(), (lq2 + lj2).contiguous(), prm, t, accepts)
        if mark:
            mark("pcn_sweep")
        acc = float(accepts[:n_moves].sum()) / max(n_moves * n, 1)
        self._check_flow()
        self.history.mcmc_acceptance.append(acc)
        self.history.rho.append(self.rho)
        if self.sampler_kwargs.get("adapt_rho", True):
            target = self.sampler_kwargs["target_acceptance_rate"]
            self.rho = float(min(max(self.rho * math.exp(acc - target), RHO_MIN), RHO_MAX))
        x = T.inverse(z)[0]
        out = SMCSamples(x, beta=beta, parameters=self.parameters, device=self.device)
        out.log_q = torch.as_tensor(self.prior_flow.log_prob(out.x), dtype=torch.float64, device=self.device).contiguous()
        out.log_prior = self.log_prior(out)
        out.log_likelihood = self.log_likelihood(out)
        if bool(torch.isnan(out.log_q).any()):
            raise ValueError("Log proposal contains NaN values")  # smc/minipcn.py:133-134
        return out

    def mutate(self, particles: SMCSamples, beta: float, n_steps: int | None = None, idx=None, out=None,
               mark=None) -> SMCSamples:
        n_moves = int(n_steps or self.sampler_kwargs["n_steps"])
        T = self.preconditioning_transform
        if T is not None and not getattr(T, "is_identity", False):
            return self._mutate_preconditioned(particles, beta, n_moves, idx, mark)
        n = len(particles) if idx is None else idx.numel()
        dims = self.dims
        # reference measure of the pCN proposal: moments of the (resampled) ensemble; beyond `moment_particles`
        # rows (default 2^22) every (n // moment_particles)-th offspring is used - the ancestors of the sorted
        # resampler are ascending, so a strided sample is spread over the whole genealogy (a prefix would not be)
        m_use = min(n, int(self.sampler_kwargs.get("moment_particles", 1 << 22)))
        stride = max(n // max(m_use, 1), 1)
        mom = ops.weighted_moments(particles.x, idx=idx, stride=stride)  # the strided rows are read in place
        mu, sigma = ops.moments_to_mu_sigma(mom, dims, floor=1e-150)
        tpcn = getattr(self, "step_fn", "pcn") == "tpcn"
        nu = 0.0
        if tpcn:  # Student-t reference measure t_nu(mu, diag sigma^2) fitted to (a sample of) the ensemble
            t_stride = max(n // min(m_use, 1 << 20), 1)
            mu, sigma, nu = fit_student_t_diag(particles.x, mu, sigma, idx=idx, stride=t_stride)
            self.history.__dict__.setdefault("tpcn_nu", []).append(nu)
        if mark:
            mark("moments")
        if self.rng.is_host:  # parity mode: noise drawn on the host in a fixed order, then uploaded
            xi = torch.from_numpy(self.rng.rng.standard_normal((n_moves, n, dims))).to(self.device)
            log_u = torch.from_numpy(np.log(1.0 - self.rng.rng.random((n_moves, n)))).to(self.device)
            gam = (torch.from_numpy(self.rng.rng.gamma(0.5 * (dims + nu), 1.0, (n_moves, n))).to(self.device)
                   if tpcn else None)
            prm = ops.make_pcn_params(beta, self.rho, mu, sigma, n_moves, xi=xi, log_u=log_u, nu=nu, gamma=gam)
        else:
            prm = ops.make_pcn_params(beta, self.rho, mu, sigma, n_moves, seed=self.rng.seed,
                                      stream0=self.rng.next_stream(n_moves), index0=getattr(self, "index0", 0),
                                      total_accepts_only=True, nu=nu)
        fused = self.fused and (not tpcn or (ops.pcn_sweep_tpcn_supported(dims) and self.rho < 1.0
                                             and particles.x.data_ptr() % 16 == 0))
        if not fused and not self.rng.is_host:
            prm.reserved = 0  # the unfused kernels count accepts per move
        fused_kernel = True  # a sweep kernel ran the whole mutation (and counted the NaN log q itself)
        if fused:
            x, ll, lp, lq, accepts = ops.pcn_sweep(particles.x, particles.log_likelihood, particles.log_prior,
                                                   particles.log_q, idx, self._log_likelihood, self._log_prior,
                                                   self.prior_flow.density, prm, out=out)
            if mark:
                mark("pcn_sweep")
            self.n_likelihood_evaluations += n * n_moves
        elif self._flow_sweep_ok(particles, n_moves, tpcn):
            # built-in prior / likelihood, coupling flow as the proposal density (C4): gather, all moves with
            # log q(x') on tcgen05 inside the move, and the final store in one kernel (csrc/flow_move.cu)
            prm.reserved = 1
            x, ll, lp, lq, accepts = ops.pcn_sweep_flow(particles.x, particles.log_likelihood, particles.log_prior,
                                                        particles.log_q, idx, self._log_likelihood, self._log_prior,
                                                        self.prior_flow.struct(), prm, out=out)
            if mark:
                mark("pcn_sweep")
            self.n_likelihood_evaluations += n * n_moves
        else:
            fused_kernel = False
            if idx is not None:
                x, ll, lp, lq = ops.gather_rows(particles.x, idx, particles.log_likelihood, particles.log_prior,
                                                particles.log_q)
            else:
                x, ll, lp, lq = (t.clone() for t in (particles.x, particles.log_likelihood, particles.log_prior,
                                                     particles.log_q))
            accepts = torch.zeros(max(n_moves, 1), dtype=torch.int64, device=self.device)
            # built-in prior / likelihood with a proposal that only offers log_prob (a flow): one pass proposes
            # and evaluates both densities, the flow adds log q(x'), a scalar kernel accepts (csrc/pcn_eval.cu)
            semi = (not tpcn and isinstance(self._log_likelihood, Density) and isinstance(self._log_prior, Density)
                    and self.rho < 1.0 and x.data_ptr() % 16 == 0
                    and ops.pcn_propose_eval_supported(dims, self._log_likelihood, self._log_prior))
            for t in range(n_moves if semi else 0):
                xp, ll2, lp2, dr = ops.pcn_propose_eval(x, self._log_likelihood, self._log_prior, prm, t)
                lq2 = torch.as_tensor(self.prior_flow.log_prob(xp), dtype=torch.float64, device=self.device).contiguous()
                ops.pcn_accept_scalar(x, ll, lp, lq, xp, ll2, lp2, lq2, dr, prm, t, accepts)
                self.n_likelihood_evaluations += n
            for t in range(0 if semi else n_moves):
                if tpcn:
                    ops.tpcn_scales(x, prm, t)  # sqrt(s_i) of this move, attached to prm
                xp = ops.pcn_propose(x, prm, t)
                prop = SMCSamples(xp, beta=beta, device=self.device, parameters=self.parameters)
                lq2 = torch.as_tensor(self.prior_flow.log_prob(xp), dtype=torch.float64, device=self.device).contiguous()
                prop.log_q = lq2
                lp2 = self.log_prior(prop)
                prop.log_prior = lp2
                ll2 = self.log_likelihood(prop)
                ops.pcn_accept(x, ll, lp, lq, xp, ll2, lp2, lq2, prm, t, accepts)
            accepts = accepts[:n_moves]
        # the sweeps count the final log q that are NaN themselves (asp_pcn_sweep: the entry behind the accept counts):
        # one device -> host read serves the acceptance and the reference's NaN check (smc/minipcn.py:133-134)
        nan_count = getattr(prm, "_nan", None) if fused_kernel else None
        if nan_count is not None:
            acc_sum, n_nan = torch.cat([accepts.sum().reshape(1), nan_count]).tolist()
        else:
            acc_sum, n_nan = float(accepts.sum()), None
        acc = acc_sum / max(n_moves * n, 1)
        self._check_flow()  # right after the mutation's own device -> host read
        self.history.mcmc_acceptance.append(acc)
        self.history.rho.append(self.rho)
        if self.sampler_kwargs.get("adapt_rho", True):
            target = self.sampler_kwargs["target_acceptance_rate"]
            self.rho = float(min(max(self.rho * math.exp(acc - target), RHO_MIN), RHO_MAX))
        out = SMCSamples(x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=beta, parameters=self.parameters,
                         device=self.device)
        if (n_nan > 0) if n_nan is not None else bool(torch.isnan(lq).any()):
            raise ValueError("Log proposal contains NaN values")  # smc/minipcn.py:133-134
        return out


def fit_student_t_diag(x: torch.Tensor, mu0: torch.Tensor, sigma0: torch.Tensor, idx=None, stride: int = 1,
                       n_iter: int = 30, nu0: float = 10.0, nu_min: float = 1.0, nu_max: float = 1e6):
    """EM fit of t_nu(mu, diag sigma^2) to the rows ``x[idx[::stride]]`` (``idx`` None: ``x[::stride]``) on the
    device (``asp_student_t_fit``, csrc/student_t.cu: one kernel per EM iteration, the nu equation solved in-kernel;
    spec ``oracle/pcn.py::fit_student_t_diag``).  -> (mu, sigma, nu)"""
    mu, sigma, nu = ops.student_t_fit(x, mu0, sigma0, idx=idx, stride=stride, n_iter=n_iter, nu0=nu0, nu_min=nu_min,
                                      nu_max=nu_max)
    return mu, torch.clamp_min(sigma, 1e-150).contiguous(), nu


class MiniPCNSMC(SMCSampler):
    """``sampler="smc"`` / ``"minipcn_smc"`` (reference ``aspire.py:288-289``); its
    ``beta_tolerance`` default is 1e-6 (``smc/minipcn.py:40``), not the base class's 1e-8."""

    default_beta_tolerance = 1e-6
    default_step_fn = "tpcn"  # smc/minipcn.py:48; built-in densities run it in the fused sweep (params->nu > 0)
