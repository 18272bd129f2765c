"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the dev container only (needs ``/root/reference``)::

    python -m oracle.make_golden

The reference's own tests hold closed-form known answers for this path
(``tests/test_samples.py:162-185, 637-670, 673-759``; restated in
``tests/test_reference_kats.py``) but no golden index / beta-schedule / logZ
vectors, so these fixtures are produced from the reference itself:
its ``SMCSamples`` methods, ``Samples.compute_weights``,
``SMCSampler.determine_beta`` and the full ``SMCSampler.sample`` loop with an
injected ``mutate`` (our restated pCN, because ``minipcn`` is absent).
"""

from __future__ import annotations

import logging
import os

import numpy as np

from . import pcn, problems, ref_shim
from . import densities as dn

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class _Flow:
    """Duck-typed proposal for the reference sampler: diagonal Gaussian."""

    def __init__(self, q, rng, xp):
        self.q, self.rng, self.xp = q, rng, xp

    def sample_and_log_prob(self, n):
        x = self.q["mean"] + self.q["sigma"] * self.rng.standard_normal((n, len(self.q["mean"])))
        return x, dn.evaluate(self.q, x)

    def log_prob(self, x):
        return dn.evaluate(self.q, np.asarray(x))


def _reference_objects():
    ref_shim.import_reference()
    import array_api_compat.numpy as xp
    from aspire.samplers.smc.base import SMCSampler
    from aspire.samples import Samples, SMCSamples
    from aspire.utils import effective_sample_size

    return xp, SMCSampler, Samples, SMCSamples, effective_sample_size


def reweight_case(name, problem, n, seed, beta0, target_eff=0.5, tol=1e-6, min_step=0.0, max_step=1.0):
    xp, SMCSampler, Samples, SMCSamples, ess_fn = _reference_objects()
    x, ll, lp, lq = problems.draw(problem, n, seed)
    smc = SMCSamples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=beta0, xp=xp, dtype="float64")

    class _S(SMCSampler):
        def mutate(self, particles, beta, n_steps=None):  # pragma: no cover
            return particles

    s = _S(lambda s_: dn.evaluate(problem["like"], s_.x), lambda s_: dn.evaluate(problem["prior"], s_.x),
           problem["dims"], _Flow(problem["q"], np.random.default_rng(0), xp), xp, dtype="float64",
           rng=np.random.default_rng(0))
    s.adaptive = True
    s.adaptive_min_beta_step = False
    s.target_efficiency = target_eff
    s.target_efficiency_rate = 1.0
    beta, _ = s.determine_beta(smc, beta0, np.nan, min_step, max_step, tol)
    lw = smc.log_weights(beta)
    from aspire.utils import logsumexp

    w = np.exp(lw - logsumexp(lw))
    rng_seed = seed + 1000
    out = smc.resample(beta, rng=np.random.default_rng(rng_seed))
    idx = np.random.default_rng(rng_seed).choice(n, size=n, replace=True, p=w)
    assert np.array_equal(out.x, x[idx])
    u = np.random.default_rng(rng_seed).random(n)
    np.savez_compressed(
        os.path.join(OUT, f"reweight_{name}.npz"),
        x=x, log_l=ll, log_p=lp, log_q=lq, beta0=beta0, target_eff=target_eff, tol=tol,
        min_step=min_step, max_step=max_step,
        beta=beta,
        ess=float(ess_fn(smc.log_weights(beta))),
        ess_target=float(ess_fn(smc.log_weights(1.0))),
        log_evidence_ratio=float(smc.log_evidence_ratio(beta)),
        log_evidence_ratio_var=float(smc.log_evidence_ratio_variance(beta)),
        uw=smc.unnormalized_log_weights(beta), w=w, u=u, idx=idx,
    )
    print(f"reweight_{name}: n={n} beta {beta0} -> {beta:.9f}")


def importance_case(name, problem, n, seed):
    xp, _, Samples, _, _ = _reference_objects()
    x, ll, lp, lq = problems.draw(problem, n, seed)
    s = Samples(x=x, log_likelihood=ll, log_prior=lp, log_q=lq, xp=xp, dtype="float64")
    np.savez_compressed(
        os.path.join(OUT, f"importance_{name}.npz"),
        log_l=ll, log_p=lp, log_q=lq, log_w=s.log_w, weights=s.weights,
        log_evidence=float(s.log_evidence), evidence=float(s.evidence),
        evidence_error=float(s.evidence_error), log_evidence_error=float(s.log_evidence_error),
        effective_sample_size=float(s.effective_sample_size),
    )
    print(f"importance_{name}: logZ={float(s.log_evidence):.6f} ess={float(s.effective_sample_size):.1f}")


def loop_case(name, problem, n, seed, n_moves=4, rho=0.5):
    """Full reference tempering loop (smc/base.py:215-488) with the restated
    pCN as ``mutate``; noise from the sampler's numpy Generator."""
    xp, SMCSampler, Samples, SMCSamples, _ = _reference_objects()
    rng = np.random.default_rng(seed)
    like, prior, q = problem["like"], problem["prior"], problem["q"]
    record = dict(mu=[], sigma=[], acc=[])

    class _S(SMCSampler):
        def mutate(self, particles, beta, n_steps=None):
            # reference measure: moments of the (resampled) ensemble
            x = np.asarray(particles.x)
            mu, sigma = x.mean(axis=0), x.std(axis=0)
            m = n_moves
            xi = self.rng.standard_normal((m, len(x), x.shape[1]))
            log_u = np.log(1.0 - self.rng.random((m, len(x))))
            xo, ll, lp, lq, acc = pcn.sweep(
                x, particles.log_likelihood, particles.log_prior, particles.log_q,
                like=like, prior=prior, q=q, beta=beta, mu=mu, sigma=sigma, rho=rho, n_moves=m,
                xi=xi, log_u=log_u)
            record["mu"].append(mu), record["sigma"].append(sigma), record["acc"].append(acc)
            self.history.mcmc_acceptance.append(float(acc.sum()) / (m * len(x)))
            return SMCSamples(x=xo, log_likelihood=ll, log_prior=lp, log_q=lq, beta=beta, xp=self.xp,
                              dtype=self.dtype, parameters=self.parameters)

    s = _S(lambda s_: dn.evaluate(like, np.asarray(s_.x)), lambda s_: dn.evaluate(prior, np.asarray(s_.x)),
           problem["dims"], _Flow(q, rng, xp), xp, dtype="float64", rng=rng)
    s.sampler_kwargs = {}
    out = s.sample(n, adaptive=True, target_efficiency=0.5, beta_tolerance=1e-6, store_sample_history=False)
    h = s.history
    np.savez_compressed(
        os.path.join(OUT, f"loop_{name}.npz"),
        seed=seed, n=n, n_moves=n_moves, rho=rho,
        beta=np.array(h.beta), ess=np.array(h.ess), ess_target=np.array(h.ess_target),
        eff_target=np.array(h.eff_target), log_norm_ratio=np.array(h.log_norm_ratio),
        log_norm_ratio_var=np.array(h.log_norm_ratio_var), mcmc_acceptance=np.array(h.mcmc_acceptance),
        log_evidence=float(out.log_evidence), log_evidence_error=float(out.log_evidence_error),
        x_mean=np.asarray(out.x).mean(axis=0), x_std=np.asarray(out.x).std(axis=0),
        x=np.asarray(out.x), log_l=np.asarray(out.log_likelihood), log_p=np.asarray(out.log_prior),
    )
    print(f"loop_{name}: {len(h.beta)} steps, logZ={float(out.log_evidence):.4f} +/- {float(out.log_evidence_error):.4f}")


def loop_precond_case(name, problem, n, seed, n_moves=5, rho=0.4, bounded_transform="logit"):
    """The reference's tempering loop WITH a preconditioning transform: its own ``CompositeTransform`` (logit +
    affine over the prior bounds), its own ``SMCSampler.log_prob`` (smc/base.py:507-519: T^-1, tempered target,
    + log|det J|) as the target of the restated pCN, driven the way ``MiniPCNSMC.mutate`` drives minipcn
    (smc/minipcn.py:89-131: z = fit(x), chain in z, x = T^-1(z_last), the three logs recomputed at x)."""
    xp, SMCSampler, Samples, SMCSamples, _ = _reference_objects()
    from aspire.transforms import CompositeTransform

    rng = np.random.default_rng(seed)
    like, prior, q = problem["like"], problem["prior"], problem["q"]
    d = problem["dims"]
    params = [f"x{j}" for j in range(d)]
    bounds = {p: [float(prior["lo"][j]), float(prior["hi"][j])] for j, p in enumerate(params)}
    transform = CompositeTransform(parameters=params, prior_bounds=bounds, bounded_to_unbounded=True,
                                   bounded_transform=bounded_transform, affine_transform=True, xp=xp, dtype="float64")

    class _S(SMCSampler):
        def mutate(self, particles, beta, n_steps=None):
            z = np.asarray(self.fit_preconditioning_transform(particles.x))
            mu, sigma = z.mean(axis=0), z.std(axis=0)
            m = n_moves
            xi = self.rng.standard_normal((m, len(z), z.shape[1]))
            log_u = np.log(1.0 - self.rng.random((m, len(z))))
            cur = np.asarray(self.log_prob(z, beta=beta))
            a = np.sqrt(1.0 - rho * rho)
            n_acc = 0
            for t in range(m):
                zp = mu + (a * (z - mu) + (rho * sigma) * xi[t])
                new = np.asarray(self.log_prob(zp, beta=beta))
                log_a = (new - cur) + (pcn.ref_quad(zp, mu, 1.0 / sigma) - pcn.ref_quad(z, mu, 1.0 / sigma))
                acc = log_u[t] < log_a
                z = np.where(acc[:, None], zp, z)
                cur = np.where(acc, new, cur)
                n_acc += int(acc.sum())
            self.history.mcmc_acceptance.append(n_acc / (m * len(z)))
            x = np.asarray(self.preconditioning_transform.inverse(z)[0])
            out = SMCSamples(x=x, beta=beta, xp=self.xp, dtype=self.dtype, parameters=self.parameters)
            out.log_q = np.asarray(self.prior_flow.log_prob(out.x))
            out.log_prior = np.asarray(self.log_prior(out))
            out.log_likelihood = np.asarray(self.log_likelihood(out))
            return out

    s = _S(lambda s_: dn.evaluate(like, np.asarray(s_.x)), lambda s_: dn.evaluate(prior, np.asarray(s_.x)), d,
           _Flow(q, rng, xp), xp, dtype="float64", rng=rng, parameters=params, preconditioning_transform=transform)
    s.sampler_kwargs = {}
    out = s.sample(n, adaptive=True, target_efficiency=0.5, beta_tolerance=1e-6, store_sample_history=False)
    h = s.history
    np.savez_compressed(
        os.path.join(OUT, f"loop_precond_{name}.npz"),
        seed=seed, n=n, n_moves=n_moves, rho=rho, bounded_transform=np.array(bounded_transform),
        beta=np.array(h.beta), ess=np.array(h.ess), log_norm_ratio=np.array(h.log_norm_ratio),
        log_norm_ratio_var=np.array(h.log_norm_ratio_var), mcmc_acceptance=np.array(h.mcmc_acceptance),
        log_evidence=float(out.log_evidence), log_evidence_error=float(out.log_evidence_error),
        x=np.asarray(out.x), log_l=np.asarray(out.log_likelihood), log_p=np.asarray(out.log_prior),
        n_likelihood_evaluations=int(s.n_likelihood_evaluations),
    )
    print(f"loop_precond_{name}: {len(h.beta)} steps, logZ={float(out.log_evidence):.4f}, acceptance "
          f"{np.mean(h.mcmc_acceptance):.3f}, x in [{np.asarray(out.x).min():.4f}, {np.asarray(out.x).max():.4f}]")


def transforms_case(name, d, periodic, bounded, bounded_transform, affine, n, seed):
    """``CompositeTransform`` of the reference itself (numpy namespace, ``src/aspire/transforms.py:142-358``):
    fit / forward / inverse on seeded points that include values on and beyond the bounds (clamped by eps) and
    several periods away."""
    ref_shim.import_reference()
    import array_api_compat.numpy as xp
    from aspire.transforms import CompositeTransform

    from . import transforms as ot

    rng = np.random.default_rng(seed)
    params = [f"x{j}" for j in range(d)]
    bounds = {}
    for j, p in enumerate(params):
        if j in periodic:
            bounds[p] = [0.0, 2 * np.pi] if j % 2 == 0 else [-1.5, 2.5]
        elif j in bounded:
            bounds[p] = [0.0, 1.0] if j % 3 == 0 else [-2.0 - j, 3.0 + 0.5 * j]
        else:
            bounds[p] = [-np.inf, np.inf]
    t = CompositeTransform(parameters=params, periodic_parameters=[params[j] for j in periodic], prior_bounds=bounds,
                           bounded_to_unbounded=True, bounded_transform=bounded_transform, affine_transform=affine,
                           xp=xp, dtype="float64")
    x = rng.normal(size=(n, d)) * 2.0
    for j, p in enumerate(params):
        lo, hi = bounds[p]
        if j in bounded:
            x[:, j] = rng.uniform(lo, hi, n)
            x[0, j], x[1, j] = lo, hi                       # on the bounds
            x[2, j], x[3, j] = lo - 0.25, hi + 0.25         # outside: clamped to eps / 1 - eps
            x[4, j] = lo + 1e-9 * (hi - lo)
        elif j in periodic:
            x[:, j] = rng.uniform(lo - 3 * (hi - lo), hi + 3 * (hi - lo), n)
            x[0, j], x[1, j] = lo, hi
    z_fit = np.asarray(t.fit(x.copy()))
    y, lj = t.forward(x.copy())
    z_in = rng.normal(size=(n, d)) * 1.5
    z_in[0] = 40.0   # sigmoid saturates: clamp on the way back
    z_in[1] = -40.0
    x_inv, lj_inv = t.inverse(z_in.copy())
    desc = ot.from_composite_args(params, [params[j] for j in periodic], bounds, True, bounded_transform, affine)
    out = dict(kind=desc["kind"], lower=desc["lower"], upper=desc["upper"],
               bounded=np.array(desc["bounded"] or ""), eps=desc["eps"], affine=affine, x=x, z_fit=z_fit,
               y=np.asarray(y), log_j=np.asarray(lj), z_in=z_in, x_inv=np.asarray(x_inv), log_j_inv=np.asarray(lj_inv))
    if affine:
        out["mean"], out["std"] = np.asarray(t._affine_transform._mean), np.asarray(t._affine_transform._std)
    np.savez_compressed(os.path.join(OUT, f"transforms_{name}.npz"), **out)
    print(f"transforms_{name}: d={d} n={n} |log_j| max {np.abs(lj).max():.3f}")


def main():
    logging.disable(logging.CRITICAL)
    os.makedirs(OUT, exist_ok=True)
    reweight_case("gauss16", problems.gaussian_problem(16), 4000, 11, 0.0)
    reweight_case("rosen8", problems.rosenbrock_problem(8), 6000, 12, 0.0)
    reweight_case("rosen8_mid", problems.rosenbrock_problem(8), 3000, 13, 0.35, target_eff=0.7)
    reweight_case("gmm6", problems.gmm_problem(6, 4), 5000, 14, 0.1, tol=1e-8)
    reweight_case("quick4_final", problems.quickstart_problem(4), 500, 15, 0.9)
    importance_case("gauss16", problems.gaussian_problem(16), 20000, 21)
    importance_case("quick4", problems.quickstart_problem(4), 2000, 22)
    loop_case("quick4", problems.quickstart_problem(4), 500, 31)
    loop_case("rosen4", problems.rosenbrock_problem(4), 1500, 32, n_moves=6)
    transforms_main()


def transforms_main():
    logging.disable(logging.CRITICAL)
    os.makedirs(OUT, exist_ok=True)
    transforms_case("logit4", 4, periodic=[1], bounded=[0, 3], bounded_transform="logit", affine=True, n=64, seed=41)
    transforms_case("probit3", 3, periodic=[], bounded=[0, 1, 2], bounded_transform="probit", affine=False, n=48, seed=42)
    transforms_case("mixed32", 32, periodic=[10, 11, 12, 13], bounded=list(range(10)), bounded_transform="logit",
                    affine=True, n=257, seed=43)
    transforms_case("affine5", 5, periodic=[], bounded=[], bounded_transform="logit", affine=True, n=33, seed=44)
    loop_precond_case("box4", problems.box_problem(4), 1500, 51)


if __name__ == "__main__":
    main()
