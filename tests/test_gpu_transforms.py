"""Preconditioning transforms on the device (scope row 8f-4): the CUDA kernel behind
``aspire_b200.transforms`` against the numpy oracle and the reference-generated golden vectors, size-independent
properties at a large N, and the samplers running through bounded / periodic parameters.
Tolerances: 1e-12 relative on transformed values (libm vs CUDA log / exp / erfinv differ in the last ulps and the
logit amplifies them near the clamp), 1e-10 absolute on log|det J| sums of up to 32 terms of size <= 30."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["logit4", "probit3", "mixed32", "affine5"]


def _golden(golden_dir, name):
    from oracle import transforms as ot

    g = np.load(os.path.join(golden_dir, f"transforms_{name}.npz"))
    t = ot.make(g["kind"], g["lower"], g["upper"], str(g["bounded"]) or None, float(g["eps"]),
                g["mean"] if bool(g["affine"]) else None, g["std"] if bool(g["affine"]) else None)
    return g, t


def _apply(cuda, t, x, inverse):
    from aspire_b200 import _lib, ops

    dev = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=cuda)  # noqa: E731
    kind = dev(t["kind"], torch.int32) if t["kind"].any() else None
    bounded = {None: _lib.BOUNDED_NONE, "logit": _lib.BOUNDED_LOGIT, "probit": _lib.BOUNDED_PROBIT}[t["bounded"]]
    y, lj = ops.transform_apply(dev(x), kind, dev(t["lower"]), dev(t["upper"]), bounded, t["eps"],
                                None if t["mean"] is None else dev(t["mean"]),
                                None if t["std"] is None else dev(t["std"]), inverse=inverse)
    return y.cpu().numpy(), lj.cpu().numpy()


@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_reference_golden(cuda, golden_dir, name):
    g, t = _golden(golden_dir, name)
    y, lj = _apply(cuda, t, g["x"], inverse=False)
    np.testing.assert_allclose(y, g["y"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lj, g["log_j"], rtol=0, atol=1e-10)
    x, lji = _apply(cuda, t, g["z_in"], inverse=True)
    np.testing.assert_allclose(x, g["x_inv"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(lji, g["log_j_inv"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("n,dims", [(1, 1), (7, 3), (2049, 16), (100_003, 32), (5000, 64), (300, 255)])
def test_kernel_matches_oracle_on_shapes(cuda, n, dims):
    """Ragged sizes: rows that do not fill a tile, odd and even row strides, one row, one dimension."""
    from oracle import transforms as ot

    rng = np.random.default_rng(n + dims)
    kind = rng.integers(0, 3, dims).astype(np.int32)
    lower = rng.uniform(-5, 0, dims)
    upper = lower + rng.uniform(0.5, 7, dims)
    for bounded in ("logit", "probit"):
        t = ot.make(kind, lower, upper, bounded, 1e-6)
        x = rng.normal(size=(n, dims)) * 3
        for j in np.nonzero(kind == ot.BOUNDED)[0]:
            x[:, j] = rng.uniform(lower[j], upper[j], n)
        t, z_ref = ot.fit(t, x, affine=n > 1)  # one row has no spread to standardise by
        y_ref, lj_ref = ot.forward(t, x)
        y, lj = _apply(cuda, t, x, inverse=False)
        np.testing.assert_allclose(y, y_ref, rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(lj, lj_ref, rtol=1e-12, atol=1e-9)
        z = rng.normal(size=(n, dims))
        x_ref, lji_ref = ot.inverse(t, z)
        xi, lji = _apply(cuda, t, z, inverse=True)
        np.testing.assert_allclose(xi, x_ref, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(lji, lji_ref, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("name", CASES)
def test_composite_transform_fit_matches_reference_golden(cuda, golden_dir, name):
    from aspire_b200 import transforms as T

    g, t = _golden(golden_dir, name)
    d = len(t["kind"])
    params = [f"x{j}" for j in range(d)]
    bounds = {p: [float(t["lower"][j]), float(t["upper"][j])] if t["kind"][j] else [-math.inf, math.inf]
              for j, p in enumerate(params)}
    periodic = [p for j, p in enumerate(params) if t["kind"][j] == 1]
    ct = T.CompositeTransform(params, periodic_parameters=periodic, prior_bounds=bounds,
                              bounded_transform=str(g["bounded"]) or "logit", affine_transform=bool(g["affine"]),
                              device=cuda)
    assert ct._kind_host == t["kind"].tolist()
    z = ct.fit(torch.as_tensor(g["x"], device=cuda))
    np.testing.assert_allclose(z.cpu().numpy(), g["z_fit"], rtol=1e-11, atol=1e-11)
    if bool(g["affine"]):
        np.testing.assert_allclose(ct._mean.cpu().numpy(), g["mean"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(ct._std.cpu().numpy(), g["std"], rtol=1e-12)
    y, lj = ct.forward(torch.as_tensor(g["x"], device=cuda))
    np.testing.assert_allclose(lj.cpu().numpy(), g["log_j"], atol=1e-9)
    x, lji = ct.inverse(torch.as_tensor(g["z_in"], device=cuda))
    np.testing.assert_allclose(x.cpu().numpy(), g["x_inv"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(lji.cpu().numpy(), g["log_j_inv"], atol=1e-9)


def test_single_transform_classes(cuda):
    """tests/test_transforms.py:62-105, 127-169, 334-: wrap-and-invert, logit / probit round trips."""
    from aspire_b200 import transforms as T

    per = T.PeriodicTransform(lower=[0.0, -1.0], upper=[2 * math.pi, 1.0], device=cuda)
    x = torch.tensor([[-0.5, 1.5], [7.0, -3.25], [0.0, 1.0]], dtype=torch.float64, device=cuda)
    y, lj = per.forward(x)
    exp = np.array([[2 * math.pi - 0.5, -0.5], [7.0 - 2 * math.pi, 0.75], [0.0, -1.0]])
    np.testing.assert_allclose(y.cpu().numpy(), exp, rtol=1e-15, atol=1e-15)
    assert torch.equal(lj, torch.zeros(3, dtype=torch.float64, device=cuda))
    assert torch.equal(per.inverse(y)[0], y)
    for cls in (T.LogitTransform, T.ProbitTransform):
        t = cls(lower=[0.0, -2.0, 10.0], upper=[1.0, 3.0, 11.0], device=cuda)
        x = torch.tensor(np.random.default_rng(1).uniform([0, -2, 10], [1, 3, 11], (500, 3)), device=cuda)
        y, lj = t.forward(x)
        x2, lji = t.inverse(y)
        np.testing.assert_allclose(x2.cpu().numpy(), x.cpu().numpy(), rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose((lj + lji).cpu().numpy(), 0.0, atol=1e-9)
    aff = T.AffineTransform(device=cuda)
    x = torch.tensor(np.random.default_rng(2).normal(3.0, 5.0, (4000, 6)), device=cuda)
    z = aff.fit(x)
    np.testing.assert_allclose(z.mean(0).cpu().numpy(), 0.0, atol=1e-12)
    np.testing.assert_allclose(z.std(0, unbiased=False).cpu().numpy(), 1.0, rtol=1e-12)
    np.testing.assert_allclose(aff.inverse(z)[0].cpu().numpy(), x.cpu().numpy(), rtol=1e-12, atol=1e-12)
    lj = aff.forward(x)[1]
    np.testing.assert_allclose(lj.cpu().numpy(), -float(torch.log(x.std(0, unbiased=False)).sum()), rtol=1e-12)


def test_round_trip_at_scale(cuda):
    """Size-independent properties at N = 2e6, D = 32: inverse then forward is the identity on the non-periodic
    dimensions, the two Jacobians cancel, periodic / bounded outputs stay inside their ranges."""
    from aspire_b200 import ops
    from aspire_b200 import transforms as T

    n, d = 2_000_000, 32
    params = [f"x{j}" for j in range(d)]
    bounds = {p: ([0.0, 1.0] if j < 12 else [-math.pi, math.pi] if j < 16 else [-math.inf, math.inf])
              for j, p in enumerate(params)}
    t = T.CompositeTransform(params, periodic_parameters=params[12:16], prior_bounds=bounds, bounded_transform="logit",
                             affine_transform=True, device=cuda)
    x0 = ops.philox_uniform(3, 0, 4096 * d, cuda).reshape(4096, d).contiguous()
    t.fit(x0)
    z = ops.philox_normal(4, 0, n, d, cuda)
    x, lji = t.inverse(z)
    z2, ljf = t.forward(x)
    nonper = [j for j in range(d) if not 12 <= j < 16]
    assert float((z2[:, nonper] - z[:, nonper]).abs().max()) < 1e-8
    assert float((lji + ljf).abs().max()) < 1e-8
    assert float(x[:, :12].min()) > 0.0 and float(x[:, :12].max()) < 1.0
    assert float(x[:, 12:16].min()) >= -math.pi and float(x[:, 12:16].max()) < math.pi
    assert bool(torch.isfinite(lji).all())


def test_argument_errors(cuda):
    from aspire_b200 import ops
    from aspire_b200._lib import AspireB200Error

    x = torch.zeros(4, 300, dtype=torch.float64, device=cuda)
    with pytest.raises(AspireB200Error, match="dims"):
        ops.transform_apply(x, None, None, None, 0, 1e-6)
    y, lj = ops.transform_apply(torch.zeros(0, 3, dtype=torch.float64, device=cuda), None, None, None, 0, 1e-6)
    assert y.shape == (0, 3) and lj.shape == (0,)
    with pytest.raises(AspireB200Error, match="unknown bounded"):
        ops.transform_apply(x[:, :3].contiguous(), None, None, None, 7, 1e-6)


def _box_problem(cuda, dims):
    """Unnormalised Gaussian likelihood inside the unit box with a flat prior: the evidence is a product of
    normal-cdf differences."""
    from scipy.stats import norm

    from aspire_b200 import targets as tg

    mu, sigma = np.linspace(0.2, 0.9, dims), np.linspace(0.15, 0.4, dims)
    like = tg.GaussianDensity(mu, sigma, dims, normalised=False, device=cuda)
    prior = tg.UniformDensity(0.0, 1.0, dims, device=cuda)
    mass = norm.cdf((1 - mu) / sigma) - norm.cdf(-mu / sigma)
    log_z = float(np.sum(np.log(sigma * math.sqrt(2 * math.pi) * mass)))
    a, b = -mu / sigma, (1 - mu) / sigma
    mean = mu + sigma * (norm.pdf(a) - norm.pdf(b)) / mass  # truncated-normal means
    return like, prior, log_z, mean


@pytest.mark.parametrize("step_fn", ["pcn", "tpcn"])
def test_smc_with_bounded_parameters(cuda, step_fn):
    """``Aspire(prior_bounds=...)`` with ``preconditioning_kwargs={"bounded_to_unbounded": True}``: the flow lives on
    the logit image of the box, the pCN moves run in the preconditioned space (smc/minipcn.py:102-131) and every
    particle stays inside the bounds; logZ and the posterior mean match the truncated-normal closed form."""
    from aspire_b200.aspire import Aspire
    from aspire_b200.flows import TransformedFlow
    from aspire_b200.rng import PhiloxRNG

    dims = 4
    like, prior, log_z, mean = _box_problem(cuda, dims)
    params = [f"x{i}" for i in range(dims)]
    a = Aspire(log_likelihood=like, log_prior=prior, dims=dims, parameters=params,
               prior_bounds={p: [0.0, 1.0] for p in params}, flow_class="diag_gaussian", rng=PhiloxRNG(11))
    init = PhiloxRNG(12).uniform(4000 * dims, cuda).reshape(4000, dims).contiguous()
    a.fit(init)
    assert isinstance(a.flow, TransformedFlow)
    samples, hist = a.sample_posterior(
        20_000, sampler="smc", return_history=True, rng=PhiloxRNG(13), store_sample_history=False,
        preconditioning="standard", preconditioning_kwargs=dict(bounded_to_unbounded=True, affine_transform=True),
        sampler_kwargs=dict(n_steps=15, step_fn=step_fn))
    assert hist.beta[-1] == 1.0
    x = samples.x
    assert float(x.min()) > 0.0 and float(x.max()) < 1.0
    assert samples.log_evidence == pytest.approx(log_z, abs=5 * samples.log_evidence_error + 0.03)
    np.testing.assert_allclose(x.mean(0).cpu().numpy(), mean, atol=0.02)
    assert 0.05 < np.mean(hist.mcmc_acceptance) < 0.95


def test_smc_with_a_periodic_parameter(cuda):
    """Default preconditioning (aspire.py:338-341) with a periodic angle: the moves wrap around the boundary.
    L = exp(kappa cos(phi - m)) x N(y; 0, 1) under a flat prior on [0, 2 pi) x N(0, 2^2): Z = I0(kappa) / sqrt(5) ...
    evaluated numerically below."""
    from scipy.special import i0

    from aspire_b200.aspire import Aspire
    from aspire_b200.rng import PhiloxRNG

    kappa, m = 2.0, 0.3  # the mode sits next to the 0 / 2 pi seam
    two_pi = 2 * math.pi

    def log_like(s):
        return kappa * torch.cos(s.x[:, 0] - m) - 0.5 * s.x[:, 1] ** 2

    def log_prior(s):
        inside = (s.x[:, 0] >= 0) & (s.x[:, 0] < two_pi)
        lp = -math.log(two_pi) - 0.5 * (s.x[:, 1] / 2.0) ** 2 - math.log(2.0 * math.sqrt(two_pi))
        return torch.where(inside, lp, torch.full_like(lp, -math.inf))

    # int exp(-y^2/2) N(y; 0, 4) dy = sqrt(2 pi) N(0; 0, 5) = 1 / sqrt(5)
    log_z = math.log(i0(kappa)) + math.log(1.0 / math.sqrt(5.0))
    params = ["phi", "y"]
    a = Aspire(log_likelihood=log_like, log_prior=log_prior, dims=2, parameters=params,
               prior_bounds={"phi": [0.0, two_pi], "y": [-math.inf, math.inf]}, periodic_parameters=["phi"],
               bounded_to_unbounded=False, flow_class="diag_gaussian", rng=PhiloxRNG(21))
    u = PhiloxRNG(22).uniform(8000, cuda).reshape(4000, 2)
    init = torch.stack([u[:, 0] * two_pi, (u[:, 1] - 0.5) * 8.0], dim=1).contiguous()
    a.fit(init)
    samples, hist = a.sample_posterior(20_000, sampler="smc", return_history=True, rng=PhiloxRNG(23),
                                       store_sample_history=False, sampler_kwargs=dict(n_steps=15))
    t = a.sampler.preconditioning_transform
    assert t is not None and t._kind_host == [1, 0]
    phi = samples.x[:, 0]
    assert float(phi.min()) >= 0.0 and float(phi.max()) < two_pi
    # The Gaussian proposal has mass A < 1 inside the prior support and the initial particles are its draws that
    # fall inside (samplers/mcmc.py:49-110) with log_q NOT renormalised, so the estimator (the reference's too)
    # targets Z / A: that is why bounded problems want bounded_to_unbounded=True.
    from scipy.stats import norm

    mu, sg = a.flow.density.mean_np[0], a.flow.density.sigma_np[0]
    mass = norm.cdf((two_pi - mu) / sg) - norm.cdf(-mu / sg)
    assert 0.8 < mass < 0.99
    assert samples.log_evidence == pytest.approx(log_z - math.log(mass), abs=5 * samples.log_evidence_error + 0.03)
    # circular mean of the von Mises posterior
    ang = math.atan2(float(torch.sin(phi).mean()), float(torch.cos(phi).mean()))
    assert ang == pytest.approx(m, abs=0.05)
    # both sides of the seam are populated
    assert float((phi > two_pi - 1.0).double().mean()) > 0.05 and float((phi < 1.0).double().mean()) > 0.2


def test_bounded_parameters_example_runs(cuda, capsys):
    """examples/bounded_parameters.py: bounded (logit) + periodic + free parameter, standard preconditioning."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "bounded_parameters", os.path.join(os.path.dirname(__file__), "..", "examples", "bounded_parameters.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    posterior, history = mod.main()
    x = posterior.x
    assert history.beta[-1] == 1.0
    assert float(x[:, 0].min()) > 0.0 and float(x[:, 0].max()) < 1.0
    assert float(x[:, 1].min()) >= 0.0 and float(x[:, 1].max()) < 2 * math.pi
    assert abs(math.atan2(float(torch.sin(x[:, 1]).mean()), float(torch.cos(x[:, 1]).mean())) - 0.3) < 0.08
    assert float(x[:, 2].std()) == pytest.approx(3 / math.sqrt(10), rel=0.05)
    assert float(x[:, 0].mean()) == pytest.approx(0.899, abs=0.01)  # mean of N(0.95, 0.1) truncated to (0, 1)
    assert "log evidence" in capsys.readouterr().out


def test_preconditioned_loop_matches_reference_loop(cuda, golden_dir):
    """The preconditioned route (``SMCSampler._mutate_preconditioned``: z = T.fit(x), moves on z, Jacobian folded into
    the log prior / log proposal handed to the accept step, x = T^-1(z), logs recomputed) against the golden run of
    the REFERENCE's loop with its own ``CompositeTransform`` and its own ``SMCSampler.log_prob`` as the target
    (oracle/make_golden.py::loop_precond_case): same seed, same numpy Generator call order -> the same beta schedule,
    ESS, evidence increments, acceptance and final particles."""
    import os

    from aspire_b200 import targets as tg
    from aspire_b200 import transforms as T
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.samplers import SMCSampler
    from oracle import problems

    g = np.load(os.path.join(golden_dir, "loop_precond_box4.npz"))
    p = problems.box_problem(4)
    dims = p["dims"]
    params = [f"x{j}" for j in range(dims)]
    like = tg.GaussianDensity(p["like"]["mean"], p["like"]["sigma"], dims, normalised=False, device=cuda)
    prior = tg.UniformDensity(p["prior"]["lo"], p["prior"]["hi"], dims, device=cuda)
    rng = np.random.default_rng(int(g["seed"]))
    flow = DiagGaussianFlow(dims, device=cuda, mean=p["q"]["mean"], sigma=p["q"]["sigma"], rng=rng)
    transform = T.CompositeTransform(params, prior_bounds={q: [0.0, 1.0] for q in params}, bounded_to_unbounded=True,
                                     bounded_transform=str(g["bounded_transform"]), affine_transform=True, device=cuda)
    s = SMCSampler(like, prior, dims, flow, rng=rng, parameters=params, preconditioning_transform=transform)
    out = s.sample(int(g["n"]), adaptive=True, target_efficiency=0.5, beta_tolerance=1e-6,
                   sampler_kwargs=dict(n_steps=int(g["n_moves"]), rho=float(g["rho"]), adapt_rho=False),
                   store_sample_history=False)
    h = s.history
    assert len(h.beta) == len(g["beta"])
    np.testing.assert_allclose(h.beta, g["beta"], rtol=1e-9)
    np.testing.assert_allclose(h.ess, g["ess"], rtol=1e-8)
    np.testing.assert_allclose(h.log_norm_ratio, g["log_norm_ratio"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(h.log_norm_ratio_var, g["log_norm_ratio_var"], rtol=1e-8)
    np.testing.assert_allclose(h.mcmc_acceptance, g["mcmc_acceptance"], rtol=0, atol=2.0 / int(g["n"]))
    assert out.log_evidence == pytest.approx(float(g["log_evidence"]), rel=1e-9)
    x = out.x.cpu().numpy()
    close = np.isclose(x, g["x"], rtol=1e-8, atol=1e-8).all(axis=1)
    assert close.mean() > 0.995
    np.testing.assert_allclose(out.log_likelihood.cpu().numpy()[close], g["log_l"][close], rtol=1e-8, atol=1e-8)
    assert x.min() > 0.0 and x.max() < 1.0


# ---------------------------------------------------------------- flow preconditioning (transforms.py:649-748)
def test_flow_preconditioning_transform_round_trip_and_jacobian(cuda):
    """``FlowPreconditioningTransform``: ``fit`` trains a coupling flow on the device and returns the base-space
    points; ``forward`` / ``inverse`` are each other's inverse to fp32 accuracy with opposite log-Jacobians, the
    log-Jacobian equals log q(x) - log N(z), and the fitted map whitens a correlated, bounded ensemble."""
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.transforms import FlowPreconditioningTransform

    dims, n = 6, 8000
    params = [f"p{i}" for i in range(dims)]
    z0 = PhiloxRNG(1).normal(n, dims, cuda)
    x = z0.clone()
    x[:, 1] = 0.8 * z0[:, 0] + 0.6 * z0[:, 1]           # correlated pair
    x[:, 2] = torch.sigmoid(1.5 * z0[:, 2])              # inside (0, 1): goes through the bounded stage
    x[:, 3] = 3.0 * z0[:, 3] - 2.0
    bounds = {p: [-math.inf, math.inf] for p in params}
    bounds["p2"] = [0.0, 1.0]
    t = FlowPreconditioningTransform(parameters=params, prior_bounds=bounds, bounded_transform="logit",
                                     affine_transform=False, device=cuda, flow_kwargs=dict(seed=3),
                                     fit_kwargs=dict(n_epochs=30, batch_size=500, lr=3e-3))
    with pytest.raises(RuntimeError):
        t.forward(x)
    z = t.fit(x.contiguous())
    assert z.shape == (n, dims) and len(t.history.training_loss) == 30
    assert t.history.training_loss[-1] < t.history.training_loss[0]
    z2, lj = t.forward(x)
    torch.testing.assert_close(z2, z, rtol=0, atol=0)
    xb, lj_inv = t.inverse(z2)
    torch.testing.assert_close(xb, x, rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(lj_inv, -lj, rtol=0, atol=2e-3)
    # the flow whitens: unit scale, the planted correlation mostly gone
    zn = z.cpu().numpy()
    assert np.abs(zn.mean(axis=0)).max() < 0.15
    np.testing.assert_allclose(zn.std(axis=0), 1.0, atol=0.2)
    assert abs(np.corrcoef(zn[:, 0], zn[:, 1])[0, 1]) < 0.5
    # change of variables: the density of x implied by the standard normal in z integrates to ~1 (importance
    # estimate against a broad Gaussian envelope on the unbounded coordinates is not needed: compare instead with
    # the flow's own log_prob through the data transform)
    y, lj_t = t._data_transform.forward(x)
    lq = t.flow.log_prob(y)
    torch.testing.assert_close(lj, lj_t + lq - t._log_normal(z2), rtol=0, atol=1e-9)
    t.flow.check_error()
    assert isinstance(t.new_instance(), FlowPreconditioningTransform) and t.new_instance().flow is None


def test_smc_with_flow_preconditioning(cuda):
    """``sample_posterior(preconditioning="flow")`` (aspire.py:351-366): the moves run in the base space of a flow
    refitted at every tempering step; posterior moments and evidence of a correlated Gaussian problem come out right."""
    from aspire_b200.aspire import Aspire
    from aspire_b200.rng import PhiloxRNG

    dims = 4
    rho_c = 0.9
    cov = torch.eye(dims, dtype=torch.float64, device=cuda)
    cov[0, 1] = cov[1, 0] = rho_c
    prec = torch.linalg.inv(cov)
    log_norm = -0.5 * dims * math.log(2 * math.pi) - 0.5 * float(torch.logdet(cov))

    def log_like(s):  # normalised correlated Gaussian: Z = int L(x) N(x; 0, 5^2 I) dx = N(0; 0, cov + 25 I)
        return log_norm - 0.5 * torch.einsum("ni,ij,nj->n", s.x, prec, s.x)

    def log_prior(s):
        return -0.5 * (s.x / 5.0).pow(2).sum(dim=1) - dims * math.log(5.0 * math.sqrt(2 * math.pi))

    tot = cov + 25.0 * torch.eye(dims, dtype=torch.float64, device=cuda)
    log_z = -0.5 * dims * math.log(2 * math.pi) - 0.5 * float(torch.logdet(tot))
    post_cov = torch.linalg.inv(prec + torch.eye(dims, dtype=torch.float64, device=cuda) / 25.0)
    a = Aspire(log_likelihood=log_like, log_prior=log_prior, dims=dims, parameters=[f"x{i}" for i in range(dims)],
               flow_class="diag_gaussian", rng=PhiloxRNG(5), device=cuda)
    a.fit(PhiloxRNG(6).normal(4000, dims, cuda) * 3.0)
    samples, hist = a.sample_posterior(
        8000, sampler="smc", return_history=True, rng=PhiloxRNG(7), store_sample_history=False,
        preconditioning="flow", preconditioning_kwargs=dict(fit_kwargs=dict(n_epochs=12, lr=3e-3)),
        sampler_kwargs=dict(n_steps=12, step_fn="pcn"))
    from aspire_b200.transforms import FlowPreconditioningTransform

    assert isinstance(a.sampler.preconditioning_transform, FlowPreconditioningTransform)
    assert hist.beta[-1] == 1.0
    x = samples.x
    assert float(x.mean(dim=0).abs().max()) < 0.08
    got = torch.cov(x.T)
    torch.testing.assert_close(got, post_cov, rtol=0, atol=0.12)
    assert samples.log_evidence == pytest.approx(log_z, abs=5 * samples.log_evidence_error + 0.1)
