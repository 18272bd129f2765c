"""The drop-in samplers on the GPU: the whole tempering loop against the
reference's own loop (golden vectors from /root/reference's SMCSampler.sample),
the reference's SMCSamples / Samples known-answer tests on the CUDA classes, and
the Aspire facade."""
import logging
import math
import os

import numpy as np
import pytest
import scipy.special as sps

from oracle import problems
from oracle import smc_numpy as sn

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _t(a, dev):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


# ------------------------------------------------------------ full loop vs the reference's loop
@pytest.mark.parametrize("name,mk,dims", [("quick4", "quickstart_problem", 4), ("rosen4", "rosenbrock_problem", 4)])
def test_smc_loop_matches_reference_loop(cuda, name, mk, dims):
    """Same seed, same numpy Generator call order as the golden run of the
    reference's SMCSampler.sample (oracle/make_golden.py::loop_case): the beta
    schedule, ESS, evidence increments, acceptance and the final particles agree."""
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.samplers import SMCSampler

    g = np.load(os.path.join(GOLDEN, f"loop_{name}.npz"))
    prob = getattr(tg, mk)(dims)
    rng = np.random.default_rng(int(g["seed"]))
    flow = DiagGaussianFlow(dims, device=cuda, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=rng)
    s = SMCSampler(prob["like"], prob["prior"], dims, flow, rng=rng)
    assert s.fused
    out = s.sample(int(g["n"]), adaptive=True, target_efficiency=0.5, beta_tolerance=1e-6,
                   sampler_kwargs=dict(n_steps=int(g["n_moves"]), rho=float(g["rho"]), adapt_rho=False),
                   store_sample_history=False)
    h = s.history
    assert len(h.beta) == len(g["beta"])
    np.testing.assert_allclose(h.beta, g["beta"], rtol=1e-9, atol=0)
    np.testing.assert_allclose(h.ess, g["ess"], rtol=1e-8)
    np.testing.assert_allclose(h.ess_target, g["ess_target"], rtol=1e-8)
    np.testing.assert_allclose(h.eff_target, g["eff_target"], rtol=1e-12)
    np.testing.assert_allclose(h.log_norm_ratio, g["log_norm_ratio"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(h.log_norm_ratio_var, g["log_norm_ratio_var"], rtol=1e-8)
    np.testing.assert_allclose(h.mcmc_acceptance, g["mcmc_acceptance"], rtol=0, atol=2.0 / int(g["n"]))
    assert out.log_evidence == pytest.approx(float(g["log_evidence"]), rel=1e-9)
    assert out.log_evidence_error == pytest.approx(float(g["log_evidence_error"]), rel=1e-8)
    assert out.log_w is None and out.log_q is None  # samples.py:1295-1305
    x = out.x.cpu().numpy()
    close = np.isclose(x, g["x"], rtol=1e-8, atol=1e-8).all(axis=1)
    assert close.mean() > 0.995  # an accept decision can only flip when |log u - log a| ~ 1e-12
    np.testing.assert_allclose(out.log_likelihood.cpu().numpy()[close], g["log_l"][close], rtol=1e-8, atol=1e-8)


def test_unfused_path_equals_fused(cuda):
    """Python callables (the reference's calling convention) drive the same
    kernels one move at a time and give the same particles as the fused sweep."""
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.samplers import SMCSampler

    dims, n = 6, 3000
    prob = tg.rosenbrock_problem(dims)
    outs = []
    for fused in (True, False):
        rng = np.random.default_rng(77)
        flow = DiagGaussianFlow(dims, device=cuda, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=rng)
        like = prob["like"] if fused else (lambda s: prob["like"].log_prob(s.x))
        prior = prob["prior"] if fused else (lambda s: prob["prior"].log_prob(s.x))
        s = SMCSampler(like, prior, dims, flow, rng=rng)
        assert s.fused is fused
        out = s.sample(n, sampler_kwargs=dict(n_steps=3, adapt_rho=False), beta_tolerance=1e-6,
                       store_sample_history=False)
        outs.append((out, s.history, s.n_likelihood_evaluations))
    (a, ha, na), (b, hb, nb) = outs
    np.testing.assert_allclose(ha.beta, hb.beta, rtol=1e-12)
    np.testing.assert_allclose(a.x.cpu().numpy(), b.x.cpu().numpy(), rtol=1e-11, atol=1e-11)
    assert a.log_evidence == pytest.approx(b.log_evidence, rel=1e-11)
    assert na == nb


def test_smc_quickstart_statistics_philox(cuda):
    """C1 (reference README.md:42-77) in throughput mode: posterior N(0, I/2), logZ = 2 ln(pi)."""
    from aspire_b200 import targets as tg
    from aspire_b200.aspire import Aspire
    from aspire_b200.rng import PhiloxRNG

    dims = 4
    prob = tg.quickstart_problem(dims)
    a = Aspire(log_likelihood=prob["like"], log_prior=prob["prior"], dims=dims, parameters=[f"x{i}" for i in range(dims)],
               flow_class="diag_gaussian", rng=PhiloxRNG(5))
    init = PhiloxRNG(6).normal(2000, dims, cuda) * 1.3
    a.fit(init)
    samples, hist = a.sample_posterior(20_000, sampler="smc", return_history=True, rng=PhiloxRNG(7),
                                       sampler_kwargs=dict(n_steps=20), store_sample_history=False)
    assert hist.beta[-1] == 1.0
    assert samples.log_evidence == pytest.approx(2 * math.log(math.pi), abs=5 * samples.log_evidence_error + 0.02)
    x = samples.x
    assert float(x.mean().abs()) < 0.03
    assert float(x.var()) == pytest.approx(0.5, rel=0.05)
    assert samples.parameters == [f"x{i}" for i in range(dims)]
    assert a.n_likelihood_evaluations > 20_000
    # systematic resampling variant reaches the same answer
    s2 = a.sample_posterior(20_000, sampler="smc", rng=PhiloxRNG(8), sampler_kwargs=dict(n_steps=20),
                            store_sample_history=False, resampling="systematic")
    assert s2.log_evidence == pytest.approx(2 * math.log(math.pi), abs=5 * s2.log_evidence_error + 0.02)


def test_importance_sampler_c2(cuda):
    """C2 shape at reduced N: analytic log-evidence of the Gaussian problem."""
    from aspire_b200 import targets as tg
    from aspire_b200.aspire import Aspire
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG

    dims = 16
    prob = tg.gaussian_problem(dims)
    flow = DiagGaussianFlow(dims, device=cuda, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=PhiloxRNG(3))
    a = Aspire(log_likelihood=prob["like"], log_prior=prob["prior"], dims=dims, flow=flow)
    s = a.sample_posterior(1_000_000, sampler="importance")
    # Z = N(0.5; 0, 1 + 0.49) per dimension
    logz = dims * (-0.5 * math.log(2 * math.pi * 1.49) - 0.5 * 0.25 / 1.49)
    assert s.log_evidence == pytest.approx(logz, abs=6 * s.log_evidence_error + 1e-3)
    assert 0 < s.efficiency <= 1
    assert a.n_likelihood_evaluations == 1_000_000
    # against the oracle on the same particles
    r = sn.compute_weights(s.log_likelihood.cpu().numpy(), s.log_prior.cpu().numpy(), s.log_q.cpu().numpy())
    assert s.log_evidence == pytest.approx(r["log_evidence"], rel=1e-10)
    assert s.effective_sample_size == pytest.approx(r["effective_sample_size"], rel=1e-10)
    assert s.evidence_error == pytest.approx(r["evidence_error"], rel=1e-9)


# ------------------------------------------------------------ reference KATs on the CUDA classes
def test_samples_compute_weights_constant_case(cuda):  # reference tests/test_samples.py:162-185
    from aspire_b200.samples import Samples

    n, d, a = 10, 3, 1.5
    s = Samples(x=np.zeros((n, d)), log_likelihood=np.full(n, a), log_prior=np.zeros(n), log_q=np.zeros(n),
                parameters=[f"p{i}" for i in range(d)], device=cuda)
    assert np.allclose(s.log_w.cpu().numpy(), a)
    assert math.isclose(float(s.evidence), math.exp(a), rel_tol=1e-12)
    assert math.isclose(float(s.log_evidence), a, rel_tol=1e-12)
    assert math.isclose(float(s.evidence_error), 0.0, abs_tol=1e-15)
    assert math.isclose(float(s.log_evidence_error), 0.0, abs_tol=1e-15)
    assert math.isclose(float(s.effective_sample_size), n, rel_tol=1e-12)
    assert math.isclose(float(s.efficiency), 1.0, rel_tol=1e-12)
    assert np.allclose(s.scaled_weights.cpu().numpy(), 1.0)


def test_smc_unnormalized_and_normalized_log_weights_and_ratio(cuda):  # :637-670
    from aspire_b200.samples import SMCSamples

    n, d = 6, 2
    ll = np.linspace(0.0, 1.0, n)
    smc = SMCSamples(x=np.arange(n * d, dtype=float).reshape(n, d), log_likelihood=ll, log_prior=np.zeros(n),
                     log_q=np.zeros(n), beta=0.3, parameters=["a", "b"], device=cuda)
    uw = smc.unnormalized_log_weights(0.7).cpu().numpy()
    assert np.allclose(uw, 0.4 * ll)
    assert math.isclose(float(smc.log_evidence_ratio(0.7)), float(sps.logsumexp(uw) - math.log(n)), rel_tol=1e-12)
    lw = smc.log_weights(0.7).cpu().numpy()
    w = np.exp(lw - sps.logsumexp(lw))
    assert np.isfinite(w).all() and math.isclose(w.sum(), 1.0, rel_tol=1e-12)


def test_smc_log_weights_nan_guard_raises(cuda):  # :673-684
    from aspire_b200.samples import SMCSamples

    smc = SMCSamples(x=np.zeros((3, 1)), log_likelihood=np.array([0.0, np.nan, 1.0]), log_prior=np.zeros(3),
                     log_q=np.zeros(3), beta=0.0, device=cuda)
    with pytest.raises(ValueError):
        smc.log_weights(0.5)


def test_smc_resample_shapes_beta_same_beta_and_reference_indices(cuda, caplog):  # :687-717
    from aspire_b200.samples import SMCSamples

    n, d = 10, 2
    x = np.arange(n * d, dtype=float).reshape(n, d)
    ll, z = np.linspace(0, 1, n), np.zeros(n)
    smc = SMCSamples(x=x, log_likelihood=ll, log_prior=z, log_q=z, beta=0.2, parameters=["a", "b"], device=cuda)
    out = smc.resample(beta=0.8, n_samples=7, rng=np.random.default_rng(42))
    assert isinstance(out, SMCSamples) and out.x.shape == (7, d) and math.isclose(out.beta, 0.8)
    # bit-identical ancestors to the reference's rng.choice for the same seed
    _, xr, *_ = sn.resample(x, ll, z, z, 0.2, 0.8, np.random.default_rng(42), n_samples=7)
    np.testing.assert_array_equal(out.x.cpu().numpy(), xr)
    caplog.clear()
    with caplog.at_level(logging.WARNING):
        same = smc.resample(beta=0.2, n_samples=None, rng=np.random.default_rng(1))
    assert same is smc
    assert any("Resampling with the same beta value" in r.message for r in caplog.records)
    # same beta, different n -> uniform weights (samples.py:1273)
    out2 = smc.resample(beta=0.2, n_samples=25, rng=np.random.default_rng(3))
    _, xr2, *_ = sn.resample(x, ll, z, z, 0.2, 0.2, np.random.default_rng(3), n_samples=25)
    np.testing.assert_array_equal(out2.x.cpu().numpy(), xr2)


def test_smc_variance_and_standard_samples(cuda):  # :720-759
    from aspire_b200.samples import Samples, SMCSamples

    smc = SMCSamples(x=np.arange(10.0).reshape(5, 2), log_likelihood=np.linspace(0, 1, 5), log_prior=np.zeros(5),
                     log_q=np.zeros(5), beta=0.3, device=cuda, parameters=["a", "b"], log_evidence=1.23)
    var = smc.log_evidence_ratio_variance(beta=0.6)
    ref = sn.log_evidence_ratio_variance(np.linspace(0, 1, 5), np.zeros(5), np.zeros(5), 0.3, 0.6)
    assert var >= 0.0 and var == pytest.approx(ref, rel=1e-10)
    std = smc.to_standard_samples()
    assert isinstance(std, Samples) and std.log_w is None
    assert math.isclose(float(std.log_evidence), 1.23, rel_tol=1e-12)
    smc2 = smc[1:3]
    assert isinstance(smc2, SMCSamples) and smc2.x.shape[0] == 2
    assert math.isclose(float(smc2.beta), 0.3) and math.isclose(float(smc2.log_evidence), 1.23)


# ------------------------------------------------------------ initial particles, errors
def test_draw_initial_samples_filters_and_errors(cuda):  # reference tests/test_samplers/test_mcmc/test_mcmc.py:40-82
    import torch

    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import SMCSampler

    dims = 2
    flow = DiagGaussianFlow(dims, device=cuda, mean=[0, 0], sigma=[1, 1], rng=PhiloxRNG(1))
    prior = tg.UniformDensity(-1.0, 1.0, dims)  # -inf outside the box -> those draws are dropped
    like = tg.GaussianDensity(0.0, 1.0, dims)
    s = SMCSampler(like, prior, dims, flow)
    out = s.draw_initial_samples(5000)
    assert len(out) == 5000 and bool(torch.isfinite(out.log_prior).all())
    assert bool((out.x.abs() < 1).all())

    class BadFlow(DiagGaussianFlow):
        def sample_and_log_prob(self, n):
            x, lq = super().sample_and_log_prob(n)
            lq[0] = float("inf")
            return x, lq

    s2 = SMCSampler(like, prior, dims, BadFlow(dims, device=cuda, mean=[0, 0], sigma=[1, 1], rng=PhiloxRNG(2)))
    with pytest.raises(ValueError):
        s2.draw_initial_samples(100)


def test_sampler_argument_errors(cuda):
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import BetaScheduleError, SMCSampler

    dims = 3
    prob = tg.gaussian_problem(dims)
    flow = DiagGaussianFlow(dims, device=cuda, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=PhiloxRNG(1))
    s = SMCSampler(prob["like"], prob["prior"], dims, flow, rng=PhiloxRNG(2))
    with pytest.raises(ValueError):
        s.sample(100, adaptive=False)                       # smc/base.py:348-349
    with pytest.raises(ValueError):
        s.sample(100, target_efficiency=1.5)                # :96-98
    with pytest.raises(ValueError):
        s.sample(100, max_beta_step=1.0)                    # :364-366
    with pytest.raises(ValueError):
        s.sample(100, sampler_kwargs=dict(step_fn="hmc"))   # only 'pcn' and 'tpcn' exist
    # fixed schedule (adaptive=False) runs exactly n_steps trips
    s.sample(500, adaptive=False, n_steps=4, sampler_kwargs=dict(n_steps=2), store_sample_history=False)
    assert s.history.beta == [0.25, 0.5, 0.75, 1.0]
    assert issubclass(BetaScheduleError, RuntimeError)


# ------------------------------------------------------------ checkpoint / resume (smc/base.py:301-314, 521-562)
@pytest.mark.parametrize("source", ["dict", "bytes", "file"])
def test_smc_checkpoint_resume_reproduces_the_run(cuda, tmp_path, source):
    """Resuming from the checkpoint taken after iteration 2 replays the rest of the run bit for bit
    (particles, rng stream position, pCN step and history all come back from the checkpoint)."""
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import MiniPCNSMC

    dims = 4
    prob = tg.rosenbrock_problem(dims)

    def make():
        flow = DiagGaussianFlow(dims, device=cuda, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=PhiloxRNG(5))
        return MiniPCNSMC(prob["like"], prob["prior"], dims, flow, rng=PhiloxRNG(6))

    kw = dict(adaptive=True, target_efficiency=0.5, sampler_kwargs=dict(n_steps=4, rho=0.4), store_sample_history=False)
    states = []
    full = make()
    path = tmp_path / "ckpt.pkl"
    ref = full.sample(4000, checkpoint_callback=states.append, **kw)
    # one state per iteration plus the forced final one (maybe_checkpoint(force=True), smc/base.py:476): it carries
    # the evidence and is written whatever checkpoint_every is
    assert len(states) == len(full.history.beta) + 1 >= 5
    assert states[-1]["iteration"] == states[-2]["iteration"] and states[-1]["meta"]["beta"] == 1.0
    st = states[1]
    assert st["sampler"] == "MiniPCNSMC" and st["iteration"] == 2 and st["meta"]["beta"] == full.history.beta[1]
    assert set(st) >= {"samples", "config", "parameters", "history", "rng_state", "sampler_kwargs"}
    if source == "bytes":
        src = full.serialize_checkpoint(st)
    elif source == "file":
        full.default_file_checkpoint_callback(path)(st)
        src = str(path)
    else:
        src = st
    again = make()
    out = again.sample(4000, resume_from=src, **kw)
    assert again.history.beta == full.history.beta
    assert again.history.log_norm_ratio == full.history.log_norm_ratio
    np.testing.assert_array_equal(out.x.cpu().numpy(), ref.x.cpu().numpy())
    assert out.log_evidence == ref.log_evidence
    # a finished run resumes straight to the end
    done = make()
    fin = done.sample(4000, resume_from=states[-1], **kw)
    np.testing.assert_array_equal(fin.x.cpu().numpy(), ref.x.cpu().numpy())


def test_smc_checkpoint_every_with_default_callback(cuda, tmp_path):
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.samplers import MiniPCNSMC

    dims = 4
    prob = tg.gaussian_problem(dims)
    rng = np.random.default_rng(3)
    flow = DiagGaussianFlow(dims, device=cuda, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=rng)
    s = MiniPCNSMC(prob["like"], prob["prior"], dims, flow, rng=rng)
    path = tmp_path / "run.pkl"
    s.sample(2000, checkpoint_every=2, checkpoint_file_path=str(path),
             sampler_kwargs=dict(n_steps=3), store_sample_history=False)
    assert s.last_checkpoint_state is not None and s.last_checkpoint_state["iteration"] % 2 == 0
    loaded = s.load_checkpoint_from_file(path)
    assert loaded["iteration"] == s.last_checkpoint_state["iteration"]
    assert loaded["rng_state"]["bit_generator"] == "PCG64"  # numpy generator state, as in the reference
    if _no_h5py():  # the reference's HDF5 container needs h5py: refused loudly, never silently dropped
        with pytest.raises(RuntimeError):
            s.default_file_checkpoint_callback(tmp_path / "run.h5")


def _no_h5py():
    try:
        import h5py  # noqa: F401
        return False
    except ImportError:
        return True


@pytest.mark.parametrize("name,dims", [("rosenbrock", 32), ("gmm", 64), ("gauss", 16), ("rosenbrock", 8)])
def test_semi_fused_move_equals_unfused_and_fused(cuda, name, dims):
    """asp_pcn_propose_eval + asp_pcn_accept_scalar (csrc/pcn_eval.cu) walk the same trajectory as
    asp_pcn_propose / density kernels / asp_pcn_accept and as the fused sweep."""
    import torch

    from aspire_b200 import ops
    from aspire_b200 import targets as tg

    mk = dict(rosenbrock=tg.rosenbrock_problem, gmm=tg.gmm_problem, gauss=tg.gaussian_problem)[name]
    pg = mk(dims)
    n, moves, beta, rho = 5000, 4, 0.3, 0.35
    x0 = ops.philox_normal(8, 0, n, dims, cuda, mean=pg["q"].mean, sigma=pg["q"].sigma)
    logs0 = [pg[k].log_prob(x0) for k in ("like", "prior", "q")]
    mom = ops.weighted_moments(x0)
    mu, sigma = ops.moments_to_mu_sigma(mom, dims)
    assert ops.pcn_propose_eval_supported(dims, pg["like"], pg["prior"])
    results = []
    for mode in ("semi", "unfused"):
        x = x0.clone()
        ll, lp, lq = (t.clone() for t in logs0)
        prm = ops.make_pcn_params(beta, rho, mu, sigma, moves, seed=31, stream0=5, index0=17)
        acc = torch.zeros(moves, dtype=torch.int64, device=cuda)
        for t in range(moves):
            if mode == "semi":
                xp, ll2, lp2, dr = ops.pcn_propose_eval(x, pg["like"], pg["prior"], prm, t)
                np.testing.assert_allclose(ll2.cpu().numpy(), pg["like"].log_prob(xp).cpu().numpy(), rtol=1e-12, atol=1e-11)
                np.testing.assert_allclose(lp2.cpu().numpy(), pg["prior"].log_prob(xp).cpu().numpy(), rtol=1e-12, atol=1e-11)
                ops.pcn_accept_scalar(x, ll, lp, lq, xp, ll2, lp2, pg["q"].log_prob(xp), dr, prm, t, acc)
            else:
                xp = ops.pcn_propose(x, prm, t)
                ops.pcn_accept(x, ll, lp, lq, xp, pg["like"].log_prob(xp), pg["prior"].log_prob(xp),
                               pg["q"].log_prob(xp), prm, t, acc)
        results.append((x.cpu().numpy(), ll.cpu().numpy(), acc.cpu().numpy()))
    prm = ops.make_pcn_params(beta, rho, mu, sigma, moves, seed=31, stream0=5, index0=17)
    fused = ops.pcn_sweep(x0, *logs0, None, pg["like"], pg["prior"], pg["q"], prm)
    (xs, ls, as_), (xu, lu, au) = results
    np.testing.assert_array_equal(as_, au)
    np.testing.assert_array_equal(as_, fused[4].cpu().numpy())
    np.testing.assert_allclose(xs, xu, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(xs, fused[0].cpu().numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ls, fused[1].cpu().numpy(), rtol=1e-11, atol=1e-10)
    assert 0 < as_.sum() < moves * n


def test_smc_with_tpcn_step(cuda):
    """step_fn='tpcn' (the reference's default kernel): EM fit of the Student-t reference on the device matches
    the oracle's, and the sampler recovers the analytic posterior and evidence of the quickstart problem."""
    import torch

    from aspire_b200 import ops
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import MiniPCNSMC
    from aspire_b200.samplers.smc import fit_student_t_diag
    from oracle import pcn as opcn

    rng = np.random.default_rng(2)
    xs = 0.3 + np.array([0.5, 1.0, 2.0]) * rng.standard_normal((60000, 3)) / np.sqrt(rng.chisquare(4.0, 60000) / 4.0)[:, None]
    xt = torch.as_tensor(xs, device=cuda)
    mu, sg, nu = fit_student_t_diag(xt, xt.mean(dim=0), xt.std(dim=0, unbiased=False))
    mu_r, sg_r, nu_r = opcn.fit_student_t_diag(xs)
    np.testing.assert_allclose(mu.cpu().numpy(), mu_r, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sg.cpu().numpy(), sg_r, rtol=1e-9)
    assert nu == pytest.approx(nu_r, rel=1e-6) and 3.0 < nu < 5.5
    # rows selected in place through the ancestor index with a stride (what the mutation hands over)
    idx = np.sort(rng.integers(0, 60000, 50001))
    rows = xs[idx[::3]]
    mu, sg, nu = fit_student_t_diag(xt, torch.as_tensor(rows.mean(axis=0), device=cuda),
                                    torch.as_tensor(rows.std(axis=0), device=cuda), idx=torch.as_tensor(idx, device=cuda),
                                    stride=3)
    mu_r, sg_r, nu_r = opcn.fit_student_t_diag(rows)
    np.testing.assert_allclose(mu.cpu().numpy(), mu_r, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sg.cpu().numpy(), sg_r, rtol=1e-9)
    assert nu == pytest.approx(nu_r, rel=1e-6)

    dims = 4
    prob = tg.quickstart_problem(dims)
    flow = DiagGaussianFlow(dims, device=cuda, mean=np.zeros(dims), sigma=np.full(dims, 1.5), rng=PhiloxRNG(21))
    s = MiniPCNSMC(prob["like"], prob["prior"], dims, flow, rng=PhiloxRNG(22))
    out = s.sample(60000, sampler_kwargs=dict(n_steps=8, step_fn="tpcn"), store_sample_history=False)
    x = out.x.cpu().numpy()
    assert np.abs(x.mean(axis=0)).max() < 0.02
    np.testing.assert_allclose(x.var(axis=0), 0.5, rtol=0.05)
    assert out.log_evidence == pytest.approx(2.0 * math.log(math.pi), abs=0.05)  # C1: logZ = (D/2) ln pi
    assert len(s.history.tpcn_nu) == len(s.history.beta) and all(v >= 1.0 for v in s.history.tpcn_nu)


def test_readme_quickstart_example_runs(cuda, capsys):
    """examples/quickstart.py = the reference README quickstart (README.md:42-77) with torch callables."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("quickstart", os.path.join(os.path.dirname(__file__), "..", "examples",
                                                                          "quickstart.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()
    out = capsys.readouterr().out
    z = float(out.split("log evidence")[1].split()[0])
    assert abs(z - 2 * math.log(math.pi)) < 0.5      # 500 particles: a loose statistical check


def test_moment_subsample_step_matches_oracle(cuda):
    """More particles than ``moment_particles``: the pCN reference measure is fitted to every
    (n // moment_particles)-th offspring.  One tempering step (host generator, the reference's draw order) against
    numpy with the same strided subsample: same beta, same ancestors, same particles."""
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.history import SMCHistory
    from aspire_b200.samplers import SMCSampler
    from aspire_b200.samples import SMCSamples
    from oracle import pcn as opcn

    dims, n, moves, rho, m_use = 8, 20_000, 4, 0.4, 4096
    po, pg = problems.rosenbrock_problem(dims), tg.rosenbrock_problem(dims, device=cuda)
    x, ll, lp, lq = problems.draw(po, n, 21)
    rng = np.random.default_rng(9)
    flow = DiagGaussianFlow(dims, device=cuda, mean=pg["q"].mean_np, sigma=pg["q"].sigma_np, rng=rng)
    s = SMCSampler(pg["like"], pg["prior"], dims, flow, rng=rng)
    s._configure_kernel(dict(n_steps=moves, rho=rho, adapt_rho=False, moment_particles=m_use))
    s.target_efficiency, s.target_efficiency_rate, s.adaptive, s.adaptive_min_beta_step = 0.5, 1.0, True, False
    s.max_beta_step, s.resampling, s.history = 1.0, "multinomial", SMCHistory()
    start = SMCSamples(_t(x, cuda), log_likelihood=_t(ll, cuda), log_prior=_t(lp, cuda), log_q=_t(lq, cuda), beta=0.0,
                       device=cuda)
    out, beta, _ = s.tempering_step(start, 0.0, math.nan, 0.0, 1e-6)

    ref_rng = np.random.default_rng(9)
    beta_ref, _ = sn.determine_beta(ll, lp, lq, 0.0, beta_tolerance=1e-6)
    assert beta == beta_ref
    idx = sn.choice_indices(sn.resample_weights(ll, lp, lq, 0.0, beta), ref_rng.random(n))
    xi = ref_rng.standard_normal((moves, n, dims))
    log_u = np.log(1.0 - ref_rng.random((moves, n)))
    stride = n // m_use
    sub = idx[::stride]
    mu, sigma = opcn.weighted_moments(x[sub], np.ones(len(sub)))
    ref = opcn.sweep(x[idx], ll[idx], lp[idx], lq[idx], like=po["like"], prior=po["prior"], q=po["q"], beta=beta, mu=mu,
                     sigma=sigma, rho=rho, n_moves=moves, xi=xi, log_u=log_u)
    got = out.x.cpu().numpy()
    close = np.isclose(got, ref[0], rtol=1e-10, atol=1e-10).all(axis=1)
    assert close.mean() > 0.999
    assert s.history.mcmc_acceptance[-1] == pytest.approx(ref[4].sum() / (moves * n), abs=2.0 / n)
    np.testing.assert_allclose(out.log_likelihood.cpu().numpy()[close], ref[1][close], rtol=1e-9, atol=1e-9)
