"""HOST logic of the samplers in the CPU suite: the ctypes layer (``aspire_b200.ops``) is replaced by the test doubles
of tests/ops_double.py and the very test functions of the GPU suite are run against them on ``device="cpu"`` -
tempering loop against the reference's golden loop, fused / unfused route equivalence, statistics of the
quickstart problem, reference KATs on the containers, initial-sample filtering, argument errors, checkpoint /
resume, tpCN, the preconditioned (bounded / periodic) route and the reference's integration tests.  What this
establishes is that the Python above the C-ABI is right; that the kernels are right is the GPU suite's job."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))

import ops_double  # noqa: E402


@pytest.fixture
def cpu(monkeypatch):
    ops_double.install(monkeypatch)
    return "cpu"


def _mod(name):
    import importlib

    return importlib.import_module(name)


@pytest.mark.parametrize("name,mk,dims", [("quick4", "quickstart_problem", 4), ("rosen4", "rosenbrock_problem", 4)])
def test_smc_loop_matches_reference_loop(cpu, name, mk, dims):
    _mod("test_gpu_samplers").test_smc_loop_matches_reference_loop(cpu, name, mk, dims)


def test_unfused_path_equals_fused(cpu):
    _mod("test_gpu_samplers").test_unfused_path_equals_fused(cpu)


def test_smc_quickstart_statistics(cpu):
    _mod("test_gpu_samplers").test_smc_quickstart_statistics_philox(cpu)


def test_importance_sampler_c2(cpu):
    _mod("test_gpu_samplers").test_importance_sampler_c2(cpu)


@pytest.mark.parametrize("fn", ["test_samples_compute_weights_constant_case",
                                "test_smc_unnormalized_and_normalized_log_weights_and_ratio",
                                "test_smc_log_weights_nan_guard_raises", "test_smc_variance_and_standard_samples",
                                "test_draw_initial_samples_filters_and_errors", "test_sampler_argument_errors",
                                "test_smc_with_tpcn_step"])
def test_reference_kats_and_errors(cpu, fn):
    getattr(_mod("test_gpu_samplers"), fn)(cpu)


def test_smc_resample_shapes_and_reference_indices(cpu, caplog):
    _mod("test_gpu_samplers").test_smc_resample_shapes_beta_same_beta_and_reference_indices(cpu, caplog)


@pytest.mark.parametrize("source", ["dict", "bytes", "file"])
def test_smc_checkpoint_resume_reproduces_the_run(cpu, tmp_path, source):
    _mod("test_gpu_samplers").test_smc_checkpoint_resume_reproduces_the_run(cpu, tmp_path, source)


def test_smc_checkpoint_every_with_default_callback(cpu, tmp_path):
    _mod("test_gpu_samplers").test_smc_checkpoint_every_with_default_callback(cpu, tmp_path)


@pytest.mark.parametrize("step_fn", ["pcn", "tpcn"])
def test_smc_with_bounded_parameters(cpu, step_fn):
    _mod("test_gpu_transforms").test_smc_with_bounded_parameters(cpu, step_fn)


def test_smc_with_a_periodic_parameter(cpu):
    _mod("test_gpu_transforms").test_smc_with_a_periodic_parameter(cpu)


@pytest.mark.parametrize("bounded_to_unbounded", [True, False])
@pytest.mark.parametrize("sampler,kwargs", [
    ("importance", {}),
    ("minipcn_smc", {"adaptive": True, "sampler_kwargs": {"n_steps": 10}}),
])
def test_integration(cpu, bounded_to_unbounded, sampler, kwargs):
    _mod("test_gpu_integration").test_integration(cpu, bounded_to_unbounded, sampler, kwargs)


@pytest.mark.parametrize("value", [float("-inf"), float("nan")])
def test_smc_log_likelihood_with_invalid_value(cpu, value):
    _mod("test_gpu_integration").test_smc_log_likelihood_with_invalid_value(cpu, value)


def test_bounded_parameters_example(cpu, capsys):
    """examples/bounded_parameters.py end to end (bounded + periodic + free parameter, logit flow transform,
    standard preconditioning) on the doubles."""
    import importlib.util
    import math

    import torch

    spec = importlib.util.spec_from_file_location(
        "bounded_parameters", os.path.join(os.path.dirname(__file__), "..", "examples", "bounded_parameters.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    posterior, history = mod.main(device="cpu", n_samples=4000)
    x = posterior.x
    assert history.beta[-1] == 1.0
    assert float(x[:, 0].min()) > 0.0 and float(x[:, 0].max()) < 1.0
    assert float(x[:, 1].min()) >= 0.0 and float(x[:, 1].max()) < 2 * math.pi
    assert abs(math.atan2(float(torch.sin(x[:, 1]).mean()), float(torch.cos(x[:, 1]).mean())) - 0.3) < 0.1
    assert float(x[:, 2].std()) == pytest.approx(3 / math.sqrt(10), rel=0.1)
    assert "log evidence" in capsys.readouterr().out


def test_preconditioned_loop_matches_reference_loop(cpu, golden_dir):
    _mod("test_gpu_transforms").test_preconditioned_loop_matches_reference_loop(cpu, golden_dir)


def test_moment_subsample_step_matches_oracle(cpu):
    _mod("test_gpu_samplers").test_moment_subsample_step_matches_oracle(cpu)


@pytest.mark.parametrize("bounded_transform", ["logit", "probit"])
def test_transformed_flow_is_a_normalised_density_on_the_box(cpu, bounded_transform):
    """``TransformedFlow`` (flows/torch/flows.py:327-387 ``rescale`` / ``inverse_rescale``): the density it reports
    integrates to one over the bounds, and the log density returned with its samples is the one ``log_prob`` gives
    back for them."""
    import numpy as np
    import torch

    from aspire_b200 import transforms as T
    from aspire_b200.flows import DiagGaussianFlow, TransformedFlow
    from aspire_b200.rng import PhiloxRNG

    ft = T.FlowTransform(["a"], prior_bounds={"a": [-2.0, 3.0]}, bounded_transform=bounded_transform,
                         affine_transform=False, device=cpu, eps=1e-12)
    flow = TransformedFlow(DiagGaussianFlow(1, device=cpu, mean=[0.4], sigma=[0.9], rng=PhiloxRNG(5)), ft)
    x = torch.linspace(-2.0, 3.0, 200_001, dtype=torch.float64)[1:-1].reshape(-1, 1)
    q = torch.exp(flow.log_prob(x)).numpy()
    # the two edge cells left out hold up to ~5e-6 of the mass (probit tails are wide in x)
    assert np.trapezoid(q, x[:, 0].numpy()) == pytest.approx(1.0, abs=2e-5)
    xs, lq = flow.sample_and_log_prob(5000)
    assert float(xs.min()) > -2.0 and float(xs.max()) < 3.0
    np.testing.assert_allclose(flow.log_prob(xs).numpy(), lq.numpy(), rtol=1e-9, atol=1e-9)


def _quickstart_sampler(cpu, seed=3, fused=True):
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import SMCSampler

    dims = 4
    prob = tg.quickstart_problem(dims, device=cpu)
    flow = DiagGaussianFlow(dims, device=cpu, mean=[0.0] * dims, sigma=[1.3] * dims, rng=PhiloxRNG(seed))
    like = prob["like"] if fused else (lambda s: prob["like"].log_prob(s.x))
    prior = prob["prior"] if fused else (lambda s: prob["prior"].log_prob(s.x))
    return SMCSampler(like, prior, dims, flow, rng=PhiloxRNG(seed + 1))


def test_fixed_schedule_runs_exactly_n_steps(cpu):
    """adaptive=False: beta advances by 1 / n_steps (smc/base.py:162-165) and ends at exactly 1."""
    import math

    s = _quickstart_sampler(cpu)
    out = s.sample(3000, n_steps=5, adaptive=False, sampler_kwargs=dict(n_steps=4), store_sample_history=False)
    assert len(s.history.beta) == 5 and s.history.beta[-1] == 1.0
    assert s.history.beta[:4] == pytest.approx([0.2, 0.4, 0.6, 0.8])
    assert out.log_evidence == pytest.approx(2 * math.log(math.pi), abs=5 * out.log_evidence_error + 0.05)
    assert s.n_likelihood_evaluations == 3000 + 5 * 4 * 3000


def test_max_n_steps_and_beta_step_clamps(cpu):
    """max_n_steps -> min_beta_step = 1 / max_n_steps, rescaled each step (smc/base.py:198-201) and the loop stops
    after max_n_steps iterations; max_beta_step caps the increase (:202-203)."""
    s = _quickstart_sampler(cpu)
    s.sample(2000, max_n_steps=2, sampler_kwargs=dict(n_steps=3), store_sample_history=False)
    assert len(s.history.beta) <= 2
    s = _quickstart_sampler(cpu)
    s.sample(2000, max_beta_step=0.15, sampler_kwargs=dict(n_steps=3), store_sample_history=False)
    steps = [b - a for a, b in zip([0.0] + s.history.beta[:-1], s.history.beta)]
    assert max(steps) <= 0.15 + 1e-12 and len(s.history.beta) >= 7 and s.history.beta[-1] == 1.0
    s = _quickstart_sampler(cpu)
    s.sample(2000, min_beta_step=0.5, sampler_kwargs=dict(n_steps=3), store_sample_history=False)
    assert len(s.history.beta) <= 2 and s.history.beta[-1] == 1.0


def test_target_efficiency_ramp_history_and_final_samples(cpu):
    """target_efficiency=(lo, hi) ramps with beta (smc/base.py:114-121); store_sample_history keeps a host copy per
    step plus the initial one; n_final_samples resamples at beta = 1 and mutates n_final_steps times (:457-474)."""
    s = _quickstart_sampler(cpu, fused=False)
    out = s.sample(1500, target_efficiency=(0.3, 0.8), target_efficiency_rate=2.0, n_final_samples=700,
                   sampler_kwargs=dict(n_steps=3, n_final_steps=2), store_sample_history=True,
                   resampling="systematic")
    h = s.history
    # recorded at the beta just chosen (smc/base.py:412-414); the search itself aimed at the value for the old beta
    assert h.eff_target == pytest.approx([0.3 + 0.5 * b ** 2.0 for b in h.beta])
    assert h.ess[0] / 1500 == pytest.approx(0.3, abs=0.01)
    assert len(h.sample_history) == len(h.beta) + 1 and h.sample_history[0]["x"].shape == (1500, 4)
    assert out.x.shape == (700, 4) and out.log_likelihood.shape == (700,)
    assert len(h.mcmc_acceptance) == len(h.beta) + 1  # one more mutation for the final samples
    assert out.log_q is None and out.log_w is None


def test_importance_sampler_fields(cpu):
    """ImportanceSampler.sample (importance.py:6-22): every field of Samples filled, prior evaluated before the
    likelihood and visible to it."""
    import math

    import torch

    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import ImportanceSampler

    prob = tg.gaussian_problem(3, device=cpu)
    seen = []

    def log_prior(s):
        seen.append("prior")
        return prob["prior"].log_prob(s.x)

    def log_likelihood(s):
        seen.append("like")
        assert s.log_prior is not None  # docs/recipes.rst:7-27
        return prob["like"].log_prob(s.x)

    flow = DiagGaussianFlow(3, device=cpu, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=PhiloxRNG(8))
    out = ImportanceSampler(log_likelihood, log_prior, 3, flow).sample(20_000)
    assert seen == ["prior", "like"]
    for name in ("log_likelihood", "log_prior", "log_q", "log_w", "weights"):
        assert getattr(out, name).shape == (20_000,)
    assert torch.allclose(out.log_w, out.log_likelihood + out.log_prior - out.log_q)
    assert 0 < float(out.effective_sample_size) <= 20_000 and 0 < out.efficiency <= 1
    assert math.isfinite(out.log_evidence) and out.evidence == pytest.approx(math.exp(out.log_evidence))
    # Z = prod_j N(0.5; 0, 1 + 0.7^2): normalised Gaussian likelihood in x under a standard normal prior
    log_z = 3 * (-0.5 * math.log(2 * math.pi * 1.49) - 0.25 / (2 * 1.49))
    assert out.log_evidence == pytest.approx(log_z, abs=5 * out.log_evidence_error + 0.02)


def test_step_fn_defaults_follow_the_reference(cpu):
    """``MiniPCNSMC`` defaults to the Student-t step like the reference (smc/minipcn.py:48); the base class - the
    north-star kernel - to plain pCN; an explicit choice wins; the tpCN mutation hands (mu, sigma, nu) of the fitted
    Student-t reference to the sweep."""
    from aspire_b200 import targets as tg
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import MiniPCNSMC, SMCSampler

    dims = 4
    prob = tg.quickstart_problem(dims, device=cpu)
    flow = DiagGaussianFlow(dims, device=cpu, mean=[0.0] * dims, sigma=[1.3] * dims, rng=PhiloxRNG(2))
    a = MiniPCNSMC(prob["like"], prob["prior"], dims, flow, rng=PhiloxRNG(3))
    b = SMCSampler(prob["like"], prob["prior"], dims, flow, rng=PhiloxRNG(3))
    a._configure_kernel({})
    b._configure_kernel({})
    assert (a.step_fn, b.step_fn) == ("tpcn", "pcn")
    a._configure_kernel({"step_fn": "pcn"})
    assert a.step_fn == "pcn"
    out = a.sample(400, sampler_kwargs=dict(n_steps=2), store_sample_history=False)  # default kernel again: tpCN
    assert a.step_fn == "tpcn" and len(a.history.tpcn_nu) == len(a.history.beta) and len(out.x) == 400
    assert all(1.0 <= v <= 1e6 for v in a.history.tpcn_nu)
