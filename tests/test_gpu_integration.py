"""The reference's integration tests for this path (tests/integration_tests/test_integration.py:12-48, 99-166 with
the fixtures of integration_tests/conftest.py): 2-D Gaussian likelihood under a flat prior on [-10, 10]^2 given as
plain Python callables, ``prior_bounds`` with ``bounded_to_unbounded`` on and off, the importance and the
``minipcn_smc`` samplers, an existing flow handed to ``Aspire``, and a likelihood with a hole of -inf / NaN."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DIMS, MEAN, STD = 2, 2.0, 1.0
PARAMETERS = [f"x_{i}" for i in range(DIMS)]
PRIOR_BOUNDS = {p: [-10, 10] for p in PARAMETERS}
LOG_Z = DIMS * math.log(1 / 20)  # normalised likelihood, prior density 1/20 per dimension, all mass inside


def log_likelihood(samples):
    assert samples.x.shape[-1] == DIMS
    constant = math.log(1 / (STD * math.sqrt(2 * math.pi)))
    return torch.sum(constant - 0.5 * ((samples.x - MEAN) / STD) ** 2, dim=-1)


def log_prior(samples):
    # the reference's fixture writes dims * log(1/10) per dimension; a flat density either way
    assert samples.x.shape[-1] == DIMS
    x = samples.x
    val = torch.where((x >= -10) & (x <= 10), torch.full_like(x, math.log(1 / 20)), torch.full_like(x, -math.inf))
    return torch.sum(val, dim=-1)


def initial_samples(cuda, n=500):
    from aspire_b200 import Samples

    rng = np.random.default_rng(42)
    return Samples(torch.as_tensor(rng.normal(MEAN, STD, size=(n, DIMS)), device=cuda))


def make(cuda, bounded_to_unbounded, like=log_likelihood, **kw):
    from aspire_b200 import Aspire
    from aspire_b200.rng import PhiloxRNG

    return Aspire(log_likelihood=like, log_prior=log_prior, dims=DIMS, parameters=PARAMETERS,
                  prior_bounds=PRIOR_BOUNDS, flow_matching=False, bounded_to_unbounded=bounded_to_unbounded,
                  flow_class="diag_gaussian", rng=PhiloxRNG(3), **kw)


@pytest.mark.parametrize("bounded_to_unbounded", [True, False])
@pytest.mark.parametrize("sampler,kwargs", [
    ("importance", {}),
    ("minipcn_smc", {"adaptive": True, "sampler_kwargs": {"n_steps": 10}}),
    ("smc", {"adaptive": True, "sampler_kwargs": {"n_steps": 10, "step_fn": "tpcn"}}),
])
def test_integration(cuda, bounded_to_unbounded, sampler, kwargs):
    from aspire_b200.rng import PhiloxRNG

    aspire = make(cuda, bounded_to_unbounded)
    aspire.fit(initial_samples(cuda))
    out = aspire.sample_posterior(n_samples=100, sampler=sampler, **kwargs)
    assert out.x.shape == (100, DIMS) and out.parameters == PARAMETERS
    assert bool(torch.isfinite(out.x).all()) and float(out.x.abs().max()) <= 10.0
    assert bool(torch.isfinite(out.log_likelihood).all()) and bool(torch.isfinite(out.log_prior).all())
    assert math.isfinite(float(out.log_evidence))
    # and with enough particles the numbers are right: Z = 1/400, posterior N(2, 1) per dimension
    more = {} if sampler == "importance" else {"rng": PhiloxRNG(9), "store_sample_history": False}
    big = aspire.sample_posterior(n_samples=20_000, sampler=sampler, **kwargs, **more)
    assert float(big.log_evidence) == pytest.approx(LOG_Z, abs=5 * float(big.log_evidence_error) + 0.03)
    if sampler == "importance":
        w = big.weights / big.weights.sum()
        mean = (w[:, None] * big.x).sum(0)
    else:
        mean = big.x.mean(0)
    np.testing.assert_allclose(mean.cpu().numpy(), MEAN, atol=0.06)


@pytest.mark.parametrize("bounded_to_unbounded", [True, False])
def test_init_existing_flow(cuda, bounded_to_unbounded):
    aspire = make(cuda, bounded_to_unbounded)
    aspire.init_flow()
    saved = aspire.flow
    other = make(cuda, bounded_to_unbounded, flow=saved)
    assert other.flow is aspire.flow
    other.flow = saved
    assert other.flow is saved


@pytest.mark.parametrize("value", [-math.inf, math.nan])
def test_smc_log_likelihood_with_invalid_value(cuda, value):
    """A hole (r < 1) where the likelihood is -inf or NaN: initial particles inside it are dropped
    (samplers/mcmc.py:76-95), proposals into it are rejected (NaN -> -inf, smc/base.py:518) and the run completes
    with no particle inside.  (+inf, the reference's third value, would be accepted by construction of the
    Metropolis test and depends on minipcn's unpublished handling: not pinned.)"""
    def with_hole(samples):
        logl = log_likelihood(samples)
        r = torch.linalg.norm(samples.x, dim=1)
        return torch.where(r < 1.0, torch.full_like(logl, value), logl)

    init = initial_samples(cuda)
    assert not bool(torch.isfinite(with_hole(init)).all())
    aspire = make(cuda, False, like=with_hole)
    aspire.fit(init)
    out = aspire.sample_posterior(n_samples=2000, sampler="smc", sampler_kwargs={"n_steps": 10})
    assert out.x.shape == (2000, DIMS)
    assert float(torch.linalg.norm(out.x, dim=1).min()) >= 1.0
    assert bool(torch.isfinite(out.log_likelihood).all())
