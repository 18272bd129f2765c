"""Pin the numpy restatement against the UNMODIFIED reference, live.
Dev container only: skipped where /root/reference does not exist (GPU box)."""
import logging

import numpy as np
import pytest

from oracle import problems, ref_shim
from oracle import smc_numpy as sn

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    ref_shim.import_reference()
    logging.disable(logging.CRITICAL)
    import array_api_compat.numpy as xp
    from aspire.samples import Samples, SMCSamples
    from aspire.utils import effective_sample_size, logsumexp

    yield dict(xp=xp, Samples=Samples, SMCSamples=SMCSamples, ess=effective_sample_size, lse=logsumexp)
    logging.disable(logging.NOTSET)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("beta0,beta", [(0.0, 0.3), (0.2, 0.21), (0.5, 1.0)])
def test_smc_samples_methods(ref, seed, beta0, beta):
    x, ll, lp, lq = problems.draw(problems.rosenbrock_problem(6), 777, seed)
    s = ref["SMCSamples"](x=x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=beta0, xp=ref["xp"], dtype="float64")
    np.testing.assert_array_equal(s.unnormalized_log_weights(beta), sn.unnormalized_log_weights(ll, lp, lq, beta0, beta))
    np.testing.assert_array_equal(s.log_weights(beta), sn.log_weights(ll, lp, lq, beta0, beta))
    np.testing.assert_array_equal(s.log_p_t(beta), sn.log_p_t(ll, lp, lq, beta))
    assert float(s.log_evidence_ratio(beta)) == sn.log_evidence_ratio(ll, lp, lq, beta0, beta)
    assert float(s.log_evidence_ratio_variance(beta)) == sn.log_evidence_ratio_variance(ll, lp, lq, beta0, beta)
    assert float(ref["ess"](s.log_weights(beta))) == float(sn.effective_sample_size(sn.log_weights(ll, lp, lq, beta0, beta)))
    r1 = s.resample(beta, rng=np.random.default_rng(seed + 5))
    idx, xr, *_ = sn.resample(x, ll, lp, lq, beta0, beta, np.random.default_rng(seed + 5))
    np.testing.assert_array_equal(np.asarray(r1.x), xr)
    # and the choice() decomposition itself
    w = sn.resample_weights(ll, lp, lq, beta0, beta)
    u = np.random.default_rng(seed + 5).random(len(w))
    np.testing.assert_array_equal(sn.choice_indices(w, u), idx)


def test_resample_n_samples_and_same_beta(ref):
    x, ll, lp, lq = problems.draw(problems.gaussian_problem(3), 50, 3)
    s = ref["SMCSamples"](x=x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=0.4, xp=ref["xp"], dtype="float64")
    assert s.resample(0.4, rng=np.random.default_rng(1)) is s
    idx, *_ = sn.resample(x, ll, lp, lq, 0.4, 0.4, np.random.default_rng(1))
    assert idx is None
    r = s.resample(0.4, n_samples=20, rng=np.random.default_rng(2))  # uniform weights path
    _, xr, *_ = sn.resample(x, ll, lp, lq, 0.4, 0.4, np.random.default_rng(2), n_samples=20)
    np.testing.assert_array_equal(np.asarray(r.x), xr)


@pytest.mark.parametrize("seed", [4, 5])
def test_compute_weights(ref, seed):
    x, ll, lp, lq = problems.draw(problems.gaussian_problem(5), 400, seed)
    s = ref["Samples"](x=x, log_likelihood=ll, log_prior=lp, log_q=lq, xp=ref["xp"], dtype="float64")
    r = sn.compute_weights(ll, lp, lq)
    np.testing.assert_array_equal(np.asarray(s.log_w), r["log_w"])
    np.testing.assert_array_equal(np.asarray(s.weights), r["weights"])
    for k in ("log_evidence", "evidence", "evidence_error", "log_evidence_error", "effective_sample_size"):
        assert float(getattr(s, k)) == r[k], k


def test_nan_guard(ref):
    s = ref["SMCSamples"](x=np.zeros((3, 1)), log_likelihood=np.array([0.0, np.nan, 1.0]), log_prior=np.zeros(3),
                          log_q=np.zeros(3), beta=0.0, xp=ref["xp"], dtype="float64")
    with pytest.raises(ValueError):
        s.log_weights(0.5)
    with pytest.raises(ValueError):
        sn.log_weights(np.array([0.0, np.nan, 1.0]), np.zeros(3), np.zeros(3), 0.0, 0.5)


@pytest.mark.parametrize("target", [0.5, 0.9, (0.3, 0.8)])
@pytest.mark.parametrize("kw", [dict(), dict(min_beta_step=0.05), dict(max_beta_step=0.01)])
def test_determine_beta(ref, target, kw):
    from aspire.samplers.smc.base import SMCSampler

    x, ll, lp, lq = problems.draw(problems.rosenbrock_problem(4), 1500, 9)
    s = SMCSampler(None, None, 4, None, ref["xp"], dtype="float64", rng=np.random.default_rng(0))
    s.adaptive, s.adaptive_min_beta_step = True, False
    s.target_efficiency, s.target_efficiency_rate = target, 1.0
    for beta0 in (0.0, 0.05, 0.6):
        smc = ref["SMCSamples"](x=x, log_likelihood=ll, log_prior=lp, log_q=lq, beta=beta0, xp=ref["xp"], dtype="float64")
        b_ref, _ = s.determine_beta(smc, beta0, np.nan, kw.get("min_beta_step", 0.0), kw.get("max_beta_step", 1.0), 1e-7)
        b, _ = sn.determine_beta(ll, lp, lq, beta0, target_efficiency=target, beta_tolerance=1e-7, **kw)
        assert b == b_ref
