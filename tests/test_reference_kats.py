"""The closed-form known answers the reference's own tests hold for this path
(/root/reference/tests/test_samples.py:162-185, 637-684, 720-731), restated
against the numpy oracle.  (The same KATs run against the CUDA path in
tests/test_gpu_parity.py.)"""
import math

import numpy as np
import pytest
import scipy.special as sps

from oracle import smc_numpy as sn


def test_compute_weights_constant_case():  # tests/test_samples.py:162-185
    n, a = 10, 1.5
    r = sn.compute_weights(np.full(n, a), np.zeros(n), np.zeros(n))
    assert np.allclose(r["log_w"], a)
    assert math.isclose(r["evidence"], math.exp(a), rel_tol=1e-12)
    assert math.isclose(r["log_evidence"], a, rel_tol=1e-12)
    assert math.isclose(r["evidence_error"], 0.0, abs_tol=1e-15)
    assert math.isclose(r["log_evidence_error"], 0.0, abs_tol=1e-15)
    assert math.isclose(r["effective_sample_size"], n, rel_tol=1e-12)


def test_unnormalized_and_normalized_log_weights_and_ratio():  # :637-670
    n = 6
    ll, lp, lq = np.linspace(0.0, 1.0, n), np.zeros(n), np.zeros(n)
    b0, b1 = 0.3, 0.7
    uw = sn.unnormalized_log_weights(ll, lp, lq, b0, b1)
    assert np.allclose(uw, (b1 - b0) * (ll + lp))
    assert math.isclose(sn.log_evidence_ratio(ll, lp, lq, b0, b1), sps.logsumexp(uw) - math.log(n), rel_tol=1e-12)
    lw = sn.log_weights(ll, lp, lq, b0, b1)
    w = np.exp(lw - sps.logsumexp(lw))
    assert np.isfinite(w).all() and math.isclose(w.sum(), 1.0, rel_tol=1e-12)


def test_log_weights_nan_guard_raises():  # :673-684
    with pytest.raises(ValueError):
        sn.log_weights(np.array([0.0, np.nan, 1.0]), np.zeros(3), np.zeros(3), 0.0, 0.5)


def test_variance_delta_method_nonnegative():  # :720-731
    v = sn.log_evidence_ratio_variance(np.linspace(0, 1, 5), np.zeros(5), np.zeros(5), 0.3, 0.6)
    assert v >= 0.0 or np.isnan(v)


def test_resample_shapes_and_same_beta():  # :687-717
    n, d = 10, 2
    x = np.arange(n * d, dtype=float).reshape(n, d)
    ll, z = np.linspace(0, 1, n), np.zeros(n)
    idx, xr, *_ = sn.resample(x, ll, z, z, 0.2, 0.8, np.random.default_rng(42), n_samples=7)
    assert xr.shape == (7, d) and idx.dtype == np.int64
    idx2, x2, *_ = sn.resample(x, ll, z, z, 0.2, 0.2, np.random.default_rng(42))
    assert idx2 is None and x2 is x
