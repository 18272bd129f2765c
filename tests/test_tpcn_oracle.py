"""The tpCN step spec (oracle/pcn.py::tpcn_sweep; parity with minipcn is UNPINNED - the package is absent):
the step is a valid MCMC kernel (leaves a Gaussian posterior invariant), the counter-based gamma sampler has
the right law, and the EM fit recovers the parameters of Student-t data."""
import numpy as np

from oracle import densities as dn
from oracle import pcn, problems


def test_tpcn_leaves_the_target_invariant():
    rng = np.random.default_rng(3)
    d, n, moves, nu = 4, 60000, 8, 4.0
    like, prior, q = problems._gauss(0.5, 0.7, d), problems._gauss(0.0, 1.0, d), problems._gauss(0.0, 1.0, d)
    var = 1.0 / (1.0 / 0.49 + 1.0)
    mean = var * (0.5 / 0.49)
    x = mean + np.sqrt(var) * rng.standard_normal((n, d))
    ll, lp, lq = (dn.evaluate(k, x) for k in (like, prior, q))
    out = pcn.tpcn_sweep(x, ll, lp, lq, like=like, prior=prior, q=q, beta=1.0, mu=np.full(d, 0.2), sigma=np.full(d, 0.9),
                         nu=nu, rho=0.6, n_moves=moves, xi=rng.standard_normal((moves, n, d)),
                         log_u=np.log(1.0 - rng.random((moves, n))), gamma=rng.gamma(0.5 * (nu + d), 1.0, (moves, n)))
    xo = out[0]
    assert abs(xo.mean() - mean) < 5e-3 and abs(xo.var() - var) < 0.01 * var
    kurt = ((xo - xo.mean(0)) ** 4).mean(0) / xo.var(0) ** 2
    assert np.abs(kurt - 3.0).max() < 0.08          # still Gaussian: no tail distortion from the t proposal
    assert 0.3 < out[4].mean() / n < 0.95
    np.testing.assert_allclose(out[1], dn.evaluate(like, xo), rtol=1e-12, atol=1e-12)


def test_gamma_sampler_and_em_fit():
    g = pcn.gamma_mt_f32(5, 3, np.arange(100000), 18.0)
    assert abs(g.mean() - 18.0) < 0.06 and abs(g.var() - 18.0) < 0.5
    # Philox-mode sweep is reproducible and keyed by the element index (sharding-invariant)
    a = pcn.gamma_mt_f32(5, 3, np.arange(50, 60), 7.5)
    np.testing.assert_array_equal(a, pcn.gamma_mt_f32(5, 3, np.arange(0, 100), 7.5)[50:60])
    rng = np.random.default_rng(1)
    scale = np.array([0.5, 1.0, 2.0, 1.0, 0.7, 3.0])
    # multivariate t: one chi-square per row shared by all dimensions
    x = 1.5 + scale * rng.standard_normal((100000, 6)) / np.sqrt(rng.chisquare(5.0, 100000) / 5.0)[:, None]
    mu, sigma, nu = pcn.fit_student_t_diag(x)
    np.testing.assert_allclose(mu, 1.5, atol=0.03)
    np.testing.assert_allclose(sigma, scale, rtol=0.03)
    assert 4.5 < nu < 5.6
