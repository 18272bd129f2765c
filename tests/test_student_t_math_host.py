"""The scalar arithmetic the Student-t EM kernel ships (aspire_b200/csrc/student_t_math.cuh; tpCN reference measure,
the reference's default step_fn - smc/minipcn.py:48 - whose package is absent: spec oracle/pcn.py::fit_student_t_diag)
compiled for the HOST: digamma against scipy, the nu equation against the spec's bisection, and a whole EM fit replayed
with the kernel's one-pass sums (about the current mean) and the shipped M step against the spec.  Needs nvcc (host
compilation only, no GPU)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import pcn as opcn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    out = tmp_path_factory.mktemp("native") / "libstudent_t_host.so"
    cmd = [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
           os.path.join(ROOT, "tests", "native", "student_t_host.cu"), "-o", str(out)]
    subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=300)
    lib = C.CDLL(str(out))
    lib.t_host_digamma.argtypes, lib.t_host_digamma.restype = [C.c_double], C.c_double
    lib.t_host_nu.argtypes, lib.t_host_nu.restype = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_double], C.c_double
    lib.t_host_mstep.argtypes = [C.c_double] * 4 + [C.POINTER(C.c_double)] * 2
    lib.t_host_mstep.restype = None
    return lib


def test_digamma_matches_scipy(host_lib):
    from scipy.special import digamma

    x = np.concatenate([np.geomspace(1e-3, 1e7, 4000), np.arange(1, 60) * 0.5])
    got = np.array([host_lib.t_host_digamma(float(v)) for v in x])
    ref = digamma(x)
    np.testing.assert_allclose(got, ref, rtol=2e-14, atol=2e-14)


def test_nu_equation_matches_the_spec(host_lib):
    """For a grid of (nu_old, D, mean(log w - w)): the same root as the spec's 60-step bisection in log nu, including
    the two clamped cases (Gaussian-like data -> nu_max, very heavy tails -> nu_min)."""
    from scipy.special import digamma

    def spec(nu, d, mean_term, lo=1.0, hi=1e6):
        const = 1.0 + mean_term + digamma(0.5 * (nu + d)) - np.log(0.5 * (nu + d))

        def f(v):
            return -digamma(0.5 * v) + np.log(0.5 * v) + const

        if f(hi) > 0:
            return hi
        if f(lo) < 0:
            return lo
        for _ in range(60):
            mid = np.sqrt(lo * hi)
            if f(mid) > 0:
                lo = mid
            else:
                hi = mid
        return float(np.sqrt(lo * hi))

    seen = set()
    for nu in (1.0, 3.7, 10.0, 250.0):
        for d in (2, 32, 64):
            for mean_term in (-1.6, -1.2, -1.05, -1.01, -1.0005, -0.99, -3.0):
                ref = spec(nu, d, mean_term)
                got = host_lib.t_host_nu(nu, d, mean_term, 1.0, 1e6)
                assert got == pytest.approx(ref, rel=1e-9), (nu, d, mean_term)
                seen.add("max" if ref == 1e6 else "min" if ref == 1.0 else "root")
    assert seen == {"max", "min", "root"}


@pytest.mark.parametrize("dims,nu_true", [(3, 4.0), (16, 8.0), (5, 1e9)])
def test_em_fit_replayed_with_the_shipped_m_step(host_lib, dims, nu_true):
    """The kernel's formulation of one EM iteration - sums about the CURRENT mean, mu += A / sum w, sigma from
    B - A^2 / sum w, nu from the shipped equation - iterated 30 times on the host equals the spec's fit."""
    rng = np.random.default_rng(dims)
    n = 20_000
    scale = np.linspace(0.5, 2.0, dims)
    g = rng.chisquare(nu_true, n) / nu_true if nu_true < 1e8 else np.ones(n)
    x = 0.3 + scale * rng.standard_normal((n, dims)) / np.sqrt(g)[:, None]
    mu, sigma, nu = x.mean(axis=0), x.std(axis=0), 10.0
    for _ in range(30):
        v = x - mu
        r2 = ((v / sigma) ** 2).sum(axis=1)
        w = (nu + dims) / (nu + r2)
        sw, sl = w.sum(), (np.log(w) - w).sum()
        a, b = (w[:, None] * v).sum(axis=0), (w[:, None] * v * v).sum(axis=0)
        new_mu, new_sigma = np.empty(dims), np.empty(dims)
        for j in range(dims):
            m_, s_ = C.c_double(mu[j]), C.c_double(0.0)
            host_lib.t_host_mstep(float(sw), float(a[j]), float(b[j]), float(n), C.byref(m_), C.byref(s_))
            new_mu[j], new_sigma[j] = m_.value, s_.value
        nu = host_lib.t_host_nu(nu, dims, float(sl / n), 1.0, 1e6)
        mu, sigma = new_mu, new_sigma
    mu_r, sigma_r, nu_r = opcn.fit_student_t_diag(x)
    np.testing.assert_allclose(mu, mu_r, rtol=1e-10, atol=1e-11)
    np.testing.assert_allclose(sigma, sigma_r, rtol=1e-10)
    assert nu == pytest.approx(nu_r, rel=1e-7)
    if nu_true < 1e8:
        assert 0.6 * nu_true < nu < 1.6 * nu_true
    else:
        assert nu > 20.0  # Gaussian data: nu keeps growing, slowly (30 EM iterations from nu0 = 10)
