"""Sorted-uniform multinomial resampling: the numpy spec (oracle/sorted_resample.py) is a valid multinomial
resampler, and the arithmetic the kernels ship (aspire_b200/csrc/spacing_math.cuh, compiled for the host) equals the
spec bit for bit.  Needs nvcc for the second part (host compilation only, no GPU)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import philox, sorted_resample as sr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    out = tmp_path_factory.mktemp("native") / "libspacing_host.so"
    cmd = [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           os.path.join(ROOT, "tests", "native", "spacing_host.cu"), "-o", str(out)]
    subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=300)
    lib = C.CDLL(str(out))
    lib.spacing_host.restype = C.c_int
    lib.spacing_host.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.threshold_host.restype = C.c_int
    lib.threshold_host.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
    return lib


def test_spacing_is_the_quantised_exponential():
    k = np.concatenate([np.array([0, 1, 2, 2**52, 2**53 - 1, 2**53 - 2, 2**52 - 1, 2**52 + 1], dtype=np.uint64),
                        np.random.default_rng(0).integers(0, 2**53, 200_000, dtype=np.uint64)])
    q = sr.spacing_q(k)
    exact = -np.log1p(-(k.astype(np.float64) * 2.0**-53))  # u = k 2^-53 exactly
    exact[k == 2**53 - 1] = 53 * np.log(2.0)
    # floor of the exact value, except within the evaluation error (3e-14 * 2^26 ~ 2e-6 quanta) of an integer
    target = exact * 2.0**26
    assert np.all(np.abs(q.astype(np.float64) - np.floor(target)) <= 1)
    assert np.mean(q.astype(np.float64) != np.floor(target)) < 1e-4
    assert q[0] == 0 and q.max() < 2**32


def test_sorted_uniforms_are_order_statistics():
    from scipy import stats

    m = 200_000
    u = sr.sorted_uniforms(7, 3, m)
    assert np.all(np.diff(u) >= 0) and u[0] >= 0 and u[-1] < 1
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    # spacings of the order statistics: m * (u_(j+1) - u_(j)) ~ Exp(1), uncorrelated
    d = np.diff(u) * m
    assert stats.kstest(d, "expon").pvalue > 1e-3
    assert abs(np.mean((d[:-1] - 1) * (d[1:] - 1))) < 5 / np.sqrt(m)
    # another stream is another sample
    assert not np.array_equal(u, sr.sorted_uniforms(7, 4, m))


def test_offspring_counts_are_multinomial():
    """Chi-square of the offspring counts against N w, and the same first two moments as numpy's multinomial."""
    from scipy import stats

    rng = np.random.default_rng(5)
    n = 64
    w = rng.dirichlet(np.full(n, 0.7))
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    m = 50_000
    counts = np.zeros((40, n))
    for rep in range(40):
        idx = sr.resample(cdf, 11, rep, m)
        assert np.all(np.diff(idx) >= 0) and idx.min() >= 0 and idx.max() < n
        counts[rep] = np.bincount(idx, minlength=n)
    chi2 = ((counts - m * w) ** 2 / (m * w)).sum(axis=1)  # each ~ chi2(n - 1)
    assert stats.kstest(chi2, "chi2", args=(n - 1,)).pvalue > 1e-3
    var = counts.var(axis=0, ddof=1)
    assert np.all(np.abs(var / (m * w * (1 - w)) - 1) < 1.2)  # 40 replicas: loose, catches a systematic resampler


def test_matches_searchsorted_of_the_sorted_uniforms():
    """idx == searchsorted(cdf, S / total, 'right') except at rounding ties (measure ~1e-16 per slot)."""
    rng = np.random.default_rng(1)
    lw = rng.normal(0, 2.0, 30_000)
    w = np.exp(lw - lw.max())
    w[rng.random(w.size) < 0.3] = 0.0
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    idx = sr.resample(cdf, 3, 1, 30_000)
    ref = np.minimum(np.searchsorted(cdf, sr.sorted_uniforms(3, 1, 30_000), side="right"), cdf.size - 1)
    assert np.mean(idx != ref) < 1e-3
    # degenerate weights
    one_hot = np.zeros(1000)
    one_hot[321] = 1.0
    c = np.cumsum(one_hot)
    assert np.all(sr.resample(c, 3, 2, 5000) == 321)
    flat = np.cumsum(np.full(1000, 1e-3))
    flat /= flat[-1]
    got = np.bincount(sr.resample(flat, 3, 3, 100_000), minlength=1000)
    assert abs(got.mean() - 100) < 1e-9 and got.std() < 15


def test_uniform_integers_are_the_philox_uniform_stream():
    idx = np.arange(1000, dtype=np.uint64)
    k = sr.uniform_k53(9, 2, idx)
    np.testing.assert_array_equal(k.astype(np.float64) * 2.0**-53, philox.uniform_f64(9, 2, idx))


def test_shipped_arithmetic_equals_the_spec(host_lib):
    k = np.concatenate([np.array([0, 1, 2, 2**52, 2**53 - 1, 2**53 - 2, 2**52 - 1, 2**52 + 1, 2**53 - 2**26],
                                 dtype=np.uint64),
                        np.random.default_rng(2).integers(0, 2**53, 500_000, dtype=np.uint64),
                        (2**53 - np.random.default_rng(3).integers(1, 2**20, 50_000)).astype(np.uint64)])
    got = np.empty(k.size, dtype=np.uint32)
    assert host_lib.spacing_host(k.ctypes.data, k.size, got.ctypes.data) == 0
    np.testing.assert_array_equal(got, sr.spacing_q(k))
    c = np.sort(np.random.default_rng(4).random(100_000))
    c[-1] = 1.0
    for total in (12345678901234, 2**53 + 12345, 67108864 * 10**8 + 17):
        t = np.empty(c.size, dtype=np.uint64)
        assert host_lib.threshold_host(c.ctypes.data, c.size, float(total), t.ctypes.data) == 0
        np.testing.assert_array_equal(t, sr.thresholds(c, total))
