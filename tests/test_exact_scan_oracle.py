"""Pin the tile-parallel exact scan (oracle/exact_scan.py, the algorithm of
aspire_b200/csrc/scan_exact.cu) against numpy.cumsum - the DEFINITION of the
resampler's CDF (numpy Generator.choice, reference samples.py:1278)."""
import numpy as np
import pytest

from oracle import exact_scan as es


def _cases():
    rng = np.random.default_rng(7)
    n = 300_000
    lw = rng.normal(size=n) * 3
    w = np.exp(lw - lw.max())
    w /= w.sum()
    z = w.copy()
    z[rng.random(n) < 0.3] = 0.0
    onehot = np.zeros(50_000)
    onehot[777] = 1.0
    dom = np.full(50_000, 1e-30)
    dom[5] = 0.5
    return {
        "lognormal": w,
        "uniform_pow2": np.full(1 << 17, 2.0**-17),
        "uniform_3e5": np.full(300_000, 1 / 300_000),
        "zeros30": z,
        "ties": rng.integers(0, 1 << 12, size=100_000).astype(np.float64) * 2.0**-31,
        "onehot": onehot,
        "dominant": dom,
        "tiny_first": np.concatenate([[1e-300, 1e-310, 5e-324], w[:20_000]]),
        "unnormalised": rng.random(100_000) * 1e6,
        "short": rng.random(5),
        "len1": np.array([0.3]),
        "all_zero": np.zeros(5000),
        "subnormal": np.full(5000, 5e-324),
    }


@pytest.mark.parametrize("name", list(_cases()))
def test_exact_cumsum_matches_numpy(name):
    w = _cases()[name]
    out, s = es.exact_cumsum(w)
    np.testing.assert_array_equal(out, np.cumsum(w))
    assert s == np.cumsum(w)[-1]


def test_pair_composition_is_sequential_rounding():
    rng = np.random.default_rng(3)
    w = rng.integers(0, 1 << 10, size=4096).astype(np.float64) * 2.0**-20  # many exact ties in binade 0
    s0 = 1.0 + 2.0**-52 * 3
    out, s, _ = es.scan_tile(w, s0)
    ref = s0 + 0.0
    seq = np.empty_like(w)
    acc = np.float64(s0)
    for i, v in enumerate(w):
        acc = acc + v
        seq[i] = acc
    np.testing.assert_array_equal(out, seq)
    assert ref == s0


@pytest.mark.parametrize("split", [1, 2047, 2048, 150_000])
def test_two_rank_carry(split):
    w = _cases()["lognormal"]
    a, s = es.exact_cumsum(w[:split], n_total=len(w))
    b, _ = es.exact_cumsum(w[split:], s_in=s, n_total=len(w), approx_offset=float(w[:split].sum()),
                           max_in=float(w[:split].max()))
    np.testing.assert_array_equal(np.concatenate([a, b]), np.cumsum(w))


def test_dirty_tiles_are_rare():
    st = {}
    es.exact_cumsum(_cases()["lognormal"], stats=st)
    assert st["n_dirty"] <= 16
    st = {}
    es.exact_cumsum(_cases()["onehot"], stats=st)
    assert st["n_dirty"] <= 2
