"""The PCG64 restatement reproduces numpy's own stream (the reference's resampling
uniforms, ``src/aspire/samples.py:1278`` -> ``Generator.choice`` -> ``Generator.random``)."""
import numpy as np
import pytest

from oracle import pcg64


@pytest.mark.parametrize("seed", [0, 7, 20251104, 2**63 + 5])
def test_uniform_stream_matches_numpy(seed):
    rng = np.random.default_rng(seed)
    state, inc = pcg64.state_of(rng)
    want = rng.random(500)
    np.testing.assert_array_equal(pcg64.uniforms(state, inc, 500), want)
    # a later window through the jump-ahead
    np.testing.assert_array_equal(pcg64.uniforms(state, inc, 40, offset=333), want[333:373])


def test_advance_matches_bit_generator_advance():
    rng = np.random.default_rng(11)
    state, inc = pcg64.state_of(rng)
    for delta in (1, 2, 1000, 10**9 + 7, 2**70 + 3):
        ref = np.random.default_rng(11)
        ref.bit_generator.advance(delta)
        assert pcg64.advance(state, inc, delta) == pcg64.state_of(ref)[0]


def test_choice_uses_these_uniforms():
    """numpy's ``choice(n, size, p=p)`` == searchsorted(cdf, random(size), 'right') on the same stream."""
    p = np.random.default_rng(1).random(1000)
    p /= p.sum()
    a, b = np.random.default_rng(5), np.random.default_rng(5)
    idx = a.choice(1000, size=2000, p=p)
    state, inc = pcg64.state_of(b)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    np.testing.assert_array_equal(idx, np.searchsorted(cdf, pcg64.uniforms(state, inc, 2000), side="right"))
