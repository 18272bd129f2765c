"""Philox4x32-10 against the published Random123 known-answer vectors."""
import numpy as np

from oracle import philox


def _kat(ctr, key):
    r = philox.philox4x32(*[np.uint32(c) for c in ctr], key[0], key[1])
    return tuple(int(v) for v in r)


def test_random123_kat_vectors():
    assert _kat((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert _kat((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF)) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert _kat((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_streams_are_sharding_invariant():
    a = philox.normals_f32(5, 9, np.arange(0, 64), 6)
    b = np.concatenate([philox.normals_f32(5, 9, np.arange(0, 20), 6), philox.normals_f32(5, 9, np.arange(20, 64), 6)])
    np.testing.assert_array_equal(a, b)
    u = philox.uniform_f64(1, 2, np.arange(101))
    assert (u >= 0).all() and (u < 1).all()


def test_moments():
    z = philox.normals_f32(1, 2, np.arange(200_000), 4)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    z = philox.normal_f64(1, 3, np.arange(200_000), 3)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
    lu = philox.accept_log_uniform(1, 4, 2, np.arange(100_000))
    assert abs(lu.mean() + 1.0) < 2e-2


def test_seven_rounds_are_a_prefix_of_ten():
    """The sweep's noise streams use Philox4x32-7: the same round function and key schedule, stopped after
    seven rounds - three more rounds with the advanced key give the pinned 10-round output."""
    rng = np.random.default_rng(3)
    ctr = rng.integers(0, 2**32, size=(4, 64), dtype=np.uint64).astype(np.uint32)
    k0, k1 = 0xA4093822, 0x299F31D0
    r7 = philox.philox4x32(*ctr, k0, k1, rounds=7)
    k0_7 = (k0 + 7 * int(philox.W0)) & 0xFFFFFFFF
    k1_7 = (k1 + 7 * int(philox.W1)) & 0xFFFFFFFF
    r10 = philox.philox4x32(*r7, k0_7, k1_7, rounds=3)
    ref = philox.philox4x32(*ctr, k0, k1, rounds=10)
    for a, b in zip(r10, ref):
        np.testing.assert_array_equal(a, b)
    assert philox.NOISE_ROUNDS == 7


def test_noise_stream_statistics():
    """Proposal normals of the sweep (Philox4x32-7 + fp32 Box-Muller): Kolmogorov-Smirnov against N(0,1),
    tail mass, and no correlation between neighbouring particles, dimensions or moves."""
    from scipy import stats

    n, d = 200_000, 8
    z0 = philox.normals_f32(11, 5, np.arange(n), d).astype(np.float64)
    z1 = philox.normals_f32(11, 6, np.arange(n), d).astype(np.float64)  # next move
    assert stats.kstest(z0.ravel(), "norm").pvalue > 1e-3
    tail = np.mean(np.abs(z0) > 3.0)
    assert abs(tail - 2 * stats.norm.sf(3.0)) < 4 * np.sqrt(2 * stats.norm.sf(3.0) / z0.size)
    bound = 5.0 / np.sqrt(n)

    def corr(a, b):
        return abs(np.mean(a * b))

    assert corr(z0[:-1, 0], z0[1:, 0]) < bound          # neighbouring particles (counter c0)
    for j in range(d - 1):
        assert corr(z0[:, j], z0[:, j + 1]) < bound     # neighbouring dimensions (words / blocks)
    assert corr(z0[:, 0], z1[:, 0]) < bound             # consecutive moves (stream c2)
    assert abs(np.mean(z0[:, 0] ** 2 * z0[:, 1] ** 2) - 1.0) < 0.035  # the two normals of a Box-Muller pair
    lu = philox.accept_log_uniform(11, 5, 1, np.arange(n))
    assert stats.kstest(-lu, "expon").pvalue > 1e-3
    assert corr(lu + 1.0, z0[:, 0]) < bound
