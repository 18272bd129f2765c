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
