"""Multinomial resampling with sorted uniforms - numpy spec of ``asp_sorted_spacings`` / ``asp_search_sorted``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference resamples with ``Generator.choice`` (``src/aspire/samples.py:1251-1287``):
``idx = searchsorted(cumsum(w), rng.random(M), 'right')``.  Only the multiset of ancestors matters downstream
(the particle set is exchangeable), and the order statistics of M iid uniforms are distributed as the normalised
partial sums of M + 1 iid exponentials.  The B200 path's throughput mode therefore generates the uniforms already
sorted, from the same Philox stream as ``oracle/philox.py::uniform_f64``:

    k_j   = the 53-bit integer of uniform j (u_j = k_j 2^-53),               j = 0 .. M
    q_j   = floor(-ln(1 - u_j) 2^26)                                          (uint32)
    S_j   = q_0 + ... + q_j,   total = S_M                                    (integers: order-free, shard-free)
    T_k   = rint(cdf_k * float(total))
    idx_j = min(searchsorted(T, S_j, 'right'), N - 1),                        j = 0 .. M - 1

``ln`` uses IEEE add / mul / div only, in the order of ``aspire_b200/csrc/spacing_math.cuh``, so this file
reproduces the kernel's integers exactly.  PARITY: this is the repo's own spec for the throughput mode (the
reference has no sorted variant); with supplied / PCG64 uniforms the kernels keep the reference's draw order
and are pinned to ``Generator.choice`` (``tests/golden``).
"""

from __future__ import annotations

import numpy as np

from . import philox

SPACING_BITS = 26
LN2 = 0.6931471805599453
SQRT2 = 1.4142135623730951


def uniform_k53(seed: int, stream: int, index) -> np.ndarray:
    """The 53-bit integers behind ``philox.uniform_f64`` (same blocks and words)."""
    index = np.asarray(index, dtype=np.uint64)
    blk = index >> np.uint64(1)
    c0, c1 = philox._split(blk)
    k0, k1 = philox._key(seed)
    r = philox.philox4x32(c0, c1, np.uint32(stream), np.uint32(0), k0, k1)
    odd = (index & np.uint64(1)).astype(bool)
    hi = np.where(odd, r[2], r[0]).astype(np.uint64)
    lo = np.where(odd, r[3], r[1]).astype(np.uint64)
    return ((hi << np.uint64(32)) | lo) >> np.uint64(11)


def spacing_q(k) -> np.ndarray:
    """``floor(-ln(1 - k 2^-53) 2^26)`` with the kernel's operation order (spacing_math.cuh::spacing_q)."""
    k = np.asarray(k, dtype=np.uint64)
    m = (np.uint64(1) << np.uint64(53)) - k                      # [1, 2^53]
    p = np.floor(np.log2(m.astype(np.float64))).astype(np.int64)  # exact for integers <= 2^53 (checked below)
    # guard the float log2 at power-of-two boundaries
    p = np.where((np.uint64(1) << p.astype(np.uint64)) > m, p - 1, p)
    p = np.where((np.uint64(1) << (p + 1).astype(np.uint64)) <= m, p + 1, p)
    f = m.astype(np.float64) * np.ldexp(1.0, -p)
    big = f > SQRT2
    f = np.where(big, f * 0.5, f)
    p = np.where(big, p + 1, p)
    s = (f + (-1.0)) / (f + 1.0)
    z = s * s
    poly = np.full_like(z, 1.0 / 15.0)
    for c in (1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0):
        poly = poly * z + c
    s2 = 2.0 * s
    lnf = (s2 * z) * poly + s2
    e = (53 - p).astype(np.float64) * LN2 + (-lnf)
    scaled = e * 67108864.0
    return np.where(scaled > 0.0, np.floor(scaled), 0.0).astype(np.uint64).astype(np.uint32)


def spacings(seed: int, stream: int, m: int) -> np.ndarray:
    """q_0 .. q_m (m + 1 values)."""
    return spacing_q(uniform_k53(seed, stream, np.arange(m + 1, dtype=np.uint64)))


def partial_sums(q: np.ndarray):
    """-> (S_0 .. S_{m-1}, total)."""
    c = np.cumsum(q.astype(np.uint64), dtype=np.uint64)
    return c[:-1], int(c[-1])


def thresholds(cdf, total: int) -> np.ndarray:
    return np.rint(np.asarray(cdf, dtype=np.float64) * float(total)).astype(np.uint64)


def sorted_uniforms(seed: int, stream: int, m: int) -> np.ndarray:
    s, total = partial_sums(spacings(seed, stream, m))
    return s.astype(np.float64) / float(total)


def resample(cdf, seed: int, stream: int, m: int) -> np.ndarray:
    """Ancestors of the m slots (ascending)."""
    s, total = partial_sums(spacings(seed, stream, m))
    t = thresholds(cdf, total)
    return np.minimum(np.searchsorted(t, s, side="right"), len(t) - 1).astype(np.int64)
