"""Philox4x32 (10 rounds; 7 for the sweep noise) and the uniform / normal conversions of the CUDA kernels.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference draws its
randomness from ``numpy.random.Generator`` (PCG64) on the host
(``src/aspire/samples.py:1278``, ``src/aspire/samplers/smc/base.py:77``); the
B200 path's throughput mode uses a counter-based generator so that results do
not depend on the launch geometry or on how particles are sharded over GPUs.
Philox4x32-10 is the published Random123 algorithm (Salmon et al., SC'11); it
is PINNED here against the Random123 known-answer vectors
(``tests/test_philox_oracle.py``).

Counter layout (shared with ``aspire_b200/csrc/philox.cuh``):
    c0, c1 = global particle (or element) index, low / high 32 bits
    c2     = stream id (host-chosen: purpose, SMC iteration, pCN move)
    c3     = block index inside the element (dimension block, or 0xFFFFFFFF
             for the accept/reject uniform)
    key    = (seed low 32, seed high 32)
"""

from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)
ACCEPT_BLOCK = 0xFFFFFFFF
# Rounds of the pCN sweep's noise streams (proposal normals, accept uniforms, tpCN gamma variates):
# Philox4x32-7, the smallest round count Salmon et al. (SC'11, Table 2) report as passing BigCrush;
# it is the first seven rounds of the SAME round function whose 10-round form is pinned against the
# Random123 known-answer vectors in tests/test_philox_oracle.py (the 7-round vectors of Random123's
# kat_vectors file are not reproducible offline, so the pin is on the round function).  The uniform
# and fp64-normal fills (resampling, initial particles) keep 10 rounds.  Kernel: kNoiseRounds in
# aspire_b200/csrc/philox.cuh.
NOISE_ROUNDS = 7


def philox4x32(c0, c1, c2, c3, k0, k1, rounds: int = 10):
    """Vectorised Philox4x32; all inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for r in range(rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _split(index):
    index = np.asarray(index, dtype=np.uint64)
    return (index & MASK32).astype(np.uint32), (index >> np.uint64(32)).astype(np.uint32)


def _key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def uniform_f64(seed: int, stream: int, index) -> np.ndarray:
    """fp64 uniform in [0,1) per element index: Philox block ``index // 2``,
    words (2*(index%2), 2*(index%2)+1): ``((hi<<32 | lo) >> 11) * 2^-53`` - the
    same 53-bit conversion numpy uses for ``Generator.random``."""
    index = np.asarray(index, dtype=np.uint64)
    blk = index >> np.uint64(1)
    c0, c1 = _split(blk)
    k0, k1 = _key(seed)
    r = philox4x32(c0, c1, np.uint32(stream), np.uint32(0), k0, k1)
    odd = (index & np.uint64(1)).astype(bool)
    hi = np.where(odd, r[2], r[0]).astype(np.uint64)
    lo = np.where(odd, r[3], r[1]).astype(np.uint64)
    x = (hi << np.uint64(32)) | lo
    return (x >> np.uint64(11)).astype(np.float64) * 2.0**-53


def _u01_f32(r):
    """fp32 uniform in (0,1]: ``fma(float(r), 2^-32, 2^-33)``."""
    return (r.astype(np.float32) * np.float32(2.0**-32) + np.float32(2.0**-33)).astype(np.float32)


def box_muller_f32(ra, rb):
    """Two fp32 normals from two uint32 words:
    ``u = fma(ra, 2^-32, 2^-33)``, ``s = sqrt(-2 ln2 * log2(u))``,
    ``theta = fma(rb, pi 2^-31, -pi)`` in [-pi, pi], ``(s cos(theta), s sin(theta))``.
    The kernel evaluates log2 / sqrt / sin / cos on the SFU (absolute error ~1e-6
    per normal); numpy float32 here - they agree to ~1e-6, not bit for bit."""
    u = _u01_f32(ra)
    th = (rb.astype(np.float32) * np.float32(1.4629180792671596e-09) + np.float32(-3.14159265358979)).astype(np.float32)
    x = (np.float32(-1.3862943611198906) * np.log2(u)).astype(np.float32)
    s = np.sqrt(x).astype(np.float32)
    th = th.astype(np.float64)
    return (s * np.cos(th).astype(np.float32)).astype(np.float32), (s * np.sin(th).astype(np.float32)).astype(np.float32)


def normal_block_f32(seed: int, stream: int, index, block) -> np.ndarray:
    """Four fp32 normals for (element ``index``, dimension block ``block``):
    returns array [..., 4] = (z0, z1) from words (0,1), (z2, z3) from (2,3)."""
    c0, c1 = _split(index)
    k0, k1 = _key(seed)
    r = philox4x32(c0, c1, np.uint32(stream), np.asarray(block, dtype=np.uint32), k0, k1, rounds=NOISE_ROUNDS)
    z0, z1 = box_muller_f32(r[0], r[1])
    z2, z3 = box_muller_f32(r[2], r[3])
    return np.stack([z0, z1, z2, z3], axis=-1)


def normals_f32(seed: int, stream: int, index, dims: int) -> np.ndarray:
    """[len(index), dims] fp32 normals: dimension j uses block j // 4, lane j % 4."""
    index = np.asarray(index, dtype=np.uint64)
    nblk = (dims + 3) // 4
    z = normal_block_f32(seed, stream, index[:, None], np.arange(nblk, dtype=np.uint32)[None, :])
    return z.reshape(len(index), nblk * 4)[:, :dims]


def accept_log_uniform(seed: int, stream0: int, move: int, index) -> np.ndarray:
    """``log(u)`` for the accept/reject test of move ``move``: fp32 ``ln2 * log2`` (SFU) of word
    ``move & 3`` of block 0xFFFFFFFF of stream ``stream0 + (move & ~3)`` (one Philox
    call serves four moves), promoted to fp64."""
    c0, c1 = _split(index)
    k0, k1 = _key(seed)
    r = philox4x32(c0, c1, np.uint32((stream0 + (move & ~3)) & 0xFFFFFFFF), np.uint32(ACCEPT_BLOCK), k0, k1,
                   rounds=NOISE_ROUNDS)
    lu = (np.float32(0.6931471805599453) * np.log2(_u01_f32(r[move & 3]))).astype(np.float32)
    return lu.astype(np.float64)


def normal_f64(seed: int, stream: int, index, dims: int) -> np.ndarray:
    """[len(index), dims] fp64 normals (used once, for drawing initial
    particles): pair p = j // 2 of an element uses Philox block p // 2... in
    full: block b = j // 4 gives words (r0..r3); 64-bit uniforms
    ``U1 = ((r0<<32|r1)>>11 + 0.5) * 2^-53``, ``U2`` likewise from (r2,r3);
    fp64 Box-Muller ``s = sqrt(-2 log U1)``, ``(s cos(2 pi U2), s sin(2 pi U2))``
    fills dims (4b' , 4b'+1) where each block yields TWO normals: dimension j
    uses block j // 2, component j % 2."""
    index = np.asarray(index, dtype=np.uint64)
    nblk = (dims + 1) // 2
    c0, c1 = _split(index[:, None])
    k0, k1 = _key(seed)
    r = philox4x32(c0, c1, np.uint32(stream), np.arange(nblk, dtype=np.uint32)[None, :], k0, k1)
    a = ((r[0].astype(np.uint64) << np.uint64(32)) | r[1].astype(np.uint64)) >> np.uint64(11)
    b = ((r[2].astype(np.uint64) << np.uint64(32)) | r[3].astype(np.uint64)) >> np.uint64(11)
    u1 = (a.astype(np.float64) + 0.5) * 2.0**-53
    u2 = (b.astype(np.float64) + 0.5) * 2.0**-53
    s = np.sqrt(-2.0 * np.log(u1))
    z = np.stack([s * np.cos(2.0 * np.pi * u2), s * np.sin(2.0 * np.pi * u2)], axis=-1)
    return z.reshape(len(index), nblk * 2)[:, :dims]
