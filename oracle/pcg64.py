"""PCG64 (PCG XSL RR 128/64) - pure-Python restatement of numpy's default bit generator.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference resamples with ``rng.choice(n, size, p=w)`` on a
``numpy.random.Generator`` (``src/aspire/samples.py:1278``); numpy's ``choice``
draws ``rng.random(size)`` and searches the CDF, and ``default_rng`` is PCG64.
``asp_pcg64_uniform`` reproduces that stream on the device from the generator's
(state, inc), so the same seed gives the same ancestors with no host draw or
upload.  PINNED against numpy itself (``tests/test_pcg64_oracle.py``).

Algorithm (numpy/random/src/pcg64/pcg64.h, O'Neill 2014):
    state <- state * M + inc  (mod 2^128),  M = 0x2360ED051FC65DA44385DF649FCCF645
    out   <- rotr64(hi64(state) ^ lo64(state), state >> 122)      (after the step)
    double = (out >> 11) * 2^-53                                   (Generator.random)
Jump-ahead by ``delta`` steps: the affine map composed ``delta`` times,
``state <- A state + C`` with ``(A, C) = (M, inc)^delta`` by repeated squaring
(``pcg_advance_lcg_128``).
"""

from __future__ import annotations

import numpy as np

MULT = 0x2360ED051FC65DA44385DF649FCCF645
MASK128 = (1 << 128) - 1
MASK64 = (1 << 64) - 1


def state_of(rng: np.random.Generator) -> tuple[int, int]:
    """(state, inc) of a Generator whose bit generator is PCG64."""
    st = rng.bit_generator.state
    if st["bit_generator"] != "PCG64":
        raise TypeError(f"not a PCG64 generator: {st['bit_generator']}")
    return int(st["state"]["state"]), int(st["state"]["inc"])


def affine_power(delta: int, inc: int) -> tuple[int, int]:
    """(A, C) with state_{t+delta} = A state_t + C."""
    acc_mult, acc_plus = 1, 0
    cur_mult, cur_plus = MULT, inc & MASK128
    while delta > 0:
        if delta & 1:
            acc_mult = (acc_mult * cur_mult) & MASK128
            acc_plus = (acc_plus * cur_mult + cur_plus) & MASK128
        cur_plus = ((cur_mult + 1) * cur_plus) & MASK128
        cur_mult = (cur_mult * cur_mult) & MASK128
        delta >>= 1
    return acc_mult, acc_plus


def advance(state: int, inc: int, delta: int) -> int:
    a, c = affine_power(delta, inc)
    return (a * state + c) & MASK128


def output(state: int) -> int:
    hi, lo = state >> 64, state & MASK64
    x = hi ^ lo
    rot = state >> 122
    return ((x >> rot) | (x << ((-rot) & 63))) & MASK64


def uniforms(state: int, inc: int, n: int, offset: int = 0) -> np.ndarray:
    """Draws offset .. offset+n-1 of ``Generator.random`` from (state, inc)."""
    s = advance(state, inc, offset)
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        s = (s * MULT + inc) & MASK128
        out[i] = (output(s) >> 11) * 2.0**-53
    return out
