"""Sequential-order-exact fp64 prefix sum, computed tile-parallel.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The DEFINITION of the
result is ``numpy.cumsum`` (strictly sequential fp64 adds - what
``numpy.random.Generator.choice`` does to ``p`` at ``src/aspire/samples.py:1278``
of the reference).  This module restates the tile-parallel algorithm the CUDA
kernels in ``aspire_b200/csrc/scan_exact.cu`` use, stage by stage, and is
pinned against ``numpy.cumsum`` in ``tests/test_exact_scan_oracle.py``.

Idea.  While the running sum ``s`` stays inside one binade
``[2^e, 2^(e+1))`` it is ``m * u`` with ``u = 2^(e-52)`` and integer ``m``, and
``fl(s + w) = (m + round_half_even(w / u)) * u``.  Ignoring ties this is integer
addition, hence associative.  A tie (``frac(w/u) == 1/2``) rounds by the parity
of the incoming ``m``; an element is then the pair ``(inc if m even, inc if m
odd)`` and pairs compose associatively: ``h_p = f_p + g_{(p + f_p) & 1}``.
Binade crossings (``m >= 2^53``) are resolved with one genuine fp64 add.
Zero and subnormal running sums use ``e = -1022`` (``u = 2^-1074``), where the
same integer rule holds for ``m`` in ``[0, 2^53)``.
"""

from __future__ import annotations

import numpy as np

TILE = 2048
TWO53 = np.int64(1) << np.int64(53)
SAT_MAX = np.int64(1) << np.int64(60)
E_MIN = -1022


def binade(s: float) -> int:
    """Exponent ``e`` with ``2^e <= s < 2^(e+1)``, clamped to -1022 (0 and
    subnormals share the lowest binade)."""
    if not (s > 0.0):
        return E_MIN
    return max(int(np.frexp(s)[1]) - 1, E_MIN)


def _sat(a):
    return np.minimum(a, SAT_MAX)


def element_maps(w: np.ndarray, e: int):
    """Per-element increment pair under binade ``e``: (inc if incoming m even,
    inc if incoming m odd).  Elements with ``w / u >= 2^53`` saturate."""
    with np.errstate(over="ignore", invalid="ignore"):
        q = np.ldexp(w, 52 - e)
    big = ~(q < 9007199254740992.0)
    qq = np.where(big, 0.0, q)
    fl = np.floor(qq)
    frac = qq - fl
    n = fl.astype(np.int64)
    up = (frac > 0.5).astype(np.int64)
    tie = frac == 0.5
    odd = (n & 1) == 1
    d0 = n + up + (tie & odd)
    d1 = n + up + (tie & ~odd)
    d0 = np.where(big, TWO53, d0)
    d1 = np.where(big, TWO53, d1)
    return d0, d1


def compose(a0, a1, b0, b1):
    """``a`` then ``b``."""
    h0 = a0 + np.where((a0 & 1) == 0, b0, b1)
    h1 = a1 + np.where((a1 & 1) == 1, b0, b1)
    return _sat(h0), _sat(h1)


def pair_scan(d0: np.ndarray, d1: np.ndarray):
    """Inclusive Hillis-Steele scan of increment pairs."""
    f0, f1 = d0.copy(), d1.copy()
    n = len(f0)
    step = 1
    while step < n:
        h0, h1 = compose(f0[:-step], f1[:-step], f0[step:], f1[step:])
        f0[step:], f1[step:] = h0, h1
        step *= 2
    return f0, f1


def tile_aggregate(w_tile: np.ndarray, e: int):
    """Aggregate pair of one tile under binade ``e`` (stage 3 of the GPU path)."""
    d0, d1 = element_maps(w_tile, e)
    f0, f1 = pair_scan(d0, d1)
    return int(f0[-1]), int(f1[-1])


def scan_tile(w_tile: np.ndarray, s: float):
    """Exact sequential cumsum of one tile starting from running sum ``s``.
    Phase loop: integer scan under the current binade up to the first
    crossing, one genuine add, repeat.  Returns (out, s_end, n_phases)."""
    n = len(w_tile)
    out = np.empty(n, dtype=np.float64)
    pos = 0
    phases = 0
    while pos < n:
        phases += 1
        e = binade(s)
        m0 = np.int64(np.ldexp(s, 52 - e))
        d0, d1 = element_maps(w_tile[pos:], e)
        f0, f1 = pair_scan(d0, d1)
        m = m0 + (f0 if (m0 & 1) == 0 else f1)
        cross = m >= TWO53
        j = int(np.argmax(cross)) if cross.any() else len(m)
        if j > 0:
            out[pos:pos + j] = np.ldexp(m[:j].astype(np.float64), e - 52)
            s = float(out[pos + j - 1])
        pos += j
        if pos < n:
            s = float(np.float64(s) + np.float64(w_tile[pos]))
            out[pos] = s
            pos += 1
    return out, s, phases


# ---- the staged (GPU-shaped) algorithm ------------------------------------
def plan(tile_sums: np.ndarray, n_total: int, offset: float = 0.0, tile: int = TILE,
         tile_max: np.ndarray | None = None, max_in: float = 0.0):
    """Stage 2: approximate exclusive prefixes -> per-tile binade, or a flag
    ("dirty": the tile may contain / sit next to a binade crossing).
    ``tol`` bounds |sequential - approximate| / value for up to ``n_total`` adds.
    The sequential sum is monotone and never below the largest term seen, so
    the running maximum of the weights (``tile_max``, ``max_in`` from earlier
    ranks) tightens the lower bound - this keeps collapsed (one-hot-like)
    weights, whose sum sits exactly on a power of two, from flagging every
    tile."""
    tol = (float(n_total) + tile) * 2.0**-52
    incl = offset + np.cumsum(tile_sums)
    excl = np.concatenate(([offset], incl[:-1]))
    lo = excl * (1.0 - tol)
    if tile_max is not None:
        runmax = np.maximum.accumulate(np.concatenate(([max_in], tile_max)))[:-1]
        lo = np.maximum(lo, runmax)
    hi = incl * (1.0 + tol)
    e_lo = np.array([binade(v) for v in lo])
    e_hi = np.array([binade(v) for v in hi])
    clean = e_lo == e_hi
    return np.where(clean, e_lo, E_MIN - 1), excl


def exact_cumsum(w: np.ndarray, s_in: float = 0.0, n_total: int | None = None,
                 approx_offset: float | None = None, tile: int = TILE, stats: dict | None = None,
                 max_in: float = 0.0):
    """Staged algorithm: tile sums -> plan -> aggregates -> resolve -> write.
    ``s_in`` is the exact incoming running sum (previous rank), ``approx_offset``
    the approximate one the plan may use before ``s_in`` is known."""
    w = np.asarray(w, dtype=np.float64)
    n = len(w)
    n_total = n if n_total is None else n_total
    approx_offset = s_in if approx_offset is None else approx_offset
    ntiles = (n + tile - 1) // tile
    pad = ntiles * tile - n
    wp = np.concatenate([w, np.zeros(pad)]) if pad else w
    tiles = wp.reshape(ntiles, tile)
    tile_sums = tiles.sum(axis=1)                                  # stage 1
    e_tile, _ = plan(tile_sums, n_total, approx_offset, tile,      # stage 2
                     tile_max=tiles.max(axis=1), max_in=max_in)
    out = np.empty(ntiles * tile, dtype=np.float64)
    s = float(s_in)
    n_dirty = 0
    n_phases = 0
    for i in range(ntiles):                                        # stages 3-5
        e = int(e_tile[i])
        if e >= E_MIN:
            if binade(s) != e:
                raise AssertionError("plan tolerance violated")
            d0, d1 = element_maps(tiles[i], e)
            f0, f1 = pair_scan(d0, d1)
            m0 = np.int64(np.ldexp(s, 52 - e))
            m = m0 + (f0 if (m0 & 1) == 0 else f1)
            if m[-1] >= TWO53:
                raise AssertionError("clean tile crossed a binade")
            out[i * tile:(i + 1) * tile] = np.ldexp(m.astype(np.float64), e - 52)
            s = float(out[(i + 1) * tile - 1])
        else:
            n_dirty += 1
            o, s, ph = scan_tile(tiles[i], s)
            n_phases += ph
            out[i * tile:(i + 1) * tile] = o
    if stats is not None:
        stats.update(n_tiles=ntiles, n_dirty=n_dirty, n_phases=n_phases)
    return out[:n], s
