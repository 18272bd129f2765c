"""Tensor-level wrappers over the C-ABI (``include/aspire_b200.h``).

PyTorch is only the carrier: every op takes contiguous fp64 / int64 CUDA tensors,
passes ``data_ptr()`` + the current CUDA stream to the native library and returns
tensors.  Nothing here computes on the CPU and nothing falls back: without the
library or without an sm_100 device the calls raise ``AspireB200Error``.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import AspireB200Error, check

_ws_cache: dict = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise AspireB200Error(f"{name} must be a CUDA tensor (aspire_b200 has no CPU path)")
    if t.dtype != torch.float64:
        raise AspireB200Error(f"{name} must be float64, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.int64:
        raise AspireB200Error(f"{name} must be an int64 CUDA tensor")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def reduce_ws(device) -> torch.Tensor:
    """Zero-initialised scratch of the grid reductions (ticket + partials),
    one per (device, stream)."""
    key = ("reduce", torch.device(device), _stream())
    ws = _ws_cache.get(key)
    if ws is None:
        ws = torch.zeros(_lib.load().asp_reduce_workspace_bytes(), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _sized_ws(kind: str, nbytes: int, device) -> torch.Tensor:
    key = (kind, torch.device(device), _stream())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def device_check() -> None:
    check(_lib.load().asp_device_check(), "asp_device_check")


# ---- K1 -------------------------------------------------------------------
def logdensity(density, x: torch.Tensor) -> torch.Tensor:
    x = _f64(x, "x")
    if x.ndim != 2 or x.shape[1] != density.dims:
        raise AspireB200Error(f"x must be [n, {density.dims}]")
    out = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
    if x.shape[0] == 0:
        return out
    check(_lib.load().asp_logdensity(C.byref(density.struct()), x.data_ptr(), x.shape[0], out.data_ptr(), _stream()),
          "asp_logdensity")
    return out


# ---- K11 ------------------------------------------------------------------
def importance_weights(log_l, log_p, log_q):
    """-> (log_w [n], weights [n], result [8] device tensor); see the header for
    the layout of ``result``."""
    log_l, log_p, log_q = _f64(log_l, "log_l"), _f64(log_p, "log_p"), _f64(log_q, "log_q")
    n = log_l.numel()
    log_w = torch.empty_like(log_l)
    weights = torch.empty_like(log_l)
    result = torch.zeros(8, dtype=torch.float64, device=log_l.device)
    check(_lib.load().asp_importance_weights(log_l.data_ptr(), log_p.data_ptr(), log_q.data_ptr(), n,
                                             log_w.data_ptr(), weights.data_ptr(), result.data_ptr(),
                                             reduce_ws(log_l.device).data_ptr(), _stream()),
          "asp_importance_weights")
    return log_w, weights, result


# ---- K2 / K3 --------------------------------------------------------------
def reweight(log_l, log_p, log_q):
    """-> (d [n], stats [4] = {min d, max d, #NaN, #+inf}) on the device."""
    log_l, log_p, log_q = _f64(log_l, "log_l"), _f64(log_p, "log_p"), _f64(log_q, "log_q")
    n = log_l.numel()
    d = torch.empty_like(log_l)
    stats = torch.empty(4, dtype=torch.float64, device=log_l.device)
    check(_lib.load().asp_reweight(log_l.data_ptr(), log_p.data_ptr(), log_q.data_ptr(), n, d.data_ptr(),
                                   stats.data_ptr(), reduce_ws(d.device).data_ptr(), _stream()), "asp_reweight")
    return d, stats


def ess_sums(d, dbetas, stats) -> torch.Tensor:
    """-> [K, 2] device tensor of (S1, S2) per candidate ``beta - beta0``."""
    d, stats = _f64(d, "d"), _f64(stats, "stats")
    dbetas = torch.as_tensor(dbetas, dtype=torch.float64, device=d.device).reshape(-1).contiguous()
    k = dbetas.numel()
    sums = torch.empty(k, 2, dtype=torch.float64, device=d.device)
    check(_lib.load().asp_ess_sums(d.data_ptr(), d.numel(), dbetas.data_ptr(), k, stats.data_ptr(), sums.data_ptr(),
                                   reduce_ws(d.device).data_ptr(), _stream()), "asp_ess_sums")
    return sums


_STATE_WORDS = C.sizeof(_lib.BetaState) // 8


def beta_search(d, stats, beta_prev: float, target_eff: float, tol: float, n_total: int | None = None,
                allreduce=None, multi_level: bool = True, peer=None, root_guided: bool | None = None) -> _lib.BetaState:
    """Device-driven bisection of ``determine_beta`` (smc/base.py:166-186).

    ``multi_level`` (default) resolves three bisection levels per pass over ``d`` (``asp_beta_search_pass3``),
    with the same decisions and the same beta as one level per pass.
    Launches a fixed number of passes with no host round trip; ``allreduce``
    (multi-GPU) is called on the two-element sums tensor between a pass and its
    decision.  ``peer`` (a ``_lib.PeerDesc``, multi-GPU) replaces that with the all-reduce done inside the pass
    kernel over NVLink peer memory (``asp_beta_search_pass_p2p``): one launch per pass and no NCCL call.
    ``root_guided`` (default: whenever the sums need no host-side all-reduce) estimates the root of ESS/N = target
    first and evaluates only the tree node of the reference's bisection that contains it
    (``asp_beta_search_root``): same state, ~2.5x fewer passes over ``d``; the three-level passes launched after
    it are no-ops unless the node was missed three times.
    Returns the final state (one device->host read)."""
    lib = _lib.load()
    d, stats = _f64(d, "d"), _f64(stats, "stats")
    if root_guided is None:
        root_guided = multi_level and (allreduce is None or peer is not None)
    n = d.numel()
    n_total = float(n if n_total is None else n_total)
    words = lib.asp_beta_search_scratch_words()
    buf = torch.zeros(_STATE_WORDS + words + 4, dtype=torch.float64, device=d.device)
    state, scratch = buf[:_STATE_WORDS], buf[_STATE_WORDS:_STATE_WORDS + words]
    buf[_STATE_WORDS + words:] = stats  # read back with the state: min / max / #NaN / #inf of d
    ws = reduce_ws(d.device).data_ptr()
    st = _stream()
    sums = torch.zeros(14, dtype=torch.float64, device=d.device)
    peer_ref = None if peer is None else C.byref(peer)

    def read_state():
        host = buf.cpu().numpy()  # the search's one device -> host read
        out = _lib.BetaState.from_buffer_copy(host[:_STATE_WORDS].tobytes())
        out.n_nan, out.n_inf = float(host[-2]), float(host[-1])
        if out.reserved == -1:
            raise AspireB200Error("beta bisection: a peer GPU did not reach the in-kernel all-reduce within 20 s")
        return out

    def single_pass():
        if peer is not None:
            check(lib.asp_beta_search_pass_p2p(d.data_ptr(), n, state.data_ptr(), 0, None, peer_ref, ws, st),
                  "asp_beta_search_pass_p2p")
        elif allreduce is None:
            check(lib.asp_beta_search_pass(d.data_ptr(), n, state.data_ptr(), 1, ws, st), "asp_beta_search_pass")
        else:
            check(lib.asp_beta_search_pass(d.data_ptr(), n, state.data_ptr(), 0, ws, st), "asp_beta_search_pass")
            allreduce(state[7:9])  # s1, s2
            check(lib.asp_beta_search_decide(state.data_ptr(), st), "asp_beta_search_decide")

    def three_level_passes():
        for _ in range(lib.asp_beta_search_max_passes3(beta_prev, tol)):
            if peer is not None:
                check(lib.asp_beta_search_pass_p2p(d.data_ptr(), n, state.data_ptr(), 1, sums.data_ptr(), peer_ref, ws,
                                                   st), "asp_beta_search_pass_p2p")
            elif allreduce is None:
                check(lib.asp_beta_search_pass3(d.data_ptr(), n, state.data_ptr(), sums.data_ptr(), 1, ws, st),
                      "asp_beta_search_pass3")
            else:
                check(lib.asp_beta_search_pass3(d.data_ptr(), n, state.data_ptr(), sums.data_ptr(), 0, ws, st),
                      "asp_beta_search_pass3")
                allreduce(sums)
                check(lib.asp_beta_search_decide3(state.data_ptr(), sums.data_ptr(), st), "asp_beta_search_decide3")

    if root_guided:
        # init, the beta = 1 pass and the root-guided search in one call; the plain passes only after three misses
        check(lib.asp_beta_search_run(d.data_ptr(), n, stats.data_ptr(), beta_prev, target_eff, tol, n_total,
                                      state.data_ptr(), scratch.data_ptr(), peer_ref, ws, st), "asp_beta_search_run")
        out = read_state()
        if out.phase == 2:
            return out
        three_level_passes()
    else:
        check(lib.asp_beta_search_init(state.data_ptr(), beta_prev, target_eff, tol, n_total, stats.data_ptr(), st),
              "asp_beta_search_init")
        if not multi_level:
            for _ in range(lib.asp_beta_search_max_passes(beta_prev, tol)):
                single_pass()
        else:
            single_pass()  # ESS at beta = 1 (smc/base.py:170-175), then three bisection levels per pass
            three_level_passes()
    out = read_state()
    if out.phase != 2:
        raise AspireB200Error("beta bisection did not converge within the planned number of passes")
    return out


def smc_weights(d, dbeta: float, c: float, s1: float, n_total: int | None = None, want_p: bool = True,
                want_tiles: bool = True):
    """-> (p or None, result [4] device, tile_stats or None).  ``p`` is the
    normalised resampling weight ``exp(dbeta d - c) / s1``."""
    d = _f64(d, "d")
    n = d.numel()
    n_total = float(n if n_total is None else n_total)
    p = torch.empty_like(d) if want_p else None
    n_tiles = (n + _lib.CDF_TILE - 1) // _lib.CDF_TILE
    tiles = torch.empty(2 * n_tiles, dtype=torch.float64, device=d.device) if want_tiles else None
    result = torch.empty(4, dtype=torch.float64, device=d.device)
    check(_lib.load().asp_smc_weights(d.data_ptr(), n, dbeta, c, s1, n_total, _ptr(p), result.data_ptr(), _ptr(tiles),
                                      reduce_ws(d.device).data_ptr(), _stream()), "asp_smc_weights")
    return p, result, tiles


# ---- K5 -------------------------------------------------------------------
def cdf_workspace(n: int, device) -> torch.Tensor:
    return _sized_ws("cdf", _lib.load().asp_cdf_workspace_bytes(n), device)


def cdf_exact(p, normalise: bool = True, tile_stats=None, return_info: bool = False):
    """Sequential-order-exact ``cumsum(p)`` (then ``/= cdf[-1]``): bit-identical
    to what ``numpy.random.Generator.choice`` builds from ``p``."""
    lib = _lib.load()
    p = _f64(p, "p")
    n = p.numel()
    cdf = torch.empty_like(p)
    ws = cdf_workspace(n, p.device)
    st = _stream()
    if tile_stats is None:
        check(lib.asp_cdf_exact(p.data_ptr(), n, int(normalise), cdf.data_ptr(), ws.data_ptr(), st), "asp_cdf_exact")
    else:
        tile_stats = _f64(tile_stats, "tile_stats")
        check(lib.asp_cdf_set_tile_stats(tile_stats.data_ptr(), n, ws.data_ptr(), st), "asp_cdf_set_tile_stats")
        check(lib.asp_cdf_plan(n, None, ws.data_ptr(), st), "asp_cdf_plan")
        check(lib.asp_cdf_aggregate(p.data_ptr(), n, ws.data_ptr(), st), "asp_cdf_aggregate")
        check(lib.asp_cdf_resolve(p.data_ptr(), n, None, cdf.data_ptr(), None, ws.data_ptr(), st), "asp_cdf_resolve")
        check(lib.asp_cdf_write(p.data_ptr(), n, None, int(normalise), cdf.data_ptr(), ws.data_ptr(), st),
              "asp_cdf_write")
    if return_info:
        return cdf, cdf_info(ws)
    return cdf


# staged form (sharded CDF: the ranks interleave their collectives between the stages)
_cdf_ctx: dict = {}


def cdf_plan_aggregate(p, tile_stats, carry) -> None:
    """Stages 1-3 on this rank's slice.  ``carry`` = device [approx offset, running max, exact
    incoming sum (filled in later), n_total]."""
    lib = _lib.load()
    p, carry = _f64(p, "p"), _f64(carry, "carry")
    n = p.numel()
    ws = cdf_workspace(n, p.device)
    st = _stream()
    if tile_stats is None:
        check(lib.asp_cdf_tile_stats(p.data_ptr(), n, ws.data_ptr(), st), "asp_cdf_tile_stats")
    else:
        check(lib.asp_cdf_set_tile_stats(_f64(tile_stats, "tile_stats").data_ptr(), n, ws.data_ptr(), st),
              "asp_cdf_set_tile_stats")
    check(lib.asp_cdf_plan(n, carry.data_ptr(), ws.data_ptr(), st), "asp_cdf_plan")
    check(lib.asp_cdf_aggregate(p.data_ptr(), n, ws.data_ptr(), st), "asp_cdf_aggregate")
    _cdf_ctx[p.data_ptr()] = (ws, torch.empty_like(p))


def cdf_resolve(p, carry) -> torch.Tensor:
    """Stage 4 with the exact incoming sum in ``carry[2]``.  -> device [exact outgoing sum, running max]."""
    ws, cdf = _cdf_ctx[p.data_ptr()]
    out = torch.empty(2, dtype=torch.float64, device=p.device)
    check(_lib.load().asp_cdf_resolve(p.data_ptr(), p.numel(), carry.data_ptr(), cdf.data_ptr(), out.data_ptr(),
                                      ws.data_ptr(), _stream()), "asp_cdf_resolve")
    return out


def cdf_write(p, total, normalise: bool = True) -> torch.Tensor:
    """Stage 5: this rank's slice of the global CDF, normalised by the device scalar ``total``."""
    ws, cdf = _cdf_ctx.pop(p.data_ptr())
    check(_lib.load().asp_cdf_write(p.data_ptr(), p.numel(), _f64(total, "total").data_ptr(), int(normalise),
                                    cdf.data_ptr(), ws.data_ptr(), _stream()), "asp_cdf_write")
    return cdf


def cdf_resolve_write_p2p(p, peer, normalise: bool = True):
    """Stages 4 + 5 with the carry chain and the total crossing the ranks through NVLink peer memory
    (``asp_cdf_resolve_p2p`` / ``asp_cdf_write_p2p``) instead of send / recv / broadcast.
    -> (cdf slice, device [running sum before this slice, exact total])."""
    ws, cdf = _cdf_ctx.pop(p.data_ptr())
    lib, ref = _lib.load(), C.byref(peer)
    chain = torch.empty(2, dtype=torch.float64, device=p.device)
    check(lib.asp_cdf_resolve_p2p(p.data_ptr(), p.numel(), cdf.data_ptr(), chain.data_ptr(), ref, ws.data_ptr(),
                                  _stream()), "asp_cdf_resolve_p2p")
    check(lib.asp_cdf_write_p2p(p.data_ptr(), p.numel(), int(normalise), cdf.data_ptr(), chain.data_ptr(), ref,
                                ws.data_ptr(), _stream()), "asp_cdf_write_p2p")
    return cdf, chain


def cdf_info(ws) -> dict:
    info = (C.c_int64 * 4)()
    check(_lib.load().asp_cdf_info(ws.data_ptr(), info, _stream()), "asp_cdf_info")
    return dict(n_tiles=info[0], n_dirty=info[1], n_phases=info[2], error_flags=info[3])


# ---- K6 / K7 --------------------------------------------------------------
def search_multinomial(cdf, u, guided: bool | None = None) -> torch.Tensor:
    """``cdf.searchsorted(u, side='right')``.  ``guided`` (default: when there are at
    least n/8 queries) goes through a guide table - same indices, ~10x less traffic."""
    cdf, u = _f64(cdf, "cdf"), _f64(u, "u")
    n, m = cdf.numel(), u.numel()
    idx = torch.empty(m, dtype=torch.int64, device=cdf.device)
    if m == 0:
        return idx
    lib = _lib.load()
    if guided is None:
        guided = m * 8 >= n and n >= 4096
    if guided:
        ws = _sized_ws("guide", lib.asp_search_workspace_bytes(n), cdf.device)
        check(lib.asp_search_multinomial_guided(cdf.data_ptr(), n, u.data_ptr(), m, idx.data_ptr(), ws.data_ptr(),
                                                _stream()), "asp_search_multinomial_guided")
    else:
        check(lib.asp_search_multinomial(cdf.data_ptr(), n, u.data_ptr(), m, idx.data_ptr(), _stream()),
              "asp_search_multinomial")
    return idx


SPACING_TILE = 2048


class SortedSpacings:
    """Integer exponential spacings of ``m`` sorted uniforms (``oracle/sorted_resample.py``): ``q`` covers the
    elements ``[j0, j0 + tiles * 2048)``, ``tile_prefix`` / ``total`` are global."""

    def __init__(self, seed: int, stream_id: int, m: int, device, j0: int = 0, n_tiles: int | None = None):
        lib = _lib.load()
        self.seed, self.stream_id, self.m, self.j0 = int(seed) & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFF, int(m), int(j0)
        self.tiles_total = int(lib.asp_sorted_tiles(self.m))
        self.n_tiles = self.tiles_total - self.j0 // SPACING_TILE if n_tiles is None else int(n_tiles)
        self.q = torch.empty(self.n_tiles * SPACING_TILE, dtype=torch.int32, device=device)
        self.tile_sums = torch.empty(self.n_tiles, dtype=torch.int64, device=device)
        if self.n_tiles > 0:
            check(lib.asp_sorted_spacings(self.seed, self.stream_id, self.j0, self.n_tiles, self.m, self.q.data_ptr(),
                                          self.tile_sums.data_ptr(), _stream()), "asp_sorted_spacings")
        self.tile_prefix = None
        self.total = None

    def scan(self, all_tile_sums: torch.Tensor | None = None):
        """Exclusive prefix over ALL tiles (``all_tile_sums``: the all-gathered sums when sharded)."""
        sums = self.tile_sums if all_tile_sums is None else all_tile_sums
        if sums.numel() != self.tiles_total:
            raise AspireB200Error("tile sums do not cover all tiles")
        self.tile_prefix = torch.empty(self.tiles_total, dtype=torch.int64, device=sums.device)
        self.total = torch.empty(1, dtype=torch.int64, device=sums.device)
        check(_lib.load().asp_sorted_tile_scan(sums.data_ptr(), self.tiles_total, self.tile_prefix.data_ptr(),
                                               self.total.data_ptr(), _stream()), "asp_sorted_tile_scan")
        return self

    def search(self, cdf, j_lo: int = 0, j_hi: int | None = None) -> torch.Tensor:
        """Ancestors (indices into ``cdf``) of the slots ``[j_lo, j_hi)``: sorted ascending."""
        cdf = _f64(cdf, "cdf")
        j_hi = self.m if j_hi is None else int(j_hi)
        idx = torch.empty(j_hi - j_lo, dtype=torch.int64, device=cdf.device)
        if j_hi > j_lo:
            lib = _lib.load()
            ws = _sized_ws("sorted", lib.asp_search_sorted_workspace_bytes(j_hi - j_lo), cdf.device)
            check(lib.asp_search_sorted(cdf.data_ptr(), cdf.numel(), self.q.data_ptr(), self.j0,
                                        self.tile_prefix.data_ptr(), self.total.data_ptr(), int(j_lo), j_hi,
                                        idx.data_ptr(), ws.data_ptr(), _stream()), "asp_search_sorted")
        return idx

    def lower_bound(self, thresholds: torch.Tensor) -> torch.Tensor:
        """``#{j < m : S_j < t}`` for each (int64 device) threshold; needs :meth:`scan` only."""
        out = torch.empty(thresholds.numel(), dtype=torch.int64, device=thresholds.device)
        check(_lib.load().asp_sorted_lower_bound(self.seed, self.stream_id, self.tile_prefix.data_ptr(),
                                                 self.tiles_total, self.m, thresholds.data_ptr(), thresholds.numel(),
                                                 out.data_ptr(), _stream()), "asp_sorted_lower_bound")
        return out

    def window(self, j_lo: int, j_hi: int) -> "SortedSpacings":
        """Spacings covering the slots ``[j_lo, j_hi)`` with this object's global tile table: itself when it
        already covers them, else a regenerated window (counter-based: no exchange)."""
        t0, t1 = j_lo // SPACING_TILE, max((j_hi - 1) // SPACING_TILE + 1, j_lo // SPACING_TILE + 1)
        if self.j0 <= t0 * SPACING_TILE and t1 * SPACING_TILE <= self.j0 + self.n_tiles * SPACING_TILE:
            return self
        win = SortedSpacings(self.seed, self.stream_id, self.m, self.q.device, j0=t0 * SPACING_TILE, n_tiles=t1 - t0)
        win.tile_prefix, win.total = self.tile_prefix, self.total
        return win


def sorted_tiles(m: int) -> int:
    return int(_lib.load().asp_sorted_tiles(int(m)))


def sorted_bounds(sp: SortedSpacings, lo_all: torch.Tensor, m: int) -> list:
    """Slot boundaries of the owner-computes partition: ``lo_all[s]`` is the global CDF value in front of rank s's
    slice; rank s owns the slots ``[J_s, J_{s+1})``, ``J_s = #{j : S_j < rint(lo_s total)}``, ``J_G = m``.
    -> host ints [J_0 .. J_G] (one device -> host read)."""
    thr_f = torch.round(lo_all.to(torch.float64) * sp.total.to(torch.float64))
    j = sp.lower_bound(thr_f.nan_to_num(0.0).to(torch.int64).contiguous())
    j = torch.where(torch.isfinite(thr_f), j, torch.full_like(j, -1)).cpu().tolist()   # one device -> host read
    if min(j) < 0:
        return None  # a non-finite CDF boundary: a peer never handed over its running sum
    j[0] = 0
    return [int(v) for v in j] + [int(m)]


def search_multinomial_sorted(cdf, seed: int, stream_id: int, m: int) -> torch.Tensor:
    """Multinomial resampling at ``m`` sorted uniforms generated on the device (throughput mode): the same
    ``searchsorted(cdf, u, 'right')`` as :func:`search_multinomial`, ancestors ascending."""
    cdf = _f64(cdf, "cdf")
    return SortedSpacings(seed, stream_id, m, cdf.device).scan().search(cdf)


def search_systematic(cdf, u0: float, m: int, j0: int = 0, m_total: int | None = None) -> torch.Tensor:
    cdf = _f64(cdf, "cdf")
    m_total = m if m_total is None else m_total
    idx = torch.empty(m, dtype=torch.int64, device=cdf.device)
    if m == 0:
        return idx
    lib = _lib.load()
    ws = _sized_ws("systematic", lib.asp_search_systematic_workspace_bytes(m), cdf.device)
    check(lib.asp_search_systematic(cdf.data_ptr(), cdf.numel(), float(u0), j0, m, m_total, idx.data_ptr(),
                                    ws.data_ptr(), _stream()), "asp_search_systematic")
    return idx


def gather_rows(x, idx, log_l=None, log_p=None, log_q=None):
    x, idx = _f64(x, "x"), _i64(idx, "idx")
    m, dims = idx.numel(), x.shape[1]
    xo = torch.empty(m, dims, dtype=torch.float64, device=x.device)
    logs = [None if t is None else _f64(t, "log") for t in (log_l, log_p, log_q)]
    outs = [None if t is None else torch.empty(m, dtype=torch.float64, device=x.device) for t in logs]
    if m:
        check(_lib.load().asp_gather_rows(x.data_ptr(), dims, idx.data_ptr(), m, xo.data_ptr(), _ptr(logs[0]),
                                          _ptr(logs[1]), _ptr(logs[2]), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]),
                                          _stream()), "asp_gather_rows")
    return (xo, *outs)


# ---- K8 -------------------------------------------------------------------
def weighted_moments(x, d=None, dbeta: float = 0.0, c: float = 0.0, idx=None, stride: int = 1, first: int = 0) -> torch.Tensor:
    """-> device tensor [1 + 2 D] = {sum w, sum w x_j, sum w x_j^2} over rows ``idx`` (all rows
    when None), w = exp(dbeta d - c) (w = 1 when ``d`` is None).  ``stride`` / ``first``: only the rows
    ``idx[first::stride]`` (``range(first, n, stride)`` without ``idx``), read in place - no strided copy of the index."""
    x = _f64(x, "x")
    dims = x.shape[1]
    idx = None if idx is None else _i64(idx, "idx")
    n_rows = x.shape[0] if idx is None else idx.numel()
    stride, first = int(stride), int(first)
    n = max((n_rows - first + stride - 1) // stride, 0)
    d = None if d is None else _f64(d, "d")
    out = torch.empty(1 + 2 * dims, dtype=torch.float64, device=x.device)
    if n == 0:
        return out.zero_()
    if idx is None and first:
        raise AspireB200Error("weighted_moments: `first` needs an index array")
    ws = _sized_ws("moments", _lib.load().asp_moments_workspace_bytes(dims), x.device)
    idx_ptr = None if idx is None else idx.data_ptr() + 8 * first
    check(_lib.load().asp_weighted_moments(x.data_ptr(), dims, idx_ptr, stride, _ptr(d), n, dbeta, c, out.data_ptr(),
                                           ws.data_ptr(), _stream()), "asp_weighted_moments")
    return out


def moments_to_mu_sigma(m: torch.Tensor, dims: int, floor: float = 0.0):
    """One-pass mean / std (at least ``floor``) from the moment sums: one launch, device tensors, no sync."""
    m = _f64(m, "moments")
    out = torch.empty(2, dims, dtype=torch.float64, device=m.device)
    check(_lib.load().asp_moments_finish(m.data_ptr(), dims, float(floor), out[0].data_ptr(), out[1].data_ptr(),
                                         _stream()), "asp_moments_finish")
    return out[0], out[1]


def make_pcn_params(beta, rho, mu, sigma, n_moves, seed=0, stream0=0, index0=0, xi=None, log_u=None, slot=None,
                    total_accepts_only=False, nu: float = 0.0, gamma=None):
    mu, sigma = _f64(mu, "mu"), _f64(sigma, "sigma")
    prm = _lib.PcnParams()
    prm.beta, prm.rho = float(beta), float(rho)
    prm.d_mu, prm.d_sigma = mu.data_ptr(), sigma.data_ptr()
    prm.n_moves = int(n_moves)
    prm.noise_mode = 0 if xi is None else 1
    prm.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    prm.stream0 = int(stream0) & 0xFFFFFFFF
    prm.index0 = int(index0)
    prm.reserved = 1 if total_accepts_only else 0
    prm.d_xi = None if xi is None else _f64(xi, "xi").data_ptr()
    prm.d_log_u = None if log_u is None else _f64(log_u, "log_u").data_ptr()
    prm.d_slot = None if slot is None else _i64(slot, "slot").data_ptr()
    prm.nu = float(nu)  # > 0: tpCN (Student-t reference measure)
    prm.d_scale = None
    prm.d_gamma = None if gamma is None else _f64(gamma, "gamma").data_ptr()
    prm._keep = (mu, sigma, xi, log_u, slot, gamma)  # keep the tensors alive
    return prm


def pcn_sweep(x, log_l, log_p, log_q, idx, like, prior, q, prm, out=None):
    """Fused gather + ``n_moves`` pCN moves.  -> (x, log_l, log_p, log_q, accepts [n_moves] int64)."""
    x, log_l, log_p, log_q = _f64(x, "x"), _f64(log_l, "log_l"), _f64(log_p, "log_p"), _f64(log_q, "log_q")
    dims = x.shape[1]
    n = x.shape[0] if idx is None else idx.numel()
    idx = None if idx is None else _i64(idx, "idx")
    if out is None:
        out = (torch.empty(n, dims, dtype=torch.float64, device=x.device),
               *(torch.empty(n, dtype=torch.float64, device=x.device) for _ in range(3)))
    accepts = torch.zeros(prm.n_moves + 2, dtype=torch.int64, device=x.device)
    prm._nan = accepts[1:2] if prm.reserved else accepts[prm.n_moves:prm.n_moves + 1]  # final log q that are NaN
    check(_lib.load().asp_pcn_sweep(x.data_ptr(), log_l.data_ptr(), log_p.data_ptr(), log_q.data_ptr(), _ptr(idx), n,
                                    dims, C.byref(like.struct()), C.byref(prior.struct()), C.byref(q.struct()),
                                    C.byref(prm), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                    out[3].data_ptr(), accepts.data_ptr(), _stream()), "asp_pcn_sweep")
    return (*out, accepts[:1] if prm.reserved else accepts[:prm.n_moves])


def pcn_sweep_tpcn_supported(dims: int) -> bool:
    """Whether ``pcn_sweep`` runs tpCN moves (``prm.nu > 0``) in the fused kernel for this dimension."""
    return bool(_lib.load().asp_pcn_sweep_tpcn_supported(int(dims)))


def student_t_fit(x, mu0, sigma0, idx=None, stride: int = 1, m: int | None = None, n_iter: int = 30, nu0: float = 10.0,
                  nu_min: float = 1.0, nu_max: float = 1e6):
    """EM fit of t_nu(mu, diag sigma^2) (tpCN reference measure; spec ``oracle/pcn.py::fit_student_t_diag``) to the
    rows ``x[idx[i * stride]]`` (``idx`` None: ``x[i * stride]``), ``i < m``, on the device.  -> (mu, sigma, nu),
    ``nu`` read back to the host (the one device -> host read of the fit)."""
    x = _f64(x, "x")
    dims = x.shape[1]
    idx = None if idx is None else _i64(idx, "idx")
    n_rows = x.shape[0] if idx is None else idx.numel()
    if m is None:
        m = (n_rows + stride - 1) // stride
    if m < 1 or (m - 1) * stride >= n_rows:
        raise AspireB200Error("student_t_fit: row selection out of range")
    state = torch.cat([_f64(mu0, "mu0").reshape(-1), _f64(sigma0, "sigma0").reshape(-1),
                       torch.tensor([float(nu0)], dtype=torch.float64, device=x.device)]).contiguous()
    lib = _lib.load()
    ws = _sized_ws("student_t", lib.asp_student_t_workspace_bytes(dims), x.device)
    check(lib.asp_student_t_fit(x.data_ptr(), dims, _ptr(idx), int(stride), int(m), int(n_iter), float(nu_min),
                                float(nu_max), state.data_ptr(), ws.data_ptr(), _stream()), "asp_student_t_fit")
    return state[:dims].contiguous(), state[dims:2 * dims].contiguous(), float(state[2 * dims])


def pcn_sweep_flow_supported(dims: int, like, prior) -> bool:
    return bool(_lib.load().asp_pcn_sweep_flow_supported(int(dims), int(like.kind), int(prior.kind)))


def pcn_sweep_flow(x, log_l, log_p, log_q, idx, like, prior, flow_struct, prm, out=None):
    """Fused gather + ``n_moves`` pCN moves whose proposal density is a coupling flow (log q(x') on tcgen05 inside
    the move).  -> (x, log_l, log_p, log_q, accepts)."""
    x, log_l, log_p, log_q = _f64(x, "x"), _f64(log_l, "log_l"), _f64(log_p, "log_p"), _f64(log_q, "log_q")
    dims = x.shape[1]
    n = x.shape[0] if idx is None else idx.numel()
    idx = None if idx is None else _i64(idx, "idx")
    if out is None:
        out = (torch.empty(n, dims, dtype=torch.float64, device=x.device),
               *(torch.empty(n, dtype=torch.float64, device=x.device) for _ in range(3)))
    lib = _lib.load()
    ws = _sized_ws("flow_sweep", lib.asp_pcn_sweep_flow_workspace_bytes(dims), x.device)
    accepts = torch.zeros(prm.n_moves + 2, dtype=torch.int64, device=x.device)
    prm._nan = accepts[1:2] if prm.reserved else accepts[prm.n_moves:prm.n_moves + 1]  # final log q that are NaN
    check(lib.asp_pcn_sweep_flow(x.data_ptr(), log_l.data_ptr(), log_p.data_ptr(), log_q.data_ptr(), _ptr(idx), n, dims,
                                 C.byref(like.struct()), C.byref(prior.struct()), C.byref(flow_struct), C.byref(prm),
                                 out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                 accepts.data_ptr(), ws.data_ptr(), _stream()), "asp_pcn_sweep_flow")
    return (*out, accepts[:1] if prm.reserved else accepts[:prm.n_moves])


def tpcn_scales(x, prm, move: int) -> torch.Tensor:
    """tpCN: per-particle proposal scale sqrt(s_i) of move ``move`` (``oracle/pcn.py::tpcn_scale``); the
    result is attached to ``prm`` (``d_scale``) for the following ``pcn_propose``."""
    x = _f64(x, "x")
    sc = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
    check(_lib.load().asp_tpcn_scales(x.data_ptr(), x.shape[0], x.shape[1], C.byref(prm), move, sc.data_ptr(), _stream()),
          "asp_tpcn_scales")
    prm.d_scale = sc.data_ptr()
    prm._scale = sc  # keep alive
    return sc


def pcn_propose(x, prm, move: int) -> torch.Tensor:
    x = _f64(x, "x")
    xp = torch.empty_like(x)
    check(_lib.load().asp_pcn_propose(x.data_ptr(), x.shape[0], x.shape[1], C.byref(prm), move, xp.data_ptr(),
                                      _stream()), "asp_pcn_propose")
    return xp


def pcn_accept(x, log_l, log_p, log_q, x_prop, log_l_prop, log_p_prop, log_q_prop, prm, move: int, accepts):
    """In-place accept/reject of move ``move``; ``accepts`` is an int64 [n_moves] counter tensor."""
    ts = [_f64(t, "tensor") for t in (x, log_l, log_p, log_q, x_prop, log_l_prop, log_p_prop, log_q_prop)]
    for a, b in zip(ts[:4], (x, log_l, log_p, log_q)):
        if a.data_ptr() != b.data_ptr():
            raise AspireB200Error("pcn_accept updates in place: state tensors must be contiguous")
    check(_lib.load().asp_pcn_accept(*(t.data_ptr() for t in ts), x.shape[0], x.shape[1], C.byref(prm), move,
                                     accepts.data_ptr(), _stream()), "asp_pcn_accept")


def pcn_propose_eval_supported(dims: int, like, prior) -> bool:
    return bool(_lib.load().asp_pcn_propose_eval_supported(int(dims), int(like.kind), int(prior.kind)))


def pcn_propose_eval(x, like, prior, prm, move: int):
    """One pass: proposal x' of move ``move`` + built-in log prior / log likelihood at x' + r(x') - r(x).
    -> (x', log_l', log_p', dr)"""
    x = _f64(x, "x")
    n, dims = x.shape
    xp = torch.empty_like(x)
    ll2, lp2, dr = (torch.empty(n, dtype=torch.float64, device=x.device) for _ in range(3))
    check(_lib.load().asp_pcn_propose_eval(x.data_ptr(), n, dims, C.byref(like.struct()), C.byref(prior.struct()),
                                           C.byref(prm), move, xp.data_ptr(), ll2.data_ptr(), lp2.data_ptr(),
                                           dr.data_ptr(), _stream()), "asp_pcn_propose_eval")
    return xp, ll2, lp2, dr


def pcn_accept_scalar(x, log_l, log_p, log_q, x_prop, log_l_prop, log_p_prop, log_q_prop, dr, prm, move: int, accepts):
    """In-place accept/reject of move ``move`` from the scalars of ``pcn_propose_eval`` + the caller's log q(x')."""
    ts = [_f64(t, "tensor") for t in (x, log_l, log_p, log_q, x_prop, log_l_prop, log_p_prop, log_q_prop, dr)]
    for a, b in zip(ts[:4], (x, log_l, log_p, log_q)):
        if a.data_ptr() != b.data_ptr():
            raise AspireB200Error("pcn_accept_scalar updates in place: state tensors must be contiguous")
    check(_lib.load().asp_pcn_accept_scalar(*(t.data_ptr() for t in ts), x.shape[0], x.shape[1], C.byref(prm), move,
                                            accepts.data_ptr(), _stream()), "asp_pcn_accept_scalar")


# ---- K9 / K10 ---------------------------------------------------------------
def flow_logprob(flow_struct, x, want_z: bool = False):
    """Coupling-flow log_prob on tcgen05 (bf16 GEMMs, fp32 epilogue).  -> log_q [n] (, z [n, D])"""
    x = _f64(x, "x")
    n = x.shape[0]
    logq = torch.empty(n, dtype=torch.float64, device=x.device)
    z = torch.empty_like(x) if want_z else None
    if n:
        check(_lib.load().asp_flow_logprob_bf16(C.byref(flow_struct), x.data_ptr(), n, logq.data_ptr(), _ptr(z), _stream()),
              "asp_flow_logprob_bf16")
    return (logq, z) if want_z else logq


def flow_sample(flow_struct, u):
    """Base normals u [n, D] -> (x [n, D], log_q [n])."""
    u = _f64(u, "u")
    n = u.shape[0]
    x = torch.empty_like(u)
    logq = torch.empty(n, dtype=torch.float64, device=u.device)
    if n:
        check(_lib.load().asp_flow_sample_bf16(C.byref(flow_struct), u.data_ptr(), n, x.data_ptr(), logq.data_ptr(),
                                               _stream()), "asp_flow_sample_bf16")
    return x, logq


# ---- K12 ------------------------------------------------------------------
def select_uniforms(n_total: int, lo: float, hi: float, capacity: int, device, seed: int = 0, stream_id: int = 0,
                    u_all=None):
    """Owner-computes resampling: the (slot, u) pairs of the global uniform stream with
    lo <= u < hi, in slot order.  -> (slot [k] int64, u [k]) with k read back from the device."""
    lib = _lib.load()
    slot = torch.empty(capacity, dtype=torch.int64, device=device)
    u = torch.empty(capacity, dtype=torch.float64, device=device)
    count = torch.zeros(1, dtype=torch.int64, device=device)
    ws = _sized_ws("select", lib.asp_select_workspace_bytes(), device)
    u_all = None if u_all is None else _f64(u_all, "u_all")
    check(lib.asp_select_uniforms(int(seed) & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFF, _ptr(u_all), n_total,
                                  float(lo), float(hi), capacity, slot.data_ptr(), u.data_ptr(), count.data_ptr(),
                                  ws.data_ptr(), _stream()), "asp_select_uniforms")
    k = int(count.item())
    if k > capacity:
        raise AspireB200Error(f"select_uniforms: {k} offspring exceed the shard capacity {capacity}")
    return slot[:k], u[:k]


# ---- K13 ------------------------------------------------------------------
def philox_uniform(seed: int, stream_id: int, n: int, device, index0: int = 0) -> torch.Tensor:
    out = torch.empty(n, dtype=torch.float64, device=device)
    if n:
        check(_lib.load().asp_philox_uniform(int(seed) & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFF, index0, n,
                                             out.data_ptr(), _stream()), "asp_philox_uniform")
    return out


def philox_normal(seed: int, stream_id: int, n: int, dims: int, device, mean=None, sigma=None,
                  index0: int = 0) -> torch.Tensor:
    out = torch.empty(n, dims, dtype=torch.float64, device=device)
    if n:
        check(_lib.load().asp_philox_normal(int(seed) & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFF, index0, n, dims,
                                            _ptr(mean), _ptr(sigma), out.data_ptr(), _stream()), "asp_philox_normal")
    return out


def pcg64_uniform(state: int, inc: int, n: int, device, offset: int = 0) -> torch.Tensor:
    """Draws ``offset .. offset+n-1`` of ``numpy.random.Generator.random`` for a PCG64 generator with the
    given 128-bit ``(state, inc)`` (``rng.bit_generator.state['state']``), generated on the device."""
    out = torch.empty(n, dtype=torch.float64, device=device)
    m64 = (1 << 64) - 1
    if n:
        check(_lib.load().asp_pcg64_uniform((state >> 64) & m64, state & m64, (inc >> 64) & m64, inc & m64, int(offset),
                                            n, out.data_ptr(), _stream()), "asp_pcg64_uniform")
    return out


# ---- preconditioning transforms (scope row 8f-4) ----------------------------
def transform_apply(x, kind, lower, upper, bounded: int, eps: float, mean=None, std=None, inverse: bool = False,
                    want_log_j: bool = True):
    """``CompositeTransform.forward`` / ``inverse`` (transforms.py:271-316) in one pass over the rows of ``x``.
    ``kind`` int32 [D] (0 untouched, 1 periodic, 2 bounded) or None; ``mean`` / ``std`` select the affine stage.
    -> (transformed rows [N, D], log |det J| [N] or None)"""
    x = _f64(x, "x")
    if x.dim() != 2:
        raise AspireB200Error("x must be [N, D]")
    n, dims = x.shape
    t = _lib.TransformDesc()
    t.dims, t.bounded, t.affine, t.eps = dims, int(bounded), int(mean is not None), float(eps)
    keep = []
    if kind is not None:
        if kind.dtype != torch.int32 or not kind.is_cuda or kind.numel() != dims:
            raise AspireB200Error("kind must be a CUDA int32 tensor of length D")
        lower, upper = _f64(lower, "lower"), _f64(upper, "upper")
        if lower.numel() != dims or upper.numel() != dims:
            raise AspireB200Error("lower / upper must have length D")
        keep += [kind, lower, upper]
        t.d_kind, t.d_lower, t.d_upper = kind.data_ptr(), lower.data_ptr(), upper.data_ptr()
    if mean is not None:
        mean, std = _f64(mean, "mean"), _f64(std, "std")
        if mean.numel() != dims or std.numel() != dims:
            raise AspireB200Error("mean / std must have length D")
        keep += [mean, std]
        t.d_mean, t.d_std = mean.data_ptr(), std.data_ptr()
    out = torch.empty_like(x)
    log_j = torch.empty(n, dtype=torch.float64, device=x.device) if want_log_j else None
    if dims < 1 or dims > 256:
        raise AspireB200Error("asp_transform: dims must be in [1, 256]")
    if int(bounded) not in (_lib.BOUNDED_NONE, _lib.BOUNDED_LOGIT, _lib.BOUNDED_PROBIT):
        raise AspireB200Error("asp_transform: unknown bounded transform")
    if n == 0:  # an empty tensor has no storage to point at
        return out, log_j
    fn = _lib.load().asp_transform_inverse if inverse else _lib.load().asp_transform_forward
    check(fn(C.byref(t), x.data_ptr(), n, out.data_ptr(), _ptr(log_j), _stream()),
          "asp_transform_inverse" if inverse else "asp_transform_forward")
    return out, log_j


# ---- device guard ------------------------------------------------------------------------------------------
# Every C entry point launches on the CURRENT device with the current stream of that device.  The wrappers
# therefore run on the device of their tensors: when that is not the current device (``Aspire(device="cuda:1")``
# without a ``torch.cuda.set_device``), the call is made under ``torch.cuda.device(...)``, so the stream, the
# workspaces and the launch all belong to the tensors' GPU; tensors on different GPUs are refused.
def _tensor_device(args, kwargs):
    dev = None
    stack = list(args) + list(kwargs.values())
    while stack:
        a = stack.pop()
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                if dev is None:
                    dev = a.device
                elif a.device != dev:
                    raise AspireB200Error(f"tensors on different devices ({dev} and {a.device}) in one call")
        elif isinstance(a, (list, tuple)):
            stack.extend(a)
        elif isinstance(a, SortedSpacings):
            stack.append(a.__dict__.get("q"))
        elif isinstance(a, (str, torch.device)) and not isinstance(a, bool):
            try:
                d = torch.device(a)
            except (RuntimeError, TypeError):
                continue
            if d.type == "cuda" and d.index is not None and dev is None:
                dev = d
    return dev


def _on_tensor_device(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _tensor_device(args, kwargs)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def _install_device_guard():
    import types

    g = globals()
    for name, obj in list(g.items()):
        if name.startswith("_") or not isinstance(obj, types.FunctionType) or obj.__module__ != __name__:
            continue
        g[name] = _on_tensor_device(obj)
    for cls in (SortedSpacings,):
        for name, obj in list(vars(cls).items()):
            if isinstance(obj, types.FunctionType) and name in ("__init__", "scan", "search", "lower_bound", "window"):
                setattr(cls, name, _on_tensor_device(obj))


_install_device_guard()
