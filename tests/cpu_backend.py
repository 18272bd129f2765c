"""Oracle-backed stand-in for ``aspire_b200.ops`` on CPU tensors, used ONLY by the
world_size-2 gloo tests of the sharded orchestration (``aspire_b200.distributed``).
The product never imports this: without the CUDA library it raises."""
import types

import numpy as np
import torch

from oracle import densities as dn
from oracle import exact_scan as es
from oracle import pcn as opcn
from oracle import philox as oph

TILE = 2048


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a))


def reweight(ll, lp, lq):
    d = (_np(ll) + _np(lp)) - _np(lq)
    return _t(d), _t(np.array([d.min(), d.max(), float(np.isnan(d).sum()), float((d == np.inf).sum())]))


def ess_sums(d, dbetas, stats):
    d, st = _np(d), _np(stats)
    out = []
    for db in np.atleast_1d(_np(dbetas)):
        c = db * (st[1] if db >= 0 else st[0])
        e = np.exp(db * d - c)
        out.append([e.sum(), (e * e).sum()])
    return _t(np.array(out))


def beta_search(d, stats, beta_prev, target_eff, tol, n_total=None, allreduce=None):
    """The device-driven bisection of asp_beta_search_*, pass by pass."""
    dn_, st = _np(d), _np(stats)
    n_total = float(len(dn_) if n_total is None else n_total)
    s = types.SimpleNamespace(beta_min=beta_prev, beta_max=1.0, beta_try=1.0, eff_beta_max=0.0, s1_min=0.0,
                              s2_min=0.0, min_valid=0, n_evals=0, phase=0, d_min=float(st[0]), d_max=float(st[1]),
                              n_nan=float(st[2]), n_inf=float(st[3]))
    while s.phase != 2:
        db = s.beta_try - beta_prev
        c = db * (st[1] if db >= 0 else st[0])
        e = np.exp(db * dn_ - c)
        sums = _t(np.array([e.sum(), (e * e).sum()]))
        if allreduce is not None:
            allreduce(sums)
        s1, s2 = float(sums[0]), float(sums[1])
        eff = (s1 * s1 / s2) / n_total
        if s.phase == 0:
            s.eff_beta_max = eff
            if eff >= target_eff:
                s.beta_min, s.s1_min, s.s2_min, s.min_valid = 1.0, s1, s2, 1
            s.phase = 1
        elif eff >= target_eff:
            s.beta_min, s.s1_min, s.s2_min, s.min_valid = s.beta_try, s1, s2, 1
        else:
            s.beta_max = s.beta_try
        s.n_evals += 1
        if s.beta_max - s.beta_min > tol:
            s.beta_try = 0.5 * (s.beta_max + s.beta_min)
        else:
            s.phase = 2
    return s


def smc_weights(d, dbeta, c, s1, n_total=None, want_p=True, want_tiles=True):
    dn_ = _np(d)
    n_total = float(len(dn_) if n_total is None else n_total)
    u = np.exp(dbeta * dn_ - c)
    p = u / s1
    nt = (len(p) + TILE - 1) // TILE
    pp = np.concatenate([p, np.zeros(nt * TILE - len(p))]).reshape(nt, TILE)
    tiles = np.stack([pp.sum(1), pp.max(1)], axis=1).reshape(-1)
    res = np.array([((u - s1 / n_total) ** 2).sum(), p.sum(), 0.0, 0.0])
    return _t(p), _t(res), _t(tiles)


_ctx = {}


def cdf_plan_aggregate(p, tiles, carry):
    _ctx[id(p)] = None  # the oracle scan redoes the stages inside exact_cumsum


def cdf_resolve(p, carry):
    c = _np(carry)
    out, s = es.exact_cumsum(_np(p), s_in=float(c[2]), n_total=int(c[3]), approx_offset=float(c[0]), max_in=float(c[1]))
    _ctx[id(p)] = out
    return _t(np.array([s, max(float(c[1]), float(_np(p).max()))]))


def cdf_write(p, total, normalise=True):
    raw = _ctx.pop(id(p))
    return _t(raw / float(_np(total)[0]) if normalise else raw)


SPACING_TILE = 2048


def sorted_tiles(m):
    return (m + 1 + SPACING_TILE - 1) // SPACING_TILE


class SortedSpacings:
    """oracle/sorted_resample.py behind the interface of ``aspire_b200.ops.SortedSpacings``."""

    def __init__(self, seed, stream_id, m, device, j0=0, n_tiles=None):
        from oracle import sorted_resample as sr

        self.seed, self.stream_id, self.m, self.j0 = seed, stream_id, m, j0
        self.tiles_total = sorted_tiles(m)
        self.n_tiles = self.tiles_total - j0 // SPACING_TILE if n_tiles is None else n_tiles
        e = np.arange(j0, j0 + self.n_tiles * SPACING_TILE, dtype=np.uint64)
        q = sr.spacing_q(sr.uniform_k53(seed, stream_id, e)).astype(np.int64) if e.size else np.zeros(0, np.int64)
        q[e > m] = 0
        self._q = q
        self.tile_sums = _t(q.reshape(self.n_tiles, SPACING_TILE).sum(axis=1)) if e.size else torch.zeros(0, dtype=torch.int64)
        self.tile_prefix = self.total = None

    def scan(self, all_tile_sums=None):
        sums = _np(self.tile_sums if all_tile_sums is None else all_tile_sums).astype(np.int64)
        assert sums.size == self.tiles_total
        self.tile_prefix = _t(np.concatenate([[0], np.cumsum(sums)[:-1]]))
        self.total = _t(np.array([sums.sum()]))
        return self

    def window(self, j_lo, j_hi):
        t0, t1 = j_lo // SPACING_TILE, max((j_hi - 1) // SPACING_TILE + 1, j_lo // SPACING_TILE + 1)
        if self.j0 <= t0 * SPACING_TILE and t1 * SPACING_TILE <= self.j0 + self.n_tiles * SPACING_TILE:
            return self
        win = SortedSpacings(self.seed, self.stream_id, self.m, None, j0=t0 * SPACING_TILE, n_tiles=t1 - t0)
        win.tile_prefix, win.total = self.tile_prefix, self.total
        return win

    def partial_sums(self, j_lo, j_hi):
        t0 = self.j0 // SPACING_TILE
        s = np.cumsum(self._q) + int(_np(self.tile_prefix)[t0])
        return s[j_lo - self.j0:j_hi - self.j0].astype(np.uint64)

    def search(self, cdf, j_lo=0, j_hi=None):
        from oracle import sorted_resample as sr

        j_hi = self.m if j_hi is None else j_hi
        t = sr.thresholds(_np(cdf), int(_np(self.total)[0]))
        idx = np.minimum(np.searchsorted(t, self.partial_sums(j_lo, j_hi), side="right"), len(t) - 1)
        return _t(idx.astype(np.int64))


def sorted_bounds(sp, lo_all, m):
    from oracle import sorted_resample as sr

    total = int(_np(sp.total)[0])
    full = SortedSpacings(sp.seed, sp.stream_id, m, None)
    full.tile_prefix, full.total = sp.tile_prefix, sp.total
    s = full.partial_sums(0, m)
    thr = sr.thresholds(_np(lo_all), total)
    j = [int(np.searchsorted(s, t, side="left")) for t in thr]
    j[0] = 0
    return j + [m]


def search_multinomial(cdf, u, guided=None):
    return _t(_np(cdf).searchsorted(_np(u), side="right").astype(np.int64))


def weighted_moments(x, d=None, dbeta=0.0, c=0.0, idx=None, stride=1, first=0):
    xn = _np(x)
    rows = xn[first::stride] if idx is None else xn[_np(idx)[first::stride]]
    return _t(np.concatenate([[float(len(rows))], rows.sum(0), (rows * rows).sum(0)]))


def moments_to_mu_sigma(m, dims, floor=0.0):
    sw = m[0]
    mu = m[1:1 + dims] / sw
    var = m[1 + dims:1 + 2 * dims] / sw - mu * mu
    return mu.contiguous(), torch.clamp_min(torch.sqrt(torch.clamp_min(var, 0.0)), floor).contiguous()


def make_pcn_params(beta, rho, mu, sigma, n_moves, seed=0, stream0=0, index0=0, xi=None, log_u=None, slot=None,
                    total_accepts_only=False):
    return types.SimpleNamespace(beta=beta, rho=rho, mu=_np(mu), sigma=_np(sigma), n_moves=n_moves, seed=seed,
                                 stream0=stream0, index0=index0, slot=None if slot is None else _np(slot))


def pcn_sweep(x, ll, lp, lq, idx, like, prior, q, prm, out=None):
    i = slice(None) if idx is None else _np(idx)
    r = opcn.sweep(_np(x)[i], _np(ll)[i], _np(lp)[i], _np(lq)[i], like=like, prior=prior, q=q, beta=prm.beta,
                   mu=prm.mu, sigma=prm.sigma, rho=prm.rho, n_moves=prm.n_moves, seed=prm.seed,
                   stream0=prm.stream0 & 0xFFFFFFFF, index0=prm.index0, index=prm.slot)
    res = [_t(a) for a in r]
    if out is not None:
        for o, v in zip(out, res[:4]):
            o.copy_(v)
        res[:4] = out
    return tuple(res)


def evaluate(desc, x):
    return _t(dn.evaluate(desc, _np(x)))
