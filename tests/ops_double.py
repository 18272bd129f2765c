"""Test doubles for ``aspire_b200.ops`` (the ctypes layer over the C-ABI): plain torch / numpy implementations of
the same contracts, installed by the ``ops_double`` fixture of tests/test_samplers_host.py so that the HOST logic of
the samplers - argument handling, tempering loop, history, evidence bookkeeping, the three mutation routes, the
preconditioned route, the ``Aspire`` plumbing - runs in the CPU suite.  Like tests/cpu_backend.py (the double
behind the gloo tests) this is test infrastructure only: the product has no CPU path and these functions are
never importable from ``aspire_b200``.  In-kernel noise is counter-based like the kernels' (a numpy generator keyed by
seed, stream and purpose - reproducible, restartable from a checkpoint) but is NOT the Philox stream, so tests built
on it are statistical or structural; supplied noise (parity mode) is used as given."""
import math

import numpy as np
import torch

from oracle import densities as dn
from oracle import transforms as ot

F64 = torch.float64


class Doubles:
    def __init__(self, seed=1234):
        self.salt = seed

    def _rng(self, seed, stream, purpose):
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        return np.random.default_rng([self.salt, seed & 0xFFFFFFFF, seed >> 32, int(stream) & 0xFFFFFFFF, purpose])

    # ---- plumbing -----------------------------------------------------------------
    @staticmethod
    def _f64(t, name):
        assert isinstance(t, torch.Tensor) and t.dtype == F64, name
        return t if t.is_contiguous() else t.contiguous()

    @staticmethod
    def _i64(t, name):
        assert isinstance(t, torch.Tensor) and t.dtype == torch.int64, name
        return t

    @staticmethod
    def _stream():
        return 0

    @staticmethod
    def device_check():
        return None

    # ---- reweighting / bisection / weights ------------------------------------------
    @staticmethod
    def reweight(ll, lp, lq):
        d = ll + lp - lq
        stats = [float(d[~torch.isnan(d)].min()), float(d[~torch.isnan(d)].max()), float(torch.isnan(d).sum()),
                 float((d == math.inf).sum())]
        return d, torch.tensor(stats, dtype=F64)

    @staticmethod
    def ess_sums(d, dbetas, stats):
        out = []
        for db in dbetas:
            c = db * (float(stats[1]) if db >= 0 else float(stats[0]))
            e = torch.exp(db * d - c)
            out.append([float(e.sum()), float((e * e).sum())])
        return torch.tensor(out, dtype=F64)

    @staticmethod
    def beta_search(d, stats, beta_prev, target_eff, tol, n_total=None, allreduce=None, multi_level=True, peer=None):
        from types import SimpleNamespace

        n = d.numel() if n_total is None else n_total

        def ev(b):
            e = torch.exp((b - beta_prev) * (d - float(stats[1])))
            return float(e.sum()), float((e * e).sum())

        st = SimpleNamespace(n_evals=1, min_valid=0, s1_min=0.0, s2_min=0.0, beta_min=beta_prev, reserved=0, phase=2)
        s1, s2 = ev(1.0)
        st.eff_beta_max = s1 * s1 / s2 / n
        if st.eff_beta_max >= target_eff:
            st.beta_min, st.s1_min, st.s2_min, st.min_valid = 1.0, s1, s2, 1
            return st
        lo, hi = beta_prev, 1.0
        while hi - lo > tol:
            mid = 0.5 * (lo + hi)
            s1, s2 = ev(mid)
            st.n_evals += 1
            if s1 * s1 / s2 / n >= target_eff:
                lo, st.s1_min, st.s2_min, st.min_valid = mid, s1, s2, 1
            else:
                hi = mid
        st.beta_min = lo
        return st

    @staticmethod
    def smc_weights(d, dbeta, c, s1, n_total=None, want_p=True, want_tiles=True):
        n = d.numel()
        e = torch.exp(dbeta * d - c)
        mean = s1 / n
        result = torch.tensor([float(((e - mean) ** 2).sum()), float((e / s1).sum()), 0.0, 0.0], dtype=F64)
        return (e / s1 if want_p else None), result, None

    @staticmethod
    def importance_weights(ll, lp, lq):
        lw = ll + lp - lq
        w = torch.exp(lw)
        n = lw.numel()
        z = float(w.mean())
        err = float(torch.sqrt(((w - z) ** 2).sum() / (n * max(n - 1, 1))))
        res = torch.tensor([math.log(z) if z > 0 else -math.inf, z, err, abs(err / z) if z > 0 else math.nan,
                            float(w.sum() ** 2 / (w * w).sum()), float(lw.max()), float(torch.isnan(lw).sum()), 0.0],
                           dtype=F64)
        return lw, w, res

    # ---- resampling -------------------------------------------------------------------
    @staticmethod
    def cdf_exact(p, normalise=True, tile_stats=None, return_info=False):
        c = torch.from_numpy(np.cumsum(p.numpy()))
        return c / c[-1] if normalise else c

    @staticmethod
    def search_multinomial(cdf, u, guided=None):
        return torch.searchsorted(cdf, u, right=True).clamp_max(cdf.numel() - 1)

    @staticmethod
    def search_multinomial_sorted(cdf, seed, stream_id, m):
        from oracle import sorted_resample

        return torch.from_numpy(sorted_resample.resample(cdf.numpy(), seed, stream_id, m))

    @staticmethod
    def search_systematic(cdf, u0, n):
        u = (torch.arange(n, dtype=F64) + u0) / n
        return torch.searchsorted(cdf, u, right=True).clamp_max(cdf.numel() - 1)

    @staticmethod
    def gather_rows(x, idx, a=None, b=None, c=None):
        return (x[idx].contiguous(), *(None if t is None else t[idx].contiguous() for t in (a, b, c)))

    @staticmethod
    def weighted_moments(x, d=None, dbeta=0.0, c=0.0, idx=None, stride=1, first=0):
        x = x[idx[first::stride]] if idx is not None else x[first::stride]
        return torch.cat([torch.tensor([float(len(x))], dtype=F64), x.sum(0), (x * x).sum(0)])

    @staticmethod
    def moments_to_mu_sigma(m, dims, floor=0.0):
        mu = m[1:1 + dims] / m[0]
        var = m[1 + dims:1 + 2 * dims] / m[0] - mu * mu
        return mu.contiguous(), torch.clamp_min(torch.sqrt(torch.clamp_min(var, 0.0)), floor).contiguous()

    # ---- RNG ----------------------------------------------------------------------------
    def philox_uniform(self, seed, stream_id, n, device, index0=0):
        return torch.from_numpy(self._rng(seed, stream_id, 0).random(index0 + n)[index0:])

    def philox_normal(self, seed, stream_id, n, dims, device, mean=None, sigma=None, index0=0):
        z = torch.from_numpy(self._rng(seed, stream_id, 1).standard_normal((index0 + n, dims))[index0:])
        if sigma is not None:
            z = z * sigma
        if mean is not None:
            z = z + mean
        return z

    @staticmethod
    def pcg64_uniform(state, inc, n, device, offset=0):
        bg = np.random.PCG64()
        st = bg.state
        st["state"] = {"state": int(state), "inc": int(inc)}
        bg.state = st
        bg.advance(int(offset))
        return torch.from_numpy(np.random.Generator(bg).random(n))

    # ---- pCN moves (the semi-fused kernels are reported as unsupported) ---------------------
    class Prm:
        pass

    def make_pcn_params(self, beta, rho, mu, sigma, n_moves, seed=0, stream0=0, index0=0, xi=None, log_u=None,
                        slot=None, total_accepts_only=False, nu=0.0, gamma=None):
        p = self.Prm()
        p.beta, p.rho, p.mu, p.sigma, p.n_moves, p.nu = beta, rho, mu, sigma, n_moves, nu
        p.reserved, p.scale, p.xi, p.log_u, p.gamma = (1 if total_accepts_only else 0), None, xi, log_u, gamma
        p.seed, p.stream0 = seed, stream0
        return p

    def tpcn_scales(self, x, prm, move):
        n, d = x.shape
        r = (((x - prm.mu) / prm.sigma) ** 2).sum(1)
        g = prm.gamma[move] if prm.gamma is not None else torch.from_numpy(
            self._rng(prm.seed, prm.stream0 + move, 2).gamma(0.5 * (d + prm.nu), 1.0, n))
        prm.scale = torch.sqrt(0.5 * (prm.nu + r) / g)
        return prm.scale

    def pcn_propose(self, x, prm, move):
        xi = prm.xi[move] if prm.xi is not None else torch.from_numpy(
            self._rng(prm.seed, prm.stream0 + move, 3).standard_normal(tuple(x.shape)))
        sc = 1.0 if prm.scale is None else prm.scale[:, None]
        return (prm.mu + math.sqrt(1 - prm.rho ** 2) * (x - prm.mu) + prm.rho * prm.sigma * xi * sc).contiguous()

    def pcn_accept(self, x, ll, lp, lq, xp, ll2, lp2, lq2, prm, move, accepts):
        b, dims = prm.beta, x.shape[1]

        def ref(v):
            r = (((v - prm.mu) / prm.sigma) ** 2).sum(1)
            return 0.5 * r if prm.nu == 0 else 0.5 * (prm.nu + dims) * torch.log1p(r / prm.nu)

        def tempered(l_, p_, q_):
            t = (1 - b) * q_ + b * (l_ + p_)
            return torch.where(torch.isnan(t), torch.full_like(t, -math.inf), t)

        la = tempered(ll2, lp2, lq2) - tempered(ll, lp, lq) + ref(xp) - ref(x)
        log_u = prm.log_u[move] if prm.log_u is not None else torch.log(torch.from_numpy(
            1 - self._rng(prm.seed, prm.stream0 + move, 4).random(len(x))))
        acc = log_u < la
        x[acc], ll[acc], lp[acc], lq[acc] = xp[acc], ll2[acc], lp2[acc], lq2[acc]
        accepts[0 if prm.reserved else move] += int(acc.sum())

    def pcn_sweep(self, x, ll, lp, lq, idx, like, prior, q, prm, out=None):
        if idx is not None:
            x, ll, lp, lq = x[idx], ll[idx], lp[idx], lq[idx]
        x, ll, lp, lq = (t.clone().contiguous() for t in (x, ll, lp, lq))
        accepts = torch.zeros(max(prm.n_moves, 1), dtype=torch.int64)
        for t in range(prm.n_moves):
            if prm.nu > 0:
                self.tpcn_scales(x, prm, t)
            xp = self.pcn_propose(x, prm, t)
            self.pcn_accept(x, ll, lp, lq, xp, self.logdensity(like, xp), self.logdensity(prior, xp),
                            self.logdensity(q, xp), prm, t, accepts)
        return x, ll, lp, lq, (accepts[:1] if prm.reserved else accepts[:prm.n_moves])

    @staticmethod
    def pcn_propose_eval_supported(dims, like, prior):
        return False

    @staticmethod
    def student_t_fit(x, mu0, sigma0, idx=None, stride=1, m=None, n_iter=30, nu0=10.0, nu_min=1.0, nu_max=1e6):
        """oracle/pcn.py::fit_student_t_diag on the selected rows, from the given starting point."""
        from scipy.special import digamma

        rows = x[::stride] if idx is None else x[idx[::stride]]
        if m is not None:
            rows = rows[:m]
        n, d = rows.shape
        mu, sigma, nu = mu0.clone(), sigma0.clone(), float(nu0)
        for _ in range(n_iter):
            z = (rows - mu) / sigma
            r2 = (z * z).sum(dim=1)
            w = (nu + d) / (nu + r2)
            mu = (w[:, None] * rows).sum(dim=0) / w.sum()
            sigma = torch.sqrt((w[:, None] * (rows - mu) ** 2).sum(dim=0) / n)
            const = 1.0 + float((torch.log(w) - w).mean()) + float(digamma(0.5 * (nu + d))) - math.log(0.5 * (nu + d))

            def f(v):
                return -float(digamma(0.5 * v)) + math.log(0.5 * v) + const

            lo, hi = nu_min, nu_max
            if f(hi) > 0:
                nu = hi
            elif f(lo) < 0:
                nu = lo
            else:
                for _ in range(60):
                    mid = math.sqrt(lo * hi)
                    if f(mid) > 0:
                        lo = mid
                    else:
                        hi = mid
                nu = math.sqrt(lo * hi)
        return mu.contiguous(), sigma.contiguous(), float(nu)

    @staticmethod
    def pcn_sweep_flow_supported(dims, like, prior):
        return False

    @staticmethod
    def pcn_sweep_tpcn_supported(dims):
        return dims in (4, 8, 16, 32, 64)

    @staticmethod
    def logdensity(dens, x):
        return torch.from_numpy(dn.evaluate(dens.describe(), x.numpy()))

    # ---- transforms ---------------------------------------------------------------------
    @staticmethod
    def transform_apply(x, kind, lower, upper, bounded, eps, mean=None, std=None, inverse=False, want_log_j=True):
        d = x.shape[1]
        t = ot.make(np.zeros(d, np.int32) if kind is None else kind.numpy(), None if lower is None else lower.numpy(),
                    None if upper is None else upper.numpy(), {0: None, 1: "logit", 2: "probit"}[bounded], eps,
                    None if mean is None else mean.numpy(), None if std is None else std.numpy())
        if not (t["kind"] == 2).any():
            t["bounded"] = None
        if x.shape[0] == 0:
            return x.clone(), torch.zeros(0, dtype=F64)
        y, lj = (ot.inverse if inverse else ot.forward)(t, x.numpy())
        return torch.from_numpy(y), torch.from_numpy(lj)


NAMES = ["_f64", "_i64", "_stream", "device_check", "reweight", "ess_sums", "beta_search", "smc_weights",
         "importance_weights", "cdf_exact", "search_multinomial", "search_multinomial_sorted", "search_systematic", "gather_rows",
         "weighted_moments", "moments_to_mu_sigma", "philox_uniform", "philox_normal", "pcg64_uniform", "make_pcn_params",
         "tpcn_scales", "pcn_propose", "pcn_accept", "pcn_sweep", "pcn_propose_eval_supported", "pcn_sweep_tpcn_supported", "pcn_sweep_flow_supported", "student_t_fit",
         "logdensity",
         "transform_apply"]


def install(monkeypatch, seed=1234):
    """Replace the ctypes layer by the doubles for the duration of one test."""
    from aspire_b200 import _lib, ops

    dbl = Doubles(seed)
    for name in NAMES:
        monkeypatch.setattr(ops, name, getattr(dbl, name))
    monkeypatch.setattr(_lib, "DEFAULT_DEVICE", "cpu")
    return dbl
