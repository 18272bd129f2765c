"""``Samples`` / ``SMCSamples`` over fp64 CUDA tensors - the output type of the
drop-in boundary (reference: ``src/aspire/samples.py:36-595, 1208-1332``).

Same field names, methods and error behaviour as the reference dataclasses; the
arithmetic runs in the sm_100a kernels (``aspire_b200.ops``).  ``xp`` is
``torch`` and every array field is a contiguous fp64 tensor on one CUDA device.
"""

from __future__ import annotations

import logging
import math
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch

from . import _lib, ops
from .rng import as_uniform_source

logger = logging.getLogger(__name__)


def _as_dev(a, device, dtype=torch.float64):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=dtype)
    else:
        t = torch.as_tensor(np.asarray(a), dtype=dtype).to(device)
    return t.contiguous()


class HostSamples(dict):
    """numpy copy of a sample set: a dict whose keys are also attributes (``host.x``, ``host["x"]``)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __len__(self):
        x = self.get("x")
        return 0 if x is None else len(x)


@dataclass
class BaseSamples:
    """reference ``BaseSamples`` (samples.py:36-413)."""

    x: Any
    log_likelihood: Any = None
    log_prior: Any = None
    log_q: Any = None
    parameters: list | None = None
    dtype: Any = None
    xp: Any = None
    device: Any = None

    def __post_init__(self):
        self.xp = torch
        self.dtype = torch.float64  # the kernels compute in fp64 (reference parity mode)
        if self.device is None:
            self.device = (self.x.device if isinstance(self.x, torch.Tensor) and self.x.is_cuda
                           else torch.device(_lib.DEFAULT_DEVICE))
        self.device = torch.device(self.device)
        self.x = _as_dev(self.x, self.device)
        if self.x.ndim == 1:
            self.x = self.x.reshape(-1, 1)
        self.log_likelihood = _as_dev(self.log_likelihood, self.device)
        self.log_prior = _as_dev(self.log_prior, self.device)
        self.log_q = _as_dev(self.log_q, self.device)
        if self.parameters is None:
            self.parameters = [f"x_{i}" for i in range(self.dims)]

    @property
    def dims(self):
        return self.x.shape[1] if self.x.ndim > 1 else 1

    def __len__(self):
        return self.x.shape[0]

    def array_to_namespace(self, x, dtype=None):
        return _as_dev(x, self.device)

    def _common(self):
        return dict(parameters=self.parameters, device=self.device)

    def __getitem__(self, idx):
        """Row selection (samples.py:327-339): int64 index tensors go through
        the gather kernel, masks / slices through torch indexing."""
        if isinstance(idx, torch.Tensor) and idx.dtype == torch.int64 and idx.is_cuda:
            out = ops.gather_rows(self.x, idx, self.log_likelihood, self.log_prior, self.log_q)
            return self.__class__(x=out[0], log_likelihood=out[1], log_prior=out[2], log_q=out[3], **self._common())
        sel = lambda t: None if t is None else t[idx]  # noqa: E731
        return self.__class__(x=self.x[idx], log_likelihood=sel(self.log_likelihood), log_prior=sel(self.log_prior),
                              log_q=sel(self.log_q), **self._common())

    def __setitem__(self, idx, value):
        raise NotImplementedError("Setting items is not supported")

    @classmethod
    def concatenate(cls, samples: list):
        if not samples:
            raise ValueError("No samples to concatenate")
        if not all(s.parameters == samples[0].parameters for s in samples):
            raise ValueError("Parameters do not match")
        cat = lambda name: (torch.cat([getattr(s, name) for s in samples], dim=0)  # noqa: E731
                            if all(getattr(s, name) is not None for s in samples) else None)
        return cls(x=cat("x"), log_likelihood=cat("log_likelihood"), log_prior=cat("log_prior"), log_q=cat("log_q"),
                   parameters=samples[0].parameters, device=samples[0].device)

    @classmethod
    def from_samples(cls, samples, **kwargs):
        kwargs.pop("xp", None)
        kwargs.pop("dtype", None)
        device = kwargs.pop("device", samples.device)
        return cls(x=samples.x, log_likelihood=samples.log_likelihood, log_prior=samples.log_prior,
                   log_q=samples.log_q, parameters=samples.parameters, device=device, **kwargs)

    def _numpy_fields(self):
        cv = lambda t: None if t is None else t.detach().cpu().numpy()  # noqa: E731
        return dict(x=cv(self.x), log_likelihood=cv(self.log_likelihood), log_prior=cv(self.log_prior),
                    log_q=cv(self.log_q), parameters=self.parameters)

    def to_numpy(self):
        """Host copy (samples.py:281-302: the reference returns a numpy-backed dataclass of the same fields): a
        ``HostSamples`` - numpy arrays reachable as attributes (``.x``, ``.log_likelihood`` ...) like the reference's
        object, and as dictionary items (what the checkpoint payload and ``to_dict`` use)."""
        return HostSamples(self._numpy_fields())

    def to_dict(self, flat: bool = True):
        """samples.py:142-181: every field (host copies) plus one key per parameter (``flat``) or a ``"samples"``
        dictionary of parameter columns."""
        d = self.to_numpy()
        xs = d.pop("x")
        columns = {p: xs[:, i] for i, p in enumerate(self.parameters)}
        if flat:
            d.update(columns)
        else:
            d["samples"] = columns
        return d

    @classmethod
    def from_dict(cls, dictionary: dict, device=None):
        """samples.py:182-206: the inverse of ``to_dict`` for both layouts; derived quantities (weights, evidence
        of ``Samples``) are recomputed by the constructor, not read back."""
        import dataclasses

        d = dict(dictionary)
        parameters = d.pop("parameters", None)
        if "samples" in d:
            columns = d.pop("samples")
            parameters = sorted(columns) if parameters is None else parameters
        else:
            if parameters is None:
                raise ValueError("Parameters must be provided if samples are not nested in a 'samples' key")
            columns = {p: d.pop(p) for p in parameters}
        x = np.stack([np.asarray(columns[p]) for p in parameters], axis=-1)
        init = {f.name for f in dataclasses.fields(cls) if f.init} - {"x", "parameters", "xp", "dtype", "device"}
        return cls(x=x, parameters=list(parameters), device=device, **{k: v for k, v in d.items() if k in init})

    def to_dataframe(self, include: list | None = None):
        """samples.py:208-250: parameter columns plus the per-sample fields named in ``include`` (default: the
        three log densities that are set)."""
        import pandas as pd

        d = self.to_dict(flat=True)
        names = list(self._dataframe_fields) if include is None else list(include)
        data = {p: d[p] for p in self.parameters}
        for k in names:  # a field that is not set becomes a column of NaN, as in the reference
            data[k] = d[k] if d.get(k) is not None else np.full(len(self.x), np.nan)
        return pd.DataFrame(data)

    _dataframe_fields = ("log_likelihood", "log_prior", "log_q")

    def __str__(self):
        return f"No. samples: {len(self.x)}\nNo. parameters: {self.x.shape[-1]}\n"  # samples.py:276-280

    # pickling swaps the namespace module for its name (samples.py:394-413)
    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_slope_cache", None)  # device scratch of SMCSamples, rebuilt on demand
        state["xp"] = "torch"
        state["device"] = None if self.device is None else str(self.device)
        return state

    def __setstate__(self, state):
        state = dict(state)
        state["xp"] = torch
        if state.get("device") is not None:
            state["device"] = torch.device(state["device"])
        self.__dict__.update(state)


@dataclass
class Samples(BaseSamples):
    """reference ``Samples`` (samples.py:416-595)."""

    log_w: Any = field(init=False, default=None)
    weights: Any = field(init=False, default=None)
    evidence: float = field(init=False, default=None)
    evidence_error: float = field(init=False, default=None)
    log_evidence: float | None = None
    log_evidence_error: float | None = None
    effective_sample_size: float = field(init=False, default=None)

    def __post_init__(self):
        super().__post_init__()
        if all(v is not None for v in (self.log_likelihood, self.log_prior, self.log_q)):
            self.compute_weights()

    @property
    def efficiency(self):
        if self.log_w is None:
            raise RuntimeError("Samples do not contain weights!")
        return self.effective_sample_size / len(self.x)

    def compute_weights(self):
        """samples.py:457-475 on the device (``asp_importance_weights``)."""
        self.log_w, self.weights, result = ops.importance_weights(self.log_likelihood, self.log_prior, self.log_q)
        r = result.cpu().numpy()
        self.log_evidence = float(r[0])
        self.evidence = float(r[1])
        self.evidence_error = float(r[2])
        self.log_evidence_error = float(r[3])
        self.effective_sample_size = float(r[4])

    @property
    def scaled_weights(self):
        return torch.exp(self.log_w - torch.max(self.log_w))

    def rejection_sample(self, rng=None):
        """samples.py:481-495."""
        if rng is None:
            rng = np.random.default_rng()
        log_u = _as_dev(np.log(rng.uniform(size=len(self.x))), self.device)
        accept = (self.log_w - torch.max(self.log_w)) > log_u
        return self.__class__(x=self.x[accept], log_likelihood=self.log_likelihood[accept],
                              log_prior=self.log_prior[accept], **self._common())

    def to_namespace(self, xp):
        name = getattr(xp, "__name__", str(xp))
        if "numpy" in name:
            return self.to_numpy()
        if "torch" in name:
            return self
        raise ValueError(f"aspire_b200 samples live on the GPU; cannot convert to namespace {name}")

    def to_numpy(self):
        d = self._numpy_fields()
        d.update(log_evidence=self.log_evidence, log_evidence_error=self.log_evidence_error)
        if self.log_w is not None:
            d.update(log_w=self.log_w.cpu().numpy(), weights=self.weights.cpu().numpy(), evidence=self.evidence,
                     evidence_error=self.evidence_error, effective_sample_size=self.effective_sample_size)
        return d

    def __getitem__(self, idx):
        sliced = super().__getitem__(idx)
        sliced.log_evidence = self.log_evidence
        sliced.log_evidence_error = self.log_evidence_error
        return sliced

    _dataframe_fields = ("log_likelihood", "log_prior", "log_q", "log_w")  # samples.py:559-577

    def __str__(self):  # samples.py:506-515
        out = super().__str__()
        if self.log_evidence is not None and self.log_evidence_error is not None:
            out += f"Log evidence: {self.log_evidence:.2f} +/- {self.log_evidence_error:.2f}\n"
        if self.log_w is not None:
            out += f"Effective sample size: {self.effective_sample_size:.1f}\nEfficiency: {self.efficiency:.2f}\n"
        return out


@dataclass
class SMCSamples(BaseSamples):
    """reference ``SMCSamples`` (samples.py:1208-1332)."""

    beta: float | None = None
    log_evidence: float | None = None
    log_evidence_error: float | None = None

    # -- cached reweighting slope d = (logL + logP) - log_q ------------------
    def _slope(self):
        key = (self.log_likelihood.data_ptr(), self.log_prior.data_ptr(), self.log_q.data_ptr())
        cache = self.__dict__.get("_slope_cache")
        if cache is None or cache[0] != key:
            d, stats = ops.reweight(self.log_likelihood, self.log_prior, self.log_q)
            cache = (key, d, stats, stats.cpu().numpy())
            self.__dict__["_slope_cache"] = cache
        return cache[1], cache[2], cache[3]

    def _check_nan(self, beta, st):
        dbeta = beta - self.beta
        if st[2] > 0 or (dbeta == 0.0 and (st[3] > 0 or st[0] == -math.inf)) or st[3] > 0 and dbeta < 0:
            raise ValueError(f"Log weights contain NaN values for beta={beta}")  # samples.py:1246-1247

    def _sums(self, beta):
        d, stats, st = self._slope()
        self._check_nan(beta, st)
        dbeta = beta - self.beta
        s = ops.ess_sums(d, [dbeta], stats).cpu().numpy()[0]
        c = dbeta * (st[1] if dbeta >= 0 else st[0]) if dbeta != 0.0 else 0.0
        return d, dbeta, c, float(s[0]), float(s[1])

    def log_p_t(self, beta):
        """samples.py:1217-1219"""
        return (1 - beta) * self.log_q + beta * (self.log_likelihood + self.log_prior)

    def unnormalized_log_weights(self, beta: float):
        """samples.py:1221-1224"""
        return (self.beta - beta) * self.log_q + (beta - self.beta) * (self.log_likelihood + self.log_prior)

    def log_evidence_ratio(self, beta: float) -> float:
        """samples.py:1226-1228"""
        _, _, c, s1, _ = self._sums(beta)
        return c + math.log(s1) - math.log(len(self.x))

    def log_evidence_ratio_variance(self, beta: float) -> float:
        """samples.py:1230-1242"""
        d, dbeta, c, s1, _ = self._sums(beta)
        n = len(self.x)
        _, result, _ = ops.smc_weights(d, dbeta, c, s1, want_p=False, want_tiles=False)
        mean = s1 / n
        return float(result[0]) / n / (n * mean * mean) if mean != 0 else math.nan

    def effective_sample_size(self, beta: float) -> float:
        """``effective_sample_size(log_weights(beta))`` (utils.py:510-512)."""
        _, _, _, s1, s2 = self._sums(beta)
        return s1 * s1 / s2

    def log_weights(self, beta: float):
        """samples.py:1244-1249 (NaN guard; the normaliser is ADDED, as in the reference)."""
        ler = self.log_evidence_ratio(beta)
        return self.unnormalized_log_weights(beta) + ler

    def resample_weights(self, beta: float):
        """The ``p`` handed to ``rng.choice`` (samples.py:1274-1277) + the scan's tile statistics."""
        d, dbeta, c, s1, _ = self._sums(beta)
        p, _, tiles = ops.smc_weights(d, dbeta, c, s1)
        return p, tiles

    def resample_indices(self, beta, n_samples=None, rng=None, method: str = "multinomial"):
        """Ancestor indices (int64 CUDA tensor), or None for the same-beta shortcut."""
        n = len(self.x)
        n_samples = n if n_samples is None else n_samples
        if beta == self.beta:
            if n_samples == n:
                return None
            p = torch.full((n,), 1.0 / n, dtype=torch.float64, device=self.device)  # samples.py:1273
            tiles = None
        else:
            p, tiles = self.resample_weights(beta)
        cdf = ops.cdf_exact(p, tile_stats=tiles)
        src = as_uniform_source(rng)
        if method == "multinomial":
            return src.multinomial_ancestors(cdf, n_samples)
        if method == "systematic":
            return ops.search_systematic(cdf, src.scalar_uniform(), n_samples)
        raise ValueError(f"unknown resampling method {method!r}")

    def resample(self, beta, n_samples: int | None = None, rng=None, method: str = "multinomial"):
        """samples.py:1251-1287.  With ``rng`` a ``numpy.random.Generator`` the
        ancestors are bit-identical to the reference's ``rng.choice`` for the
        same weights (its uniforms are drawn on the host and uploaded)."""
        idx = self.resample_indices(beta, n_samples, rng, method)
        if idx is None:
            logger.warning("Resampling with the same beta value, returning identical samples")
            return self
        out = ops.gather_rows(self.x, idx, self.log_likelihood, self.log_prior, self.log_q)
        return self.__class__(x=out[0], log_likelihood=out[1], log_prior=out[2], log_q=out[3], beta=beta,
                              **self._common())

    def to_standard_samples(self):
        """samples.py:1295-1305: no ``log_q`` -> the returned Samples carry no weights."""
        return Samples(x=self.x, log_likelihood=self.log_likelihood, log_prior=self.log_prior,
                       parameters=self.parameters, device=self.device, log_evidence=self.log_evidence,
                       log_evidence_error=self.log_evidence_error)

    def to_numpy(self):
        d = self._numpy_fields()
        d.update(beta=self.beta, log_evidence=self.log_evidence, log_evidence_error=self.log_evidence_error)
        return d

    def __str__(self):  # samples.py:1289-1293
        out = super().__str__()
        if self.log_evidence is not None:
            out += f"Log evidence: {self.log_evidence:.2f}\n"
        return out

    def __getitem__(self, idx):
        sliced = super().__getitem__(idx)
        sliced.beta = self.beta
        sliced.log_evidence = self.log_evidence
        sliced.log_evidence_error = self.log_evidence_error
        return sliced
