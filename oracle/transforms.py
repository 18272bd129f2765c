"""numpy restatement of the reference's preconditioning transforms (scope row 8f-4).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: checked against the reference's own
``aspire.transforms`` imported through ``ref_shim`` (``tests/test_transforms_oracle.py``) and against the golden
vectors in ``tests/golden/transforms_*.npz`` produced from it by ``oracle/make_golden.py``.

A transform is described by a plain dict (what the CUDA kernel's ``asp_transform_t`` carries)::

    kind      int32 [D]   0 = untouched, 1 = periodic, 2 = bounded
    lower     f64   [D]   (kinds 1, 2)
    upper     f64   [D]
    bounded   "logit" | "probit" | None
    eps       clamp of the unit-interval value (default 1e-6)
    mean, std f64 [D] or None   affine stage, fitted by ``fit``

Stage order follows ``CompositeTransform`` (``src/aspire/transforms.py:271-316``): forward = periodic, bounded,
affine; inverse = affine, bounded, periodic; the log |det J| of the stages add up.
"""

from __future__ import annotations

import math

import numpy as np
from scipy.special import erf, erfinv

UNTOUCHED, PERIODIC, BOUNDED = 0, 1, 2


def make(kind, lower=None, upper=None, bounded=None, eps=1e-6, mean=None, std=None):
    kind = np.asarray(kind, dtype=np.int32)
    d = len(kind)
    lower = np.zeros(d) if lower is None else np.asarray(lower, dtype=np.float64)
    upper = np.ones(d) if upper is None else np.asarray(upper, dtype=np.float64)
    if bounded not in (None, "logit", "probit"):
        raise ValueError(f"Unknown bounded transform: {bounded}")  # transforms.py:229-232
    if (kind == BOUNDED).any():
        if bounded is None:
            raise ValueError("bounded dimensions need bounded='logit' or 'probit'")
        if ((upper - lower)[kind == BOUNDED] == 0.0).any():  # transforms.py:514-519
            raise ValueError("Current floating precision is too small for specified parameter ranges")
    return dict(kind=kind, lower=lower, upper=upper, bounded=bounded, eps=float(eps),
                mean=None if mean is None else np.asarray(mean, dtype=np.float64),
                std=None if std is None else np.asarray(std, dtype=np.float64))


def from_composite_args(parameters, periodic_parameters=None, prior_bounds=None, bounded_to_unbounded=True,
                        bounded_transform="probit", affine_transform=True, eps=1e-6):
    """The per-dimension description ``CompositeTransform.__init__`` (``transforms.py:143-248``) arrives at;
    ``affine_transform`` only says whether ``fit`` will add the affine stage."""
    periodic_parameters = list(periodic_parameters or [])
    if periodic_parameters and not prior_bounds:
        raise ValueError("Must specify prior bounds to use periodic parameters.")  # transforms.py:161-164
    d = len(parameters)
    kind, lower, upper = np.zeros(d, dtype=np.int32), np.zeros(d), np.ones(d)
    for j, p in enumerate(parameters):
        if prior_bounds is None:
            continue
        lo, hi = (float(v) for v in prior_bounds[p])
        if p in periodic_parameters:
            kind[j], lower[j], upper[j] = PERIODIC, lo, hi
        elif bounded_to_unbounded and math.isfinite(lo) and math.isfinite(hi):  # transforms.py:185-191
            kind[j], lower[j], upper[j] = BOUNDED, lo, hi
    return make(kind, lower, upper, bounded_transform if (kind == BOUNDED).any() else None, eps)


def _stages_forward(t, x):
    """periodic + bounded stages.  -> (y, log|det J|)"""
    y = np.array(x, dtype=np.float64, copy=True)
    y = np.atleast_2d(y)
    log_j = np.zeros(len(y))
    per, bnd = t["kind"] == PERIODIC, t["kind"] == BOUNDED
    if per.any():  # PeriodicTransform.forward, transforms.py:425-427
        lo, width = t["lower"][per], (t["upper"] - t["lower"])[per]
        y[:, per] = lo + (y[:, per] - lo) % width
    if bnd.any():
        lo, denom = t["lower"][bnd], (t["upper"] - t["lower"])[bnd]
        u = (y[:, bnd] - lo) / denom                               # to_unit_interval, transforms.py:476-493
        log_j = log_j - np.log(denom).sum()
        u = np.clip(u, t["eps"], 1.0 - t["eps"])
        if t["bounded"] == "logit":                                # utils.py:196-219
            y[:, bnd] = np.log(u) - np.log1p(-u)
            log_j = log_j + (-np.log(u) - np.log1p(-u)).sum(-1)
        else:                                                      # ProbitTransform.forward, transforms.py:548-556
            v = erfinv(2 * u - 1) * math.sqrt(2)
            y[:, bnd] = v
            log_j = log_j + 0.5 * (math.log(2 * math.pi) + v**2).sum(-1)
    return y, log_j


def forward(t, x):
    """``CompositeTransform.forward`` (``transforms.py:271-292``).  -> (y [N, D], log|det J| [N])"""
    y, log_j = _stages_forward(t, x)
    if t["mean"] is not None:  # AffineTransform.forward, transforms.py:624-628
        y = (y - t["mean"]) / t["std"]
        log_j = log_j - np.log(np.abs(t["std"])).sum()
    return y, log_j


def inverse(t, y):
    """``CompositeTransform.inverse`` (``transforms.py:294-316``).  -> (x [N, D], log|det J| [N])"""
    x = np.array(y, dtype=np.float64, copy=True)
    x = np.atleast_2d(x)
    log_j = np.zeros(len(x))
    per, bnd = t["kind"] == PERIODIC, t["kind"] == BOUNDED
    if t["mean"] is not None:  # AffineTransform.inverse, transforms.py:630-634
        x = x * t["std"] + t["mean"]
        log_j = log_j + np.log(np.abs(t["std"])).sum()
    if bnd.any():
        lo, denom = t["lower"][bnd], (t["upper"] - t["lower"])[bnd]
        v = x[:, bnd]
        if t["bounded"] == "logit":                                # sigmoid, utils.py:222-245
            s = np.divide(1, 1 + np.exp(-v))
            s = np.clip(s, t["eps"], 1.0 - t["eps"])
            log_j = log_j + (np.log(s) + np.log1p(-s)).sum(-1)
        else:                                                      # ProbitTransform.inverse, transforms.py:558-565
            log_j = log_j - (0.5 * (math.log(2 * math.pi) + v**2)).sum(-1)
            s = 0.5 * (1 + erf(v / math.sqrt(2)))
        x[:, bnd] = denom * s + lo                                 # from_unit_interval, transforms.py:495-512
        log_j = log_j + np.log(denom).sum()
    if per.any():  # PeriodicTransform.inverse, transforms.py:429-431
        lo, width = t["lower"][per], (t["upper"] - t["lower"])[per]
        x[:, per] = lo + (x[:, per] - lo) % width
    return x, log_j


def fit(t, x, affine=True):
    """``CompositeTransform.fit`` (``transforms.py:250-269``): the affine stage is fitted to the output of the
    other stages, mean / population standard deviation per dimension (``AffineTransform.fit``, ``:618-622``, numpy
    namespace).  -> (fitted description, transformed x)"""
    t = dict(t, mean=None, std=None)
    y, _ = _stages_forward(t, x)
    if affine:
        t["mean"], t["std"] = y.mean(0), y.std(0)
        y = (y - t["mean"]) / t["std"]
    return t, y
