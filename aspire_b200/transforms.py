"""Preconditioning / bounded-parameter transforms on the device (scope row 8f-4).

Same classes, constructor arguments and method contracts as the reference's ``src/aspire/transforms.py``
(``fit(x) -> y``, ``forward(x) -> (y, log|det J|)``, ``inverse(y) -> (x, log|det J|)``, ``new_instance``,
``config_dict``) so that a sampler written against one runs against the other; the arithmetic is one CUDA kernel
(``asp_transform_forward`` / ``asp_transform_inverse``, ``csrc/transform.cu``) instead of a chain of masked
array-API updates.  Tensors are fp64 CUDA tensors; there is no CPU path.  Spec: ``oracle/transforms.py``, pinned
against the reference's own classes.

``FlowPreconditioningTransform`` (``transforms.py:649-748``) refits a coupling flow to the ensemble at every SMC
step with the device fit (``asp_flow_fit_step``, row 8f-1) and maps through it on tcgen05.  Not built here: the
HDF5 ``save`` / ``load`` (``h5py`` absent; ``state_dict`` carries the same content)."""

from __future__ import annotations

import math
from typing import Any

import torch

from . import _lib, ops

UNTOUCHED, PERIODIC, BOUNDED = 0, 1, 2
_BOUNDED_KINDS = {None: _lib.BOUNDED_NONE, "logit": _lib.BOUNDED_LOGIT, "probit": _lib.BOUNDED_PROBIT}


class BaseTransform:
    """``transforms.py:24-122``.  ``xp`` / ``dtype`` are accepted for signature compatibility; the carrier is
    always a torch fp64 CUDA tensor."""

    name = "base"
    requires_prior_bounds = False
    is_identity = False

    def __init__(self, xp=None, dtype: Any = None, device=None):
        self.xp = torch
        self.dtype = torch.float64
        self.device = torch.device(device if device is not None else _lib.DEFAULT_DEVICE)

    def _rows(self, x) -> torch.Tensor:
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device)
        return (x[None, :] if x.dim() == 1 else x).contiguous()

    def fit(self, x):
        raise NotImplementedError

    def forward(self, x):
        raise NotImplementedError

    def inverse(self, y):
        raise NotImplementedError

    def config_dict(self) -> dict:
        return {"name": self.name, "dtype": "float64"}

    def new_instance(self, xp=None, dtype: Any = None):
        return self.__class__(device=self.device,
                              **{k: v for k, v in self.config_dict().items() if k not in ("name", "dtype")})


class IdentityTransform(BaseTransform):
    """``transforms.py:125-139``."""

    name = "identity"
    is_identity = True

    def fit(self, x):
        return self._rows(x)

    def forward(self, x):
        x = self._rows(x)
        return x, torch.zeros(len(x), dtype=torch.float64, device=self.device)

    inverse = forward

    def new_instance(self, xp=None, dtype: Any = None):
        return IdentityTransform(device=self.device)


class _KernelTransform(BaseTransform):
    """Everything below is the one kernel with a different per-dimension description."""

    def _setup(self, kind, lower, upper, bounded: str | None, eps: float):
        if bounded not in _BOUNDED_KINDS:
            raise ValueError(f"Unknown bounded transform: {bounded}")  # transforms.py:229-232
        self._kind_host = [int(k) for k in kind]
        dev = self.device
        self._kind = torch.tensor(self._kind_host, dtype=torch.int32, device=dev) if any(self._kind_host) else None
        self._lower = torch.as_tensor(lower, dtype=torch.float64, device=dev).reshape(-1).contiguous()
        self._upper = torch.as_tensor(upper, dtype=torch.float64, device=dev).reshape(-1).contiguous()
        self._bounded = bounded if BOUNDED in self._kind_host else None
        self.eps = float(eps)
        self._mean = self._std = None
        if self._bounded is not None:
            lo, hi = self._lower.cpu(), self._upper.cpu()
            mask = torch.tensor([k == BOUNDED for k in self._kind_host])
            if bool(((hi - lo)[mask] == 0.0).any()):  # transforms.py:514-519
                raise ValueError("Current floating precision (float64) is too small for specified parameter ranges")

    @property
    def dims(self) -> int:
        return len(self._kind_host)

    def _apply(self, x, inverse: bool, affine: bool = True):
        x = self._rows(x)
        if x.shape[1] != self.dims:
            raise ValueError(f"expected rows of {self.dims} values, got {x.shape[1]}")
        mean, std = (self._mean, self._std) if affine else (None, None)
        return ops.transform_apply(x, self._kind, self._lower, self._upper, _BOUNDED_KINDS[self._bounded], self.eps,
                                   mean, std, inverse=inverse)

    def fit(self, x):
        return self.forward(x)[0]

    def forward(self, x):
        return self._apply(x, inverse=False)

    def inverse(self, y):
        return self._apply(y, inverse=True)


class PeriodicTransform(_KernelTransform):
    """``transforms.py:411-437``: ``lower + (x - lower) mod (upper - lower)`` both ways, unit Jacobian."""

    name = "periodic"
    requires_prior_bounds = True

    def __init__(self, lower, upper, xp=None, dtype=None, device=None):
        super().__init__(xp, dtype, device)
        lower = torch.as_tensor(lower, dtype=torch.float64).reshape(-1)
        self._setup([PERIODIC] * len(lower), lower, upper, None, 0.0)

    def config_dict(self):
        return super().config_dict() | {"lower": self._lower.tolist(), "upper": self._upper.tolist()}


class BoundedTransform(_KernelTransform):
    """``transforms.py:440-534``: [lower, upper] -> [0, 1] -> the real line."""

    name = "bounded"
    requires_prior_bounds = True
    _which: str | None = None

    def __init__(self, lower, upper, xp=None, eps: float = 1e-6, dtype=None, device=None):
        super().__init__(xp, dtype, device)
        if self._which is None:
            raise NotImplementedError("Subclasses must implement forward method.")
        lower = torch.atleast_1d(torch.as_tensor(lower, dtype=torch.float64)).reshape(-1)
        self._setup([BOUNDED] * len(lower), lower, torch.atleast_1d(torch.as_tensor(upper, dtype=torch.float64)),
                    self._which, eps)

    def config_dict(self):
        return super().config_dict() | {"lower": self._lower.tolist(), "upper": self._upper.tolist(), "eps": self.eps}


class LogitTransform(BoundedTransform):
    """``transforms.py:573-606``."""

    name = "logit"
    _which = "logit"


class ProbitTransform(BoundedTransform):
    """``transforms.py:537-570``."""

    name = "probit"
    _which = "probit"


class AffineTransform(_KernelTransform):
    """``transforms.py:609-646``: standardise every dimension with the mean / standard deviation of the rows it is
    fitted to (population standard deviation, as the reference computes in its numpy namespace)."""

    name = "affine"

    def __init__(self, xp=None, dtype=None, device=None, dims: int | None = None):
        super().__init__(xp, dtype, device)
        self._kind_host = [UNTOUCHED] * (dims or 0)

    def fit(self, x):
        x = self._rows(x)
        self._setup([UNTOUCHED] * x.shape[1], torch.zeros(x.shape[1]), torch.ones(x.shape[1]), None, 0.0)
        self._mean = x.mean(0).contiguous()
        self._std = x.std(0, unbiased=False).contiguous()
        self.log_abs_det_jacobian = -torch.log(torch.abs(self._std)).sum()
        return self.forward(x)[0]

    def _apply(self, x, inverse, affine=True):
        if self._mean is None:
            raise RuntimeError("AffineTransform must be fitted first")
        return super()._apply(x, inverse, affine)

    def config_dict(self):
        return BaseTransform.config_dict(self)

    def state_dict(self) -> dict:
        return {"mean": self._mean, "std": self._std}


class CompositeTransform(_KernelTransform):
    """``transforms.py:142-358``: periodic wrap, bounded -> unbounded, affine - chosen per parameter from the prior
    bounds exactly as the reference does (finite two-sided bounds and not periodic -> bounded, ``:185-191``)."""

    name = "composite"

    def __init__(self, parameters, periodic_parameters=None, prior_bounds=None, bounded_to_unbounded: bool = True,
                 bounded_transform: str = "probit", affine_transform: bool = True, device=None, xp=None,
                 eps: float = 1e-6, dtype: Any = None):
        super().__init__(xp, dtype, device)
        if periodic_parameters and not prior_bounds:
            raise ValueError("Must specify prior bounds to use periodic parameters.")  # transforms.py:161-164
        if bounded_transform not in ("probit", "logit"):
            raise ValueError(f"Unknown bounded transform: {bounded_transform}")
        self.parameters = list(parameters)
        self.periodic_parameters = list(periodic_parameters or [])
        self.bounded_to_unbounded = bounded_to_unbounded
        self.bounded_transform = bounded_transform
        self.affine_transform = affine_transform
        d = len(self.parameters)
        kind, lower, upper = [UNTOUCHED] * d, [0.0] * d, [1.0] * d
        if prior_bounds is None:
            self.prior_bounds = None
            self.bounded_parameters = None
        else:
            self.prior_bounds = {k: [float(v) for v in torch.as_tensor(prior_bounds[k], dtype=torch.float64).tolist()]
                                 for k in self.parameters}
            self.bounded_parameters = [] if bounded_to_unbounded else None
            for j, p in enumerate(self.parameters):
                lo, hi = self.prior_bounds[p]
                if p in self.periodic_parameters:
                    kind[j], lower[j], upper[j] = PERIODIC, lo, hi
                elif bounded_to_unbounded and math.isfinite(lo) and math.isfinite(hi):
                    kind[j], lower[j], upper[j] = BOUNDED, lo, hi
                    self.bounded_parameters.append(p)
        self._setup(kind, lower, upper, bounded_transform, eps)

    @property
    def is_identity(self) -> bool:  # nothing to apply: the samplers keep their fused kernels
        return not any(self._kind_host) and not self.affine_transform

    def fit(self, x):
        """``transforms.py:250-269``: the affine stage is fitted to the output of the other stages."""
        if not self.affine_transform:
            return self._apply(x, inverse=False, affine=False)[0]
        y0 = self._apply(x, inverse=False, affine=False)[0]
        self._mean = y0.mean(0).contiguous()
        self._std = y0.std(0, unbiased=False).contiguous()
        return self._apply(x, inverse=False)[0]

    def _apply(self, x, inverse, affine=True):
        if affine and self.affine_transform and self._mean is None:
            raise RuntimeError("CompositeTransform with an affine stage must be fitted first")
        return super()._apply(x, inverse, affine and self.affine_transform)

    def new_instance(self, xp=None, dtype: Any = None):
        return self.__class__(parameters=self.parameters, periodic_parameters=self.periodic_parameters,
                              prior_bounds=self.prior_bounds, bounded_to_unbounded=self.bounded_to_unbounded,
                              bounded_transform=self.bounded_transform, affine_transform=self.affine_transform,
                              device=self.device, eps=self.eps)

    def config_dict(self):
        return BaseTransform.config_dict(self) | {
            "parameters": self.parameters, "periodic_parameters": self.periodic_parameters,
            "prior_bounds": self.prior_bounds, "bounded_to_unbounded": self.bounded_to_unbounded,
            "bounded_transform": self.bounded_transform, "affine_transform": self.affine_transform,
            "eps": self.eps, "device": str(self.device)}

    def state_dict(self) -> dict:
        return {"affine_transform": {"mean": self._mean, "std": self._std}} if self.affine_transform else {}


class FlowTransform(CompositeTransform):
    """``transforms.py:361-409``: the data transform of a flow - a ``CompositeTransform`` without periodic
    parameters."""

    def __init__(self, parameters, prior_bounds=None, bounded_to_unbounded: bool = True,
                 bounded_transform: str = "probit", affine_transform: bool = True, device=None, xp=None,
                 eps: float = 1e-6, dtype=None):
        super().__init__(parameters=parameters, periodic_parameters=[], prior_bounds=prior_bounds,
                         bounded_to_unbounded=bounded_to_unbounded, bounded_transform=bounded_transform,
                         affine_transform=affine_transform, device=device, xp=xp, eps=eps, dtype=dtype)

    def new_instance(self, xp=None, dtype=None):
        return self.__class__(parameters=self.parameters, prior_bounds=self.prior_bounds,
                              bounded_to_unbounded=self.bounded_to_unbounded,
                              bounded_transform=self.bounded_transform, affine_transform=self.affine_transform,
                              device=self.device, eps=self.eps)

    def config_dict(self):
        cfg = super().config_dict()
        cfg.pop("periodic_parameters", None)
        return cfg


class FlowPreconditioningTransform(BaseTransform):
    """``transforms.py:649-748``: preconditioning by a normalising flow that is refitted to the ensemble at every
    SMC step - ``fit(x)`` builds and trains a new flow on x and returns the base-space points, ``forward`` /
    ``inverse`` map through it with the log-Jacobian of the whole chain (bounded / periodic data transform, then the
    flow).  The flow is the library's ``CouplingFlow``: trained by the device kernels (``asp_flow_fit_step``),
    evaluated on tcgen05 (``asp_flow_logprob_bf16`` gives z and log q(x), ``asp_flow_sample_bf16`` the way back);
    ``log|det J| = log q(x) - log N(z)`` in either direction.  The coupling MLPs run in bf16 / fp32, so the
    round trip x -> z -> x closes to ~1e-6, like the reference's fp32 flows."""

    name = "flow_preconditioning"

    def __init__(self, parameters, flow_backend: str = "b200", prior_bounds=None, bounded_to_unbounded: bool = True,
                 bounded_transform: str = "probit", affine_transform: bool = True, periodic_parameters=None,
                 device=None, xp=None, eps: float = 1e-6, dtype: Any = None, flow_matching: bool = False,
                 flow_kwargs: dict | None = None, fit_kwargs: dict | None = None):
        super().__init__(xp, dtype, device)
        if flow_matching:
            raise NotImplementedError("flow matching is out of scope")
        self.parameters = list(parameters)
        self.periodic_parameters = list(periodic_parameters or [])
        self.prior_bounds = prior_bounds
        self.bounded_to_unbounded = bounded_to_unbounded
        self.bounded_transform = bounded_transform
        self.affine_transform = affine_transform
        self.eps = eps
        self.flow_backend = flow_backend
        self.flow_matching = flow_matching
        self.flow_kwargs = {k: v for k, v in dict(flow_kwargs or {}).items() if k not in ("flow_class", "dtype")}
        self.fit_kwargs = dict(fit_kwargs or {})
        # the flow standardises its input itself, so the data transform keeps its affine stage only when asked
        self._data_transform = CompositeTransform(
            parameters=self.parameters, periodic_parameters=self.periodic_parameters, prior_bounds=prior_bounds,
            bounded_to_unbounded=bounded_to_unbounded, bounded_transform=bounded_transform,
            affine_transform=affine_transform, device=self.device, eps=eps)
        self.flow = None
        self.history = None

    def _log_normal(self, z: torch.Tensor) -> torch.Tensor:
        return -0.5 * (z * z).sum(dim=1) - 0.5 * z.shape[1] * math.log(2.0 * math.pi)

    def fit(self, x):
        from .flows import CouplingFlow

        x = self._rows(x)
        self.flow = CouplingFlow(dims=len(self.parameters), device=self.device, **self.flow_kwargs)
        y = x if self._data_transform.is_identity else self._data_transform.fit(x)
        self.history = self.flow.fit(y, **self.fit_kwargs)
        return self.forward(x)[0]

    def _require(self):
        if self.flow is None:
            raise RuntimeError("FlowPreconditioningTransform must be fitted first")

    def forward(self, x):
        self._require()
        x = self._rows(x)
        y, lj = (x, 0.0) if self._data_transform.is_identity else self._data_transform.forward(x)
        z, log_q = self.flow.forward(y)
        return z, lj + (log_q - self._log_normal(z))

    def inverse(self, z):
        self._require()
        z = self._rows(z)
        y, log_q = self.flow.inverse(z)
        lj_flow = self._log_normal(z) - log_q
        if self._data_transform.is_identity:
            return y, lj_flow
        x, lj = self._data_transform.inverse(y)
        return x, lj + lj_flow

    def new_instance(self, xp=None, dtype: Any = None):
        return self.__class__(parameters=self.parameters, periodic_parameters=self.periodic_parameters,
                              prior_bounds=self.prior_bounds, bounded_to_unbounded=self.bounded_to_unbounded,
                              bounded_transform=self.bounded_transform, affine_transform=self.affine_transform,
                              device=self.device, eps=self.eps, flow_backend=self.flow_backend,
                              flow_matching=self.flow_matching, flow_kwargs=self.flow_kwargs, fit_kwargs=self.fit_kwargs)

    def config_dict(self):
        return BaseTransform.config_dict(self) | {
            "parameters": self.parameters, "periodic_parameters": self.periodic_parameters,
            "prior_bounds": self.prior_bounds, "bounded_to_unbounded": self.bounded_to_unbounded,
            "bounded_transform": self.bounded_transform, "affine_transform": self.affine_transform,
            "eps": self.eps, "device": str(self.device), "flow_backend": self.flow_backend,
            "flow_kwargs": self.flow_kwargs, "fit_kwargs": self.fit_kwargs}

    def save(self, h5_file, path="data_transform"):
        raise NotImplementedError("FlowPreconditioningTransform does not support save method yet.")  # transforms.py:745-748
