"""``Aspire`` facade - the call a user of the reference makes
(``src/aspire/aspire.py:33-570``): ``Aspire(...).fit(samples)`` then
``.sample_posterior(n, sampler="smc" | "importance", ...)``.

Only the dispatch that sits on the SMC / importance hot path is mirrored
(``init_flow`` :175-205, ``get_sampler_class`` :272-305, ``init_sampler`` :307-381, ``sample_posterior``
:383-570, ``sample_flow`` :891-907), including the bounded / periodic / affine preconditioning transforms
(``aspire_b200.transforms``) and flow preconditioning (a coupling flow refitted on the device at every SMC step);
the HDF5 config / flow files are not built.
"""

from __future__ import annotations

import logging
from inspect import signature
from typing import Any, Callable

import torch

from . import _lib
from .flows import CouplingFlow, DiagGaussianFlow, Flow
from .samplers import ImportanceSampler, MiniPCNSMC
from .samples import Samples

logger = logging.getLogger(__name__)

_FLOW_CLASSES = {"diag_gaussian": DiagGaussianFlow, "DiagGaussianFlow": DiagGaussianFlow,
                 "coupling": CouplingFlow, "CouplingFlow": CouplingFlow}


class Aspire:
    def __init__(
        self,
        *,
        log_likelihood: Callable,
        log_prior: Callable,
        dims: int,
        parameters: list[str] | None = None,
        periodic_parameters: list[str] | None = None,
        prior_bounds: dict | None = None,
        bounded_to_unbounded: bool = True,
        bounded_transform: str = "logit",
        device: str | None = None,
        xp: Callable | None = None,
        flow: Flow | None = None,
        flow_backend: str = "b200",
        flow_matching: bool = False,
        eps: float = 1e-6,
        dtype: Any | str | None = None,
        **kwargs,
    ) -> None:
        if (periodic_parameters or prior_bounds) and parameters is None:
            raise ValueError("prior_bounds / periodic_parameters are keyed by parameter name: pass `parameters`")
        self.periodic_parameters = list(periodic_parameters or [])
        self.prior_bounds = prior_bounds
        self.bounded_to_unbounded = bounded_to_unbounded
        self.bounded_transform = bounded_transform
        if flow_matching:
            raise NotImplementedError("flow matching is out of scope")
        self.log_likelihood = log_likelihood
        self.log_prior = log_prior
        self.dims = dims
        self.parameters = parameters
        self.device = torch.device(device if device is not None else _lib.DEFAULT_DEVICE)
        self.eps = eps
        self.flow_backend = flow_backend
        self.flow_kwargs = kwargs
        self.xp = torch
        self.dtype = torch.float64
        self._flow = flow
        self._sampler = None

    @property
    def flow(self):
        return self._flow

    @flow.setter
    def flow(self, flow):
        self._flow = flow

    @property
    def sampler(self):
        return self._sampler

    def config_dict(self, **kwargs) -> dict:
        """Subset of ``Aspire.config_dict`` (aspire.py:820-860) that concerns this path."""
        cfg = dict(dims=self.dims, parameters=self.parameters, periodic_parameters=self.periodic_parameters,
                   prior_bounds=self.prior_bounds, bounded_to_unbounded=self.bounded_to_unbounded,
                   bounded_transform=self.bounded_transform, device=str(self.device), eps=self.eps,
                   flow_backend=self.flow_backend, flow_kwargs=dict(self.flow_kwargs))
        if self._sampler is not None:
            cfg["sampler_config"] = self._sampler.config_dict(**kwargs)
        return cfg

    @property
    def n_likelihood_evaluations(self):
        return None if self._sampler is None else self._sampler.n_likelihood_evaluations

    # ---- flow ---------------------------------------------------------------
    def init_flow(self):
        kw = dict(self.flow_kwargs)
        # the reference's default is a learned flow for any D (flows/torch/flows.py:140,164): here the coupling flow
        # (tcgen05 evaluation, device fit) where its kernels cover D; flow_class="diag_gaussian" selects the
        # analytic proposal of the benchmark configs, which keeps the SMC mutation on the fused sweep
        default_class = "coupling" if 2 <= self.dims <= 64 and self.device.type == "cuda" else "diag_gaussian"
        flow_class = kw.pop("flow_class", default_class)
        if isinstance(flow_class, str):
            if flow_class not in _FLOW_CLASSES:
                raise ValueError(f"Unknown flow class {flow_class!r}; available: {sorted(_FLOW_CLASSES)}")
            flow_class = _FLOW_CLASSES[flow_class]
        self._flow = flow_class(dims=self.dims, device=self.device, dtype=self.dtype, **kw)
        if self.prior_bounds is not None and self.bounded_to_unbounded:
            # aspire.py:182-191: the flow is defined on the unbounded image of the bounded parameters (the built
            # flows standardise their input themselves, so the transform carries no affine stage)
            from .flows import TransformedFlow
            from .transforms import FlowTransform

            t = FlowTransform(parameters=self.parameters, prior_bounds=self.prior_bounds, bounded_to_unbounded=True,
                              bounded_transform=self.bounded_transform, affine_transform=False, device=self.device,
                              eps=self.eps)
            if not t.is_identity:
                self._flow = TransformedFlow(self._flow, t)

    def fit(self, samples, **kwargs):
        """``Aspire.fit`` (aspire.py:208-270): build the flow if needed and fit it to ``samples``."""
        if self._flow is None:
            self.init_flow()
        return self._flow.fit(samples, **kwargs)

    def sample_flow(self, n_samples: int = 1):
        x, log_q = self._flow.sample_and_log_prob(n_samples)
        return Samples(x=x, log_q=log_q, parameters=self.parameters, device=self.device)

    # ---- samplers -------------------------------------------------------------
    def get_sampler_class(self, sampler_type: str):
        """aspire.py:272-305 (only the two samplers on the hot path are built)."""
        if sampler_type == "importance":
            return ImportanceSampler
        if sampler_type in ("smc", "minipcn_smc", "minipcn"):
            return MiniPCNSMC
        # other names resolve through the same entry-point group as the reference (aspire.py:293-304)
        from importlib.metadata import entry_points

        eps = {ep.name.lower(): ep for ep in entry_points(group="aspire.samplers")}
        if sampler_type.lower() in eps:
            return eps[sampler_type.lower()].load()
        raise ValueError(f"Unknown sampler type: {sampler_type}. aspire_b200 builds 'importance' and 'smc'; "
                         f"entry points in 'aspire.samplers': {sorted(eps)}")

    def init_sampler(self, sampler_type: str, preconditioning=None, preconditioning_kwargs=None, **kwargs):
        """aspire.py:307-381: ``preconditioning`` None / "none" -> no transform; "default" / "standard" -> a
        ``CompositeTransform`` over the prior bounds and periodic parameters (by default without the bounded and
        affine stages, :338-341); "flow" -> ``FlowPreconditioningTransform`` (:351-366)."""
        cls = self.get_sampler_class(sampler_type)
        if sampler_type != "importance" and preconditioning is None:
            preconditioning = "default"
        preconditioning = preconditioning.lower() if preconditioning else None
        transform = None
        if preconditioning in ("standard", "default"):
            kw = dict(preconditioning_kwargs or {})
            kw.setdefault("affine_transform", False)
            kw.setdefault("bounded_to_unbounded", False)
            kw.setdefault("bounded_transform", "logit")
            if self.parameters is not None and (self.prior_bounds is not None or kw["affine_transform"]):
                from .transforms import CompositeTransform

                transform = CompositeTransform(parameters=self.parameters, prior_bounds=self.prior_bounds,
                                               periodic_parameters=self.periodic_parameters, device=self.device, **kw)
                if transform.is_identity:
                    transform = None
        elif preconditioning == "flow":
            from .transforms import FlowPreconditioningTransform

            kw = dict(preconditioning_kwargs or {})
            kw.setdefault("affine_transform", False)
            parameters = self.parameters or [f"x_{i}" for i in range(self.dims)]
            transform = FlowPreconditioningTransform(
                parameters=parameters, flow_backend=self.flow_backend, flow_kwargs=self.flow_kwargs,
                periodic_parameters=self.periodic_parameters, bounded_to_unbounded=self.bounded_to_unbounded,
                prior_bounds=self.prior_bounds, device=self.device, **kw)
        elif preconditioning not in (None, "none", "identity"):
            raise ValueError(f"Unknown preconditioning: {preconditioning}")
        if transform is not None:
            kwargs["preconditioning_transform"] = transform
        return cls(log_likelihood=self.log_likelihood, log_prior=self.log_prior, dims=self.dims,
                   prior_flow=self._flow, xp=self.xp, dtype=self.dtype, parameters=self.parameters, **kwargs)

    def sample_posterior(self, n_samples: int = 1000, sampler: str = "importance", xp: Any = None,
                         return_history: bool = False, preconditioning: str | None = None,
                         preconditioning_kwargs: dict | None = None, **kwargs):
        """aspire.py:383-570: keyword arguments are split between the sampler's
        constructor and its ``sample`` method by signature inspection (:467-480)."""
        if self._flow is None:
            raise RuntimeError("fit() must be called (or a flow passed) before sample_posterior()")
        cls = self.get_sampler_class(sampler)
        init_names = signature(cls.__init__).parameters
        init_kwargs = {k: v for k, v in kwargs.items() if k in init_names and k != "self"}
        kwargs = {k: v for k, v in kwargs.items() if k not in init_names or k == "self"}
        self._sampler = self.init_sampler(sampler, preconditioning=preconditioning,
                                          preconditioning_kwargs=preconditioning_kwargs, **init_kwargs)
        samples = self._sampler.sample(n_samples, **kwargs)
        if xp is not None:
            samples = samples.to_namespace(xp)
        if hasattr(samples, "parameters"):
            samples.parameters = self.parameters or samples.parameters
        logger.info(f"Number of likelihood evaluations: {self.n_likelihood_evaluations}")
        if return_history:
            return samples, self._sampler.history
        return samples
