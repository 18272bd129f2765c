"""``Aspire`` facade - the call a user of the reference makes
(``src/aspire/aspire.py:33-570``): ``Aspire(...).fit(samples)`` then
``.sample_posterior(n, sampler="smc" | "importance", ...)``.

Only the dispatch that sits on the SMC / importance hot path is mirrored
(``get_sampler_class`` :272-305, ``init_sampler`` :307-381, ``sample_posterior``
:383-570, ``sample_flow`` :891-907); checkpointing, HDF5 config and the
bounded-parameter preconditioning transforms are out of scope.
"""

from __future__ import annotations

import logging
from inspect import signature
from typing import Any, Callable

import torch

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
        if periodic_parameters or prior_bounds:
            raise NotImplementedError("periodic / bounded parameter transforms are scope row 8f-4 (SURVEY.md)")
        if flow_matching:
            raise NotImplementedError("flow matching is out of scope")
        self.log_likelihood = log_likelihood
        self.log_prior = log_prior
        self.dims = dims
        self.parameters = parameters
        self.device = torch.device(device if device is not None else "cuda")
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
        cfg = dict(dims=self.dims, parameters=self.parameters, device=str(self.device), eps=self.eps,
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
        flow_class = kw.pop("flow_class", "diag_gaussian")
        if isinstance(flow_class, str):
            if flow_class not in _FLOW_CLASSES:
                raise ValueError(f"Unknown flow class {flow_class!r}; available: {sorted(_FLOW_CLASSES)}")
            flow_class = _FLOW_CLASSES[flow_class]
        self._flow = flow_class(dims=self.dims, device=self.device, dtype=self.dtype, **kw)

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
        if preconditioning not in (None, "default", "identity"):
            raise NotImplementedError("preconditioning transforms are scope row 8f-4 (SURVEY.md)")
        cls = self.get_sampler_class(sampler_type)
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
