"""Sampler contract of the drop-in boundary (reference ``src/aspire/samplers/base.py:19-147``):
same constructor, likelihood-call counter and ``config_dict``; checkpoint helpers are out of
scope."""

from __future__ import annotations

import functools
import logging
from typing import Any, Callable

import torch

from ..samples import Samples

logger = logging.getLogger(__name__)


def track_calls(fn):
    """Record the arguments of every call (reference ``utils.py:1003-1030``) for ``config_dict``."""

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        calls = self.__dict__.setdefault("_sample_calls", [])
        calls.append(dict(args=args, kwargs=dict(kwargs)))
        return fn(self, *args, **kwargs)

    return wrapper


class Sampler:
    def __init__(
        self,
        log_likelihood: Callable,
        log_prior: Callable,
        dims: int,
        prior_flow,
        xp: Callable = None,
        dtype: Any | str | None = None,
        parameters: list[str] | None = None,
        preconditioning_transform: Callable | None = None,
    ):
        self.prior_flow = prior_flow
        self._log_likelihood = log_likelihood
        self._log_prior = log_prior
        self.dims = dims
        self.xp = torch
        self.backend_str = "torch"
        self.dtype = torch.float64
        self.parameters = parameters
        self.history = None
        self.n_likelihood_evaluations = 0
        if preconditioning_transform is not None and not getattr(preconditioning_transform, "is_identity", False):
            raise NotImplementedError(
                "aspire_b200 samplers run in the original parameter space; preconditioning / bounded-parameter "
                "transforms are a later scope row (SURVEY 8f-4)")
        self.preconditioning_transform = None
        self.device = torch.device(getattr(prior_flow, "device", None) or "cuda")

    def log_likelihood(self, samples) -> torch.Tensor:
        """Counts evaluations like the reference (samplers/base.py:81-87)."""
        self.n_likelihood_evaluations += len(samples)
        return self._as_vec(self._log_likelihood(samples), len(samples))

    def log_prior(self, samples) -> torch.Tensor:
        return self._as_vec(self._log_prior(samples), len(samples))

    def _as_vec(self, v, n):
        t = torch.as_tensor(v, dtype=torch.float64, device=self.device).reshape(-1)
        if t.numel() != n:
            raise ValueError(f"callable returned {t.numel()} values for {n} samples")
        return t.contiguous()

    @track_calls
    def sample(self, n_samples: int) -> Samples:
        raise NotImplementedError

    def config_dict(self, include_sample_calls: str | bool = "last") -> dict:
        config = {"sampler_class": self.__class__.__name__}
        calls = self.__dict__.get("_sample_calls", [])
        if include_sample_calls is not False and calls:
            config["sample_calls"] = calls[-1] if str(include_sample_calls).lower() == "last" else {
                i: c for i, c in enumerate(calls)}
        return config
