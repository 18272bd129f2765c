"""Sampler contract of the drop-in boundary (reference ``src/aspire/samplers/base.py:19-147``):
same constructor, likelihood-call counter, ``config_dict`` and checkpoint helpers
(``samplers/base.py:158-276``: the state is a picklable dict; files are pickle blobs - the
reference's HDF5 container needs ``h5py``, which this image does not have)."""

from __future__ import annotations

import functools
import logging
import pickle
from pathlib import Path
from typing import Any, Callable

import torch

from .. import _lib
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
        # samplers/base.py:63-66; None / an identity transform keep the fused kernels of the samplers
        self.preconditioning_transform = preconditioning_transform
        self.device = torch.device(getattr(prior_flow, "device", None) or _lib.DEFAULT_DEVICE)

    def fit_preconditioning_transform(self, x):
        """samplers/base.py:68-75."""
        if self.preconditioning_transform is None:
            return x
        return self.preconditioning_transform.fit(torch.as_tensor(x, dtype=torch.float64, device=self.device))

    def log_likelihood(self, samples) -> torch.Tensor:
        """Counts evaluations like the reference (samplers/base.py:81-87)."""
        self.n_likelihood_evaluations += len(samples)
        return self._as_vec(self._log_likelihood(samples), len(samples))

    def log_prior(self, samples) -> torch.Tensor:
        return self._as_vec(self._log_prior(samples), len(samples))

    def _check_flow(self) -> None:
        """Flow kernels report failures (a tensor-core barrier time-out) through a device flag; the samplers read it
        where they synchronise anyway and raise ``AspireB200Error`` instead of carrying garbage log q forward."""
        check = getattr(self.prior_flow, "check_error", None)
        if check is not None:
            check()

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

    # ---- checkpointing (reference samplers/base.py:158-276) ---------------------
    def _checkpoint_extra_state(self) -> dict:
        return {}

    def _restore_extra_state(self, state: dict) -> None:
        _ = state

    def build_checkpoint_state(self, samples, iteration: int | None = None, meta: dict | None = None,
                               include_sample_calls: str | bool = "last") -> dict:
        """Picklable payload with the reference's keys; ``samples`` is the host copy (``to_numpy()``)."""
        host = samples.to_numpy() if hasattr(samples, "to_numpy") else samples
        state = {"sampler": self.__class__.__name__, "iteration": iteration, "samples": host,
                 "config": self.config_dict(include_sample_calls=include_sample_calls), "parameters": self.parameters,
                 "meta": meta or {}}
        state.update(self._checkpoint_extra_state())
        return state

    def serialize_checkpoint(self, state: dict, protocol: int | None = None) -> bytes:
        return pickle.dumps(state, protocol=pickle.HIGHEST_PROTOCOL if protocol is None else int(protocol))

    def default_checkpoint_callback(self, state: dict) -> None:
        self._last_checkpoint_state = state
        self._last_checkpoint_bytes = self.serialize_checkpoint(state)

    def default_file_checkpoint_callback(self, file_path) -> Callable[[dict], None]:
        """Overwrite ``file_path`` with the latest checkpoint.  ``.pkl`` files hold the pickled state;
        ``.h5`` / ``.hdf5`` (the reference's container: dataset ``/checkpoint/state``) need h5py."""
        if file_path is None:
            return self.default_checkpoint_callback
        file_path = Path(file_path)
        is_hdf = file_path.name.lower().endswith((".h5", ".hdf5"))
        if is_hdf:
            try:
                import h5py  # noqa: F401
            except ImportError as e:  # pragma: no cover - depends on the image
                raise RuntimeError("HDF5 checkpoints need h5py; use a .pkl checkpoint_file_path") from e

        def _callback(state: dict) -> None:
            self.default_checkpoint_callback(state)
            if is_hdf:
                import h5py
                import numpy as np

                with h5py.File(file_path, "a") as f:
                    grp = f.require_group("checkpoint")
                    if "state" in grp:
                        del grp["state"]
                    grp.create_dataset("state", data=np.frombuffer(self._last_checkpoint_bytes, dtype=np.uint8))
            else:
                tmp = file_path.with_suffix(file_path.suffix + ".tmp")
                with open(tmp, "wb") as f:
                    f.write(self._last_checkpoint_bytes)
                tmp.replace(file_path)

        return _callback

    def load_checkpoint_from_file(self, file_path, h5_path: str = "checkpoint", dsetname: str = "state") -> dict:
        file_path = Path(file_path)
        if file_path.name.lower().endswith((".h5", ".hdf5")):
            import h5py

            with h5py.File(file_path, "r") as f:
                blob = f[h5_path][dsetname][...].tobytes()
        else:
            with open(file_path, "rb") as f:
                blob = f.read()
        return pickle.loads(blob)

    def restore_from_checkpoint(self, source):
        """-> (host sample dict, state) from a file path, pickled bytes or a state dict."""
        if isinstance(source, (str, Path)):
            state = self.load_checkpoint_from_file(source)
        elif isinstance(source, bytes):
            state = pickle.loads(source)
        elif isinstance(source, dict):
            state = source
        else:
            raise TypeError("Unsupported checkpoint source type.")
        saved = state.get("samples")
        if saved is None:
            raise ValueError("Checkpoint missing samples.")
        self._restore_extra_state(state)
        return saved, state

    @property
    def last_checkpoint_state(self):
        return getattr(self, "_last_checkpoint_state", None)

    @property
    def last_checkpoint_bytes(self):
        return getattr(self, "_last_checkpoint_bytes", None)
