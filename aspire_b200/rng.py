"""Randomness sources for the samplers.

The reference passes a ``numpy.random.Generator`` (or ``orng.ArrayRNG``) around
(``src/aspire/samplers/smc/base.py:49,64,77``).  Two sources exist here:

``HostGenerator``  wraps a ``numpy.random.Generator``: draws on the host in the
                   reference's order and uploads - PARITY mode; with the same
                   seed the ancestors are bit-identical to the reference's.
``Pcg64Generator`` a ``HostGenerator`` whose bit generator is PCG64 (numpy's default): the
                   resampling uniforms - the bulk draw of the step - are generated ON THE
                   DEVICE from the generator's (state, inc) (``asp_pcg64_uniform``) and the
                   host generator is advanced by the same count, so the stream, and hence
                   the ancestors, are bit-identical to the reference's with no host draw or
                   upload.  Scalars and normals still come from the host generator.
``PhiloxRNG``      device-side counter-based Philox4x32 (``asp_philox_*``, the sorted
                   resampling uniforms of ``asp_sorted_spacings`` and the in-kernel
                   stream of ``asp_pcn_sweep``) - THROUGHPUT mode; results are
                   independent of launch geometry and of how particles are sharded
                   across GPUs.
"""

from __future__ import annotations

import os

import numpy as np
import torch

from . import ops


class PhiloxRNG:
    """Counter-based device RNG: every draw takes a fresh stream id."""

    def __init__(self, seed: int | None = None):
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self._stream = 0

    def next_stream(self, count: int = 1) -> int:
        s = self._stream
        self._stream += count
        if self._stream >= 2**32:
            raise OverflowError("Philox stream ids exhausted")
        return s

    def uniform(self, n: int, device, index0: int = 0) -> torch.Tensor:
        return ops.philox_uniform(self.seed, self.next_stream(), n, device, index0=index0)

    def multinomial_ancestors(self, cdf: torch.Tensor, m: int) -> torch.Tensor:
        """``searchsorted(cdf, u, 'right')`` at ``m`` iid uniforms.  Throughput mode generates the uniforms already
        sorted (``asp_sorted_spacings`` / ``asp_search_sorted``, ``oracle/sorted_resample.py``): the same multiset
        law, ancestors ascending, every HBM access sequential."""
        return ops.search_multinomial_sorted(cdf, self.seed, self.next_stream(), m)

    def scalar_uniform(self) -> float:
        # one host-side Philox-free draw is enough for the systematic offset: derive it from the seed/stream
        s = self.next_stream()
        return float(np.random.default_rng([self.seed & 0xFFFFFFFF, self.seed >> 32, s]).random())

    def normal(self, n: int, dims: int, device, mean=None, sigma=None, index0: int = 0) -> torch.Tensor:
        return ops.philox_normal(self.seed, self.next_stream(), n, dims, device, mean=mean, sigma=sigma, index0=index0)

    @property
    def is_host(self) -> bool:
        return False


class HostGenerator:
    """numpy ``Generator`` draws, uploaded (reference-order parity mode)."""

    def __init__(self, rng: np.random.Generator):
        self.rng = rng

    def uniform(self, n: int, device, index0: int = 0) -> torch.Tensor:
        return torch.from_numpy(self.rng.random(n)).to(device)

    def multinomial_ancestors(self, cdf: torch.Tensor, m: int) -> torch.Tensor:
        """The reference's draw order: ``searchsorted(cdf, rng.random(m), 'right')`` (``Generator.choice``)."""
        return ops.search_multinomial(cdf, self.uniform(m, cdf.device))

    def scalar_uniform(self) -> float:
        return float(self.rng.random())

    def normal(self, n: int, dims: int, device, mean=None, sigma=None, index0: int = 0) -> torch.Tensor:
        z = torch.from_numpy(self.rng.standard_normal((n, dims))).to(device)
        if sigma is not None:
            z = z * sigma
        if mean is not None:
            z = z + mean
        return z

    @property
    def is_host(self) -> bool:
        return True


class Pcg64Generator(HostGenerator):
    """``HostGenerator`` with the bulk uniforms drawn on the device (same PCG64 stream)."""

    def uniform(self, n: int, device, index0: int = 0) -> torch.Tensor:
        bg = self.rng.bit_generator
        st = bg.state
        if st.get("has_uint32", 0):  # a buffered 32-bit half would be consumed first: keep numpy's order
            return super().uniform(n, device, index0)
        u = ops.pcg64_uniform(int(st["state"]["state"]), int(st["state"]["inc"]), n, device)
        bg.advance(n)
        return u


def _is_pcg64(rng) -> bool:
    bg = getattr(rng, "bit_generator", None)
    return bg is not None and type(bg).__name__ == "PCG64" and hasattr(bg, "advance")


def as_uniform_source(rng):
    if rng is None:
        return PhiloxRNG()
    if isinstance(rng, (PhiloxRNG, HostGenerator)):
        return rng
    if isinstance(rng, np.random.Generator):
        return Pcg64Generator(rng) if _is_pcg64(rng) else HostGenerator(rng)
    if isinstance(rng, (int, np.integer)):
        return PhiloxRNG(int(rng))
    if hasattr(rng, "random"):
        return HostGenerator(rng)
    raise TypeError(f"unsupported rng {type(rng)!r}")
