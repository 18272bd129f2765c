"""Affine-coupling normalising flow - numpy spec of ``asp_flow_logprob_bf16`` /
``asp_flow_sample_bf16`` (include/aspire_b200.h).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY UNPINNED against the reference's flow: the reference wraps the third-party
``zuko`` package (``src/aspire/flows/torch/flows.py:8,156-168``; unpinned in
``pyproject.toml:38-42``), which is neither under ``/root/reference`` nor installed
here.  Pinned to the reference: the call contract ``log_prob(x)`` /
``sample_and_log_prob(n)`` (``flows/torch/flows.py:327-387``), the affine data transform
``(x - mean) / std`` with ``log|J| = -sum log std`` folded into the first layer
(``src/aspire/transforms.py:618-628``) and zuko's published soft-clipped log-scale
``s / (1 + |s / log(1e-3)|)``.  Architecture (ours): T coupling layers
with alternating half masks, conditioner MLP(D/2 -> H -> H -> D) with ReLU, standard-normal base.

``emulate_bf16=True`` rounds the GEMM operands (activations and weights) to bfloat16 and
keeps everything else in fp64: the arithmetic model of the tcgen05 kernel (bf16 x bf16 ->
fp32 accumulate, fp32 epilogue), used to state the kernel's tolerance.
"""

from __future__ import annotations

import math

import numpy as np

LIMIT = math.log(1e3)
LOG_2PI = math.log(2.0 * math.pi)


def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float64."""
    f = np.asarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def soft_clip(raw):
    return raw / (1.0 + np.abs(raw / LIMIT))


def init_params(dims: int, hidden: int = 64, n_layers: int = 3, seed: int = 0, scale: float = 0.5) -> dict:
    rng = np.random.default_rng(seed)
    half = dims // 2
    layers = []
    for _ in range(n_layers):
        layers.append(dict(
            W1=rng.normal(size=(hidden, half)) * scale / math.sqrt(half), b1=rng.normal(size=hidden) * 0.1,
            W2=rng.normal(size=(hidden, hidden)) * scale / math.sqrt(hidden), b2=rng.normal(size=hidden) * 0.1,
            W3=rng.normal(size=(dims, hidden)) * scale / math.sqrt(hidden), b3=rng.normal(size=dims) * 0.1))
    return dict(dims=dims, hidden=hidden, mean=rng.normal(size=dims), std=0.5 + rng.random(dims), layers=layers)


def _conditioner(layer, za, emulate_bf16):
    q = bf16_round if emulate_bf16 else (lambda a: np.asarray(a, dtype=np.float64))
    h = np.maximum(q(za) @ q(layer["W1"]).T + layer["b1"], 0.0)
    h = np.maximum(q(h) @ q(layer["W2"]).T + layer["b2"], 0.0)
    o = q(h) @ q(layer["W3"]).T + layer["b3"]
    half = o.shape[1] // 2
    return o[:, :half], soft_clip(o[:, half:])   # shift, log-scale


def _halves(t, half):
    a = slice(0, half) if t % 2 == 0 else slice(half, 2 * half)
    b = slice(half, 2 * half) if t % 2 == 0 else slice(0, half)
    return a, b


def forward(params: dict, x: np.ndarray, emulate_bf16: bool = False):
    """x -> (z, log_prob(x))."""
    d = params["dims"]
    half = d // 2
    z = (np.asarray(x, dtype=np.float64) - params["mean"]) / params["std"]
    ladj = np.full(len(z), -np.sum(np.log(params["std"])))
    for t, layer in enumerate(params["layers"]):
        a, b = _halves(t, half)
        shift, log_s = _conditioner(layer, z[:, a], emulate_bf16)
        z = z.copy()
        z[:, b] = z[:, b] * np.exp(log_s) + shift
        ladj = ladj + log_s.sum(axis=1)
    return z, -0.5 * np.sum(z * z, axis=1) - 0.5 * d * LOG_2PI + ladj


def inverse(params: dict, u: np.ndarray, emulate_bf16: bool = False):
    """base noise u -> (x, log_prob(x))."""
    d = params["dims"]
    half = d // 2
    z = np.asarray(u, dtype=np.float64).copy()
    ladj = np.full(len(z), -np.sum(np.log(params["std"])))
    base = -0.5 * np.sum(z * z, axis=1) - 0.5 * d * LOG_2PI
    for t in reversed(range(len(params["layers"]))):
        a, b = _halves(t, half)
        shift, log_s = _conditioner(params["layers"][t], z[:, a], emulate_bf16)
        z[:, b] = (z[:, b] - shift) * np.exp(-log_s)
        ladj = ladj + log_s.sum(axis=1)
    return z * params["std"] + params["mean"], base + ladj


def pack_umma_kmajor(w: np.ndarray, k_pad: int | None = None) -> np.ndarray:
    """[N, K] weight matrix -> bf16 bit patterns (uint16) in the tcgen05 no-swizzle K-major
    core-matrix layout: element (n, k) at byte (k/8)*(N*16) + (n/8)*128 + (n%8)*16 + (k%8)*2."""
    n, k = w.shape
    k_pad = k if k_pad is None else k_pad
    wp = np.zeros((n, k_pad), dtype=np.float64)
    wp[:, :k] = w
    bits = (bf16_round(wp).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)
    out = np.zeros(n * k_pad, dtype=np.uint16)
    nn, kk = np.meshgrid(np.arange(n), np.arange(k_pad), indexing="ij")
    off = (kk // 8) * (n * 8) + (nn // 8) * 64 + (nn % 8) * 8 + (kk % 8)   # in uint16 units
    out[off.reshape(-1)] = bits.reshape(-1)
    return out
