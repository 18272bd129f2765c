"""Proposal flows behind the reference's ``Flow`` plug-in contract
(``src/aspire/flows/base.py:11-98``: ``log_prob``, ``sample``,
``sample_and_log_prob``, ``fit``, class attribute ``xp``).

``DiagGaussianFlow`` is the analytic diagonal-Gaussian proposal of the benchmark
configs ("q = diagonal Gaussian fitted to 2,000 draws"); its density
is a built-in ``asp_density_t`` so the fused pCN sweep can evaluate ``log_q``
in-kernel.  Any other object with the same methods works through the unfused path.
"""

from __future__ import annotations

import logging
import math

import numpy as np
import torch

from . import _lib, ops
from .history import FlowHistory
from .rng import as_uniform_source
from .targets import GaussianDensity

logger = logging.getLogger(__name__)


class Flow:
    """Minimal mirror of the reference ``Flow`` base class."""

    xp = torch

    def __init__(self, dims: int, device=None, data_transform=None, dtype=None):
        self.dims = int(dims)
        self.device = torch.device(device if device is not None else _lib.DEFAULT_DEVICE)
        self.data_transform = data_transform
        self.dtype = torch.float64

    def log_prob(self, x):
        raise NotImplementedError

    def sample(self, n_samples):
        return self.sample_and_log_prob(n_samples)[0]

    def sample_and_log_prob(self, n_samples):
        raise NotImplementedError

    def fit(self, samples, **kwargs) -> FlowHistory:
        raise NotImplementedError

    def check_error(self):
        """Kernels that report failures through a device flag raise here (``CouplingFlow``); a no-op otherwise."""


class TransformedFlow(Flow):
    """A flow that lives in the space of a data transform (``flows/torch/flows.py:327-387``: ``rescale`` /
    ``inverse_rescale``): ``log_prob(x) = inner.log_prob(T(x)) + log|det J_T(x)|``; samples are drawn in the
    transformed space and mapped back, their log density corrected by the Jacobian of the way back.  Used when
    prior bounds are given, so that the proposal has no mass outside them (``aspire.py:175-205``)."""

    def __init__(self, inner: Flow, data_transform):
        super().__init__(inner.dims, inner.device, data_transform, None)
        self.inner = inner

    def check_error(self):
        self.inner.check_error()

    def fit(self, samples, **kwargs) -> FlowHistory:
        x = samples.x if hasattr(samples, "x") else samples
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device).contiguous()
        return self.inner.fit(self.data_transform.fit(x), **kwargs)

    def log_prob(self, x):
        z, log_j = self.data_transform.forward(torch.as_tensor(x, dtype=torch.float64, device=self.device))
        return torch.as_tensor(self.inner.log_prob(z), dtype=torch.float64, device=self.device) + log_j

    def sample_and_log_prob(self, n_samples):
        z, log_q = self.inner.sample_and_log_prob(n_samples)
        x, log_j = self.data_transform.inverse(z)
        return x, torch.as_tensor(log_q, dtype=torch.float64, device=self.device) - log_j


class DiagGaussianFlow(Flow):
    """q(x) = N(mean, diag(sigma^2)); ``fit`` matches moments (population std)."""

    def __init__(self, dims: int, device=None, data_transform=None, dtype=None, mean=None, sigma=None, rng=None):
        super().__init__(dims, device, data_transform, dtype)
        self.rng = as_uniform_source(rng)
        self.density = None
        if mean is not None or sigma is not None:
            self._set(np.zeros(dims) if mean is None else mean, np.ones(dims) if sigma is None else sigma)

    def _set(self, mean, sigma):
        self.density = GaussianDensity(mean, sigma, self.dims, device=self.device)

    def fit(self, samples, **kwargs) -> FlowHistory:
        x = samples.x if hasattr(samples, "x") else samples
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device)
        if not x.is_contiguous():
            x = x.contiguous()
        m = ops.weighted_moments(x)
        mu, sigma = ops.moments_to_mu_sigma(m, self.dims)
        self._set(mu.cpu().numpy(), sigma.cpu().numpy())
        return FlowHistory()

    def _require(self):
        if self.density is None:
            raise RuntimeError("flow has not been fitted")

    def log_prob(self, x):
        self._require()
        return self.density.log_prob(torch.as_tensor(x, dtype=torch.float64, device=self.device))

    def sample_and_log_prob(self, n_samples):
        self._require()
        x = self.rng.normal(n_samples, self.dims, self.device, mean=self.density.mean, sigma=self.density.sigma)
        return x, self.density.log_prob(x)


# ---------------------------------------------------------------------------
# Affine-coupling flow on the tensor cores (scope row a12 / config C4)
# ---------------------------------------------------------------------------
SOFT_CLIP_LIMIT = 6.907755278982137  # log(1e3), zuko's MonotonicAffineTransform clip


def _bf16_bits(t: torch.Tensor) -> torch.Tensor:
    """float tensor -> bf16 bit patterns (int16 view), round-to-nearest-even."""
    return t.to(torch.float32).to(torch.bfloat16).view(torch.int16)


def pack_umma_kmajor(w: torch.Tensor, k_pad: int | None = None) -> torch.Tensor:
    """[N, K] weights -> bf16 bits in the tcgen05 no-swizzle K-major core-matrix layout:
    element (n, k) at byte (k/8)*(N*16) + (n/8)*128 + (n%8)*16 + (k%8)*2  (see oracle/flow.py)."""
    n, k = w.shape
    k_pad = k if k_pad is None else k_pad
    wp = torch.zeros(n, k_pad, dtype=torch.float32, device=w.device)
    wp[:, :k] = w
    bits = _bf16_bits(wp)
    nn = torch.arange(n, device=w.device).view(-1, 1).expand(n, k_pad)
    kk = torch.arange(k_pad, device=w.device).view(1, -1).expand(n, k_pad)
    off = (kk // 8) * (n * 8) + (nn // 8) * 64 + (nn % 8) * 8 + (kk % 8)
    out = torch.zeros(n * k_pad, dtype=torch.int16, device=w.device)
    out[off.reshape(-1)] = bits.reshape(-1)
    return out


class _NativeTrainer:
    """State of ``asp_flow_fit_step``: flat fp32 master weights, Adam moments, gradient and workspace buffers."""

    def __init__(self, flow: "CouplingFlow", batch_cap: int):
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        self.flow, self.lib, self.C = flow, lib, C
        k, n_layers, dev = flow.kdims, flow.n_layers, flow.device
        n = lib.asp_flow_fit_params(k, n_layers)
        self.w = torch.cat([torch.cat([layer[name].detach().float().reshape(-1)
                                       for name in ("W1", "b1", "W2", "b2", "W3", "b3")]) for layer in flow.layers]).contiguous()
        assert self.w.numel() == n
        self.m, self.v, self.g = (torch.zeros_like(self.w) for _ in range(3))
        self.ws = torch.zeros(lib.asp_flow_fit_workspace_floats(k, n_layers, batch_cap), dtype=torch.float32, device=dev)
        half = k // 2
        mask = torch.stack([flow._out_mask(t)[:half] for t in range(n_layers)]).to(torch.uint8).contiguous()
        self.mask = mask
        self.mean32, self.inv_std32 = flow.mean.float().contiguous(), (1.0 / flow.std).float().contiguous()
        self.log_const = 0.5 * k * math.log(2.0 * math.pi) + float(torch.log(flow.std).sum())
        t = _lib.FlowTrain()
        t.kdims, t.hidden, t.n_layers, t.batch_cap = k, flow.hidden, n_layers, int(batch_cap)
        t.d_w, t.d_m, t.d_v, t.d_g = (x.data_ptr() for x in (self.w, self.m, self.v, self.g))
        t.d_out_mask = mask.data_ptr() if flow._pad else None
        t.d_mean, t.d_inv_std, t.d_ws = self.mean32.data_ptr(), self.inv_std32.data_ptr(), self.ws.data_ptr()
        self.t, self.steps = t, 0

    def step(self, x32: torch.Tensor, lr: float, clip_grad) -> torch.Tensor:
        from ._lib import check

        x32 = x32.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=x32.device)
        self.steps += 1
        check(self.lib.asp_flow_fit_step(self.C.byref(self.t), x32.data_ptr(), x32.shape[0], lr, 0.9, 0.999, 1e-8,
                                         self.steps, float(clip_grad or 0.0), self.log_const, loss.data_ptr(),
                                         ops._stream()), "asp_flow_fit_step")
        return loss

    def loss(self, x32: torch.Tensor) -> torch.Tensor:
        from ._lib import check

        x32 = x32.contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=x32.device)
        check(self.lib.asp_flow_fit_loss(self.C.byref(self.t), x32.data_ptr(), x32.shape[0], self.log_const,
                                         loss.data_ptr(), ops._stream()), "asp_flow_fit_loss")
        return loss

    def unflatten(self, w: torch.Tensor):
        k, h, half = self.flow.kdims, self.flow.hidden, self.flow.kdims // 2
        shapes = [("W1", (h, half)), ("b1", (h,)), ("W2", (h, h)), ("b2", (h,)), ("W3", (k, h)), ("b3", (k,))]
        layers, off = [], 0
        for _ in range(self.flow.n_layers):
            layer = {}
            for name, shape in shapes:
                n = int(np.prod(shape))
                layer[name] = w[off:off + n].reshape(shape).clone()
                off += n
            layers.append(layer)
        return layers


class CouplingFlow(Flow):
    """T affine-coupling layers (alternating half masks), conditioner MLP(D/2 -> 64 -> 64 -> D)
    with ReLU, soft-clipped log-scale, standard-normal base, input standardisation.

    ``log_prob`` / ``sample_and_log_prob`` run on tcgen05 (``asp_flow_*_bf16``); ``fit`` trains the fp32 master
    weights with the library's forward / backward / Adam kernels (``asp_flow_fit_step``).

    The kernels are built for 16, 32 and 64 dimensions; any ``2 <= dims <= 64`` runs on the next kernel size with
    PADDING coordinates: the real coordinates are spread over both halves of the mask, a padding coordinate has
    mean 0 / std 1, enters as 0 and - its rows of every ``W3`` / ``b3`` being held at zero - leaves every layer as 0
    with zero log-scale, so it contributes exactly ``-1/2 log 2 pi`` to the base density, which is added back.
    (The reference's default flow works for any D: ``flows/torch/flows.py:140,164``.)
    """

    hidden = 64

    def __init__(self, dims: int, device=None, data_transform=None, dtype=None, n_layers: int = 3, rng=None,
                 seed: int = 0):
        super().__init__(dims, device, data_transform, dtype)
        if not 2 <= dims <= 64:
            raise ValueError("CouplingFlow kernels cover 2 <= dims <= 64")
        kdims = 16 if dims <= 16 else (32 if dims <= 32 else 64)
        n_first = (dims + 1) // 2  # real coordinates in the first half of the mask; the rest go to the second half
        pos = list(range(n_first)) + [kdims // 2 + j for j in range(dims - n_first)]
        self._pos = torch.tensor(pos, dtype=torch.int64, device=self.device)
        self._pad = kdims - dims
        self._real = torch.zeros(kdims, dtype=torch.bool, device=self.device)
        self._real[self._pos] = True
        self.kdims = dims = kdims  # the weights, mean / std and the kernels work in the kernel's dimension
        self.n_layers = int(n_layers)
        self.rng = as_uniform_source(rng)
        g = torch.Generator().manual_seed(seed)
        half, h = dims // 2, self.hidden
        mk = lambda *shape, s: (torch.randn(*shape, generator=g) * s).to(self.device)  # noqa: E731
        self.layers = [dict(W1=mk(h, half, s=1.0 / math.sqrt(half)), b1=torch.zeros(h, device=self.device),
                            W2=mk(h, h, s=1.0 / math.sqrt(h)), b2=torch.zeros(h, device=self.device),
                            W3=mk(dims, h, s=0.01), b3=torch.zeros(dims, device=self.device))
                       for _ in range(self.n_layers)]
        self.mean = torch.zeros(dims, dtype=torch.float64, device=self.device)
        self.std = torch.ones(dims, dtype=torch.float64, device=self.device)
        self._struct = None
        self._hold_padding()

    # -- padding ---------------------------------------------------------------
    def _out_mask(self, t: int) -> torch.Tensor:
        """Rows of layer t's W3 / b3 (shift rows then log-scale rows) that belong to real coordinates."""
        half = self.kdims // 2
        b = self._real[half:] if t % 2 == 0 else self._real[:half]  # the half this layer transforms
        return torch.cat([b, b])

    def _hold_padding(self):
        """Padding coordinates are never shifted or scaled: their W3 rows and b3 entries stay zero."""
        if not self._pad:
            return
        with torch.no_grad():
            for t, layer in enumerate(self.layers):
                m = self._out_mask(t).to(layer["W3"].dtype)
                layer["W3"].mul_(m[:, None])
                layer["b3"].mul_(m)

    def _mask_padding_grads(self, layers):
        """Gradients of the padding rows are dropped before clipping / the optimiser step (the kernels never form
        them: ``asp_flow_fit_step`` zeroes dO on those rows)."""
        if not self._pad:
            return
        for t, layer in enumerate(layers):
            m = self._out_mask(t).to(layer["W3"].dtype)
            if layer["W3"].grad is not None:
                layer["W3"].grad.mul_(m[:, None])
            if layer["b3"].grad is not None:
                layer["b3"].grad.mul_(m)

    def _embed(self, x: torch.Tensor) -> torch.Tensor:
        if not self._pad:
            return x.contiguous()
        out = torch.zeros(x.shape[0], self.kdims, dtype=x.dtype, device=x.device)
        out[:, self._pos] = x
        return out

    def _extract(self, xp: torch.Tensor) -> torch.Tensor:
        return xp if not self._pad else xp[:, self._pos].contiguous()

    @property
    def _pad_const(self) -> float:
        return 0.5 * self._pad * math.log(2.0 * math.pi)

    # -- parameters ------------------------------------------------------------
    def set_params(self, params: dict):
        """Load parameters given as numpy arrays (keys as in ``oracle/flow.py``)."""
        t32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).to(self.device)  # noqa: E731
        self.layers = [{k: t32(v) for k, v in layer.items()} for layer in params["layers"]]
        self.n_layers = len(self.layers)
        self.mean = torch.as_tensor(np.asarray(params["mean"]), dtype=torch.float64).to(self.device)
        self.std = torch.as_tensor(np.asarray(params["std"]), dtype=torch.float64).to(self.device)
        self._struct = None
        self._hold_padding()

    def struct(self):
        if self._struct is None:
            from . import _lib

            half = self.kdims // 2
            ka = max(16, half)
            blobs, bias = [], torch.zeros(self.n_layers, 3, 64, dtype=torch.float32, device=self.device)
            for i, layer in enumerate(self.layers):
                blobs += [pack_umma_kmajor(layer["W1"], ka), pack_umma_kmajor(layer["W2"]), pack_umma_kmajor(layer["W3"])]
                bias[i, 0, :self.hidden] = layer["b1"]
                bias[i, 1, :self.hidden] = layer["b2"]
                bias[i, 2, :self.kdims] = layer["b3"]
            weights = torch.cat(blobs).contiguous()
            assert weights.numel() * 2 == _lib.load().asp_flow_weights_bytes(self.kdims, self.n_layers)
            inv_std = (1.0 / self.std).contiguous()
            err = torch.zeros(1, dtype=torch.int32, device=self.device)
            s = _lib.Flow()
            s.dims, s.hidden, s.n_layers = self.kdims, self.hidden, self.n_layers
            s.d_weights, s.d_bias = weights.data_ptr(), bias.data_ptr()
            s.d_mean, s.d_inv_std, s.d_std = self.mean.data_ptr(), inv_std.data_ptr(), self.std.data_ptr()
            s.log_det0 = -float(torch.log(self.std).sum())
            s.d_error = err.data_ptr()
            s._keep = (weights, bias, inv_std, err, self.mean, self.std)
            self._struct = s
        return self._struct

    def check_error(self):
        """Raise if a tcgen05 flow kernel reported a tensor-core barrier time-out since the last check (the flag
        is then reset).  One tiny device -> host read: the samplers call it where they synchronise anyway (after
        the initial draw, after every mutation, after the importance weights), so a time-out can never feed
        garbage log q into the weights or the accept test unnoticed."""
        if getattr(self, "_struct", None) is None:
            return
        err = self._struct._keep[3]
        if int(err.item()) != 0:
            from ._lib import AspireB200Error

            err.zero_()
            raise AspireB200Error("tcgen05 flow kernel: a tensor-core barrier timed out")

    _check_error = check_error

    # -- hot path ----------------------------------------------------------------
    def log_prob(self, x):
        x = self._embed(torch.as_tensor(x, dtype=torch.float64, device=self.device))
        logq = ops.flow_logprob(self.struct(), x)
        return logq + self._pad_const if self._pad else logq

    def forward(self, x):
        x = self._embed(torch.as_tensor(x, dtype=torch.float64, device=self.device))
        logq, z = ops.flow_logprob(self.struct(), x, want_z=True)
        return self._extract(z), (logq + self._pad_const if self._pad else logq)

    def inverse(self, z):
        """Base-space points -> (x, log q(x)): the sampling direction on given z (``asp_flow_sample_bf16``)."""
        u = self._embed(torch.as_tensor(z, dtype=torch.float64, device=self.device))
        x, logq = ops.flow_sample(self.struct(), u)
        return self._extract(x), (logq + self._pad_const if self._pad else logq)

    def sample_and_log_prob(self, n_samples):
        u = self._embed(self.rng.normal(n_samples, self.dims, self.device))  # padding base coordinates: 0
        x, logq = ops.flow_sample(self.struct(), u)
        return self._extract(x), (logq + self._pad_const if self._pad else logq)

    # -- torch restatement of the forward pass: the autograd route of fit(native=False) and the reference the device
    #    fit kernels are tested against ----------------------------------------------------------------------------
    def _torch_forward(self, x32, layers):
        half = self.kdims // 2
        z = (x32 - self.mean.float()) / self.std.float()
        ladj = -torch.log(self.std.float()).sum().expand(z.shape[0])
        for t, layer in enumerate(layers):
            a = slice(0, half) if t % 2 == 0 else slice(half, self.kdims)
            b = slice(half, self.kdims) if t % 2 == 0 else slice(0, half)
            hdn = torch.relu(z[:, a] @ layer["W1"].T + layer["b1"])
            hdn = torch.relu(hdn @ layer["W2"].T + layer["b2"])
            o = hdn @ layer["W3"].T + layer["b3"]
            shift, raw = o[:, :half], o[:, half:]
            log_s = raw / (1.0 + torch.abs(raw / SOFT_CLIP_LIMIT))
            zb = z[:, b] * torch.exp(log_s) + shift
            z = torch.cat([z[:, a], zb], dim=1) if t % 2 == 0 else torch.cat([zb, z[:, a]], dim=1)
            ladj = ladj + log_s.sum(dim=1)
        return -0.5 * (z * z).sum(dim=1) - 0.5 * self.kdims * math.log(2 * math.pi) + ladj

    def fit(self, samples, n_epochs: int = 100, lr: float = 1e-3, batch_size: int = 500,
            validation_fraction: float = 0.2, clip_grad: float | None = None, lr_annealing: bool = False,
            patience: int | None = None, **kwargs) -> FlowHistory:
        """Maximum-likelihood fit with the reference's training protocol (``ZukoFlow.fit``,
        flows/torch/flows.py:170-325): affine data transform from the samples (mean, unbiased std -
        ``transforms.py:618-621``), random permutation, train / validation split, Adam, optional gradient
        clipping, cosine annealing and early stopping, best-validation weights restored at the end.
        The optimisation itself - forward, backward, Adam - runs in the library's own kernels over the flow's fp32
        master weights (``asp_flow_fit_step``, csrc/flow_fit.cu).  ``native=False`` is test
        infrastructure: torch autograd over the same tensors, the reference the kernels' gradients are checked against
        and the way the CPU suite exercises this method's host logic."""
        x = samples.x if hasattr(samples, "x") else samples
        x = torch.as_tensor(x, dtype=torch.float64, device=self.device).clone()
        if bool(torch.isnan(x).any()):
            bad = torch.isnan(x).any(dim=0).nonzero(as_tuple=True)[0].tolist()
            raise ValueError(f"Training data contains NaN values in dimensions: {bad}")
        if not bool(torch.isfinite(x).all()):
            bad = (~torch.isfinite(x)).any(dim=0).nonzero(as_tuple=True)[0].tolist()
            raise ValueError(f"Training data contains infinite values in dimensions: {bad}")
        x = self._embed(x)
        self.mean = x.mean(dim=0)
        self.std = x.std(dim=0).clamp_min(1e-12)
        if self._pad:
            self.mean = torch.where(self._real, self.mean, torch.zeros_like(self.mean))
            self.std = torch.where(self._real, self.std, torch.ones_like(self.std))
        n = x.shape[0]
        g = torch.Generator(device="cpu").manual_seed(int(kwargs.get("seed", 0)))
        x32 = x[torch.randperm(n, generator=g).to(self.device)].float()
        n_val = int(validation_fraction * n)
        x_train, x_val = (x32[:-n_val], x32[-n_val:]) if n_val > 0 else (x32, x32[:0])
        logger.info(f"Training on {x_train.shape[0]} samples, validating on {x_val.shape[0]} samples.")
        native = bool(kwargs.get("native", True))  # no silent CPU route: the default fails loudly without the library
        hist = FlowHistory()
        best_val, best_state, best_epoch = float("inf"), None, 0
        if native:
            trainer = _NativeTrainer(self, batch_size)
        else:
            params = [p.requires_grad_(True) for layer in self.layers for p in layer.values()]
            opt = torch.optim.Adam(params, lr=lr)
            sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, n_epochs) if lr_annealing else None
        for epoch in range(n_epochs):
            perm = torch.randperm(x_train.shape[0], generator=g).to(self.device)
            nb = 0
            if native:
                lr_e = 0.5 * lr * (1.0 + math.cos(math.pi * epoch / n_epochs)) if lr_annealing else lr
                tot_d = torch.zeros(1, dtype=torch.float32, device=self.device)
                for i in range(0, x_train.shape[0], batch_size):
                    tot_d += trainer.step(x_train[perm[i:i + batch_size]], lr_e, clip_grad)
                    nb += 1
                tot = float(tot_d)  # one device -> host read per epoch
            else:
                tot = 0.0
                for i in range(0, x_train.shape[0], batch_size):
                    loss = -self._torch_forward(x_train[perm[i:i + batch_size]], self.layers).mean()
                    opt.zero_grad()
                    loss.backward()
                    self._mask_padding_grads(self.layers)
                    if clip_grad is not None:
                        torch.nn.utils.clip_grad_norm_(params, clip_grad)
                    opt.step()
                    self._hold_padding()
                    tot += float(loss.detach())
                    nb += 1
                if sched is not None:
                    sched.step()
            hist.training_loss.append(tot / max(nb, 1))
            val, nv = 0.0, 0
            if native:
                val_d = torch.zeros(1, dtype=torch.float32, device=self.device)
                for i in range(0, x_val.shape[0], batch_size):
                    val_d += trainer.loss(x_val[i:i + batch_size])
                    nv += 1
                val = float(val_d) if nv else 0.0
            else:
                with torch.no_grad():
                    for i in range(0, x_val.shape[0], batch_size):
                        val += float(-self._torch_forward(x_val[i:i + batch_size], self.layers).mean())
                        nv += 1
            avg_val = val / nv if nv else hist.training_loss[-1]
            hist.validation_loss.append(avg_val)
            if avg_val < best_val:
                best_val, best_epoch = avg_val, epoch
                best_state = (trainer.w.clone() if native else
                              [{k: v.detach().clone() for k, v in layer.items()} for layer in self.layers])
            if patience is not None and epoch - best_epoch >= patience:
                logger.info(f"Early stopping triggered after {patience} epochs")
                break
        if native:
            self.layers = trainer.unflatten(best_state if best_state is not None else trainer.w)
            best_state = None if best_state is None else self.layers
        if best_state is not None:
            self.layers = best_state
            logger.info(f"Loaded best model from epoch {best_epoch} with val loss {best_val:.4f}")
        else:
            self.layers = [{k: v.detach() for k, v in layer.items()} for layer in self.layers]
        self._struct = None
        self._hold_padding()
        return hist
