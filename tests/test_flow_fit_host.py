"""``CouplingFlow.fit`` follows the reference's training protocol (``ZukoFlow.fit``,
src/aspire/flows/torch/flows.py:170-325): data transform, split, history, best-validation restore, early stop.
Runs on the CPU (the optimiser is torch autograd; only log_prob / sample need the B200 kernels)."""
import numpy as np
import pytest
import torch


def _flow(dims=16):
    from aspire_b200.flows import CouplingFlow

    return CouplingFlow(dims, device="cpu", n_layers=2, seed=3)


def test_fit_protocol_and_history():
    rng = np.random.default_rng(0)
    z = rng.normal(size=(6000, 16))
    z[:, 8:] = 0.9 * z[:, :8] + 0.3 * z[:, 8:]      # second half depends on the first: what a coupling layer learns
    x = z * np.linspace(0.5, 3.0, 16) + 2.0
    f = _flow()
    h = f.fit(x, n_epochs=12, batch_size=250, validation_fraction=0.2, lr=2e-3)
    assert len(h.training_loss) == len(h.validation_loss) == 12
    assert h.training_loss[-1] < h.training_loss[0] and min(h.validation_loss) < h.validation_loss[0]
    np.testing.assert_allclose(f.mean.numpy(), x.mean(axis=0), rtol=1e-12)
    np.testing.assert_allclose(f.std.numpy(), x.std(axis=0, ddof=1), rtol=1e-12)   # torch's unbiased std
    # the restored weights are the best-validation ones
    xv = torch.as_tensor(x, dtype=torch.float32)
    assert all(not p.requires_grad for layer in f.layers for p in layer.values())
    val_now = float(-f._torch_forward(xv, f.layers).mean())
    assert np.isfinite(val_now)


def test_fit_early_stopping_clip_and_annealing():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(600, 16))
    f = _flow()
    h = f.fit(x, n_epochs=200, batch_size=200, patience=2, clip_grad=1.0, lr_annealing=True, lr=5e-2)
    assert len(h.training_loss) < 200                  # stopped early
    best = int(np.argmin(h.validation_loss))
    assert len(h.validation_loss) - 1 - best >= 2


def test_fit_rejects_non_finite_data():
    x = np.zeros((100, 16))
    x[3, 5] = np.nan
    with pytest.raises(ValueError, match="NaN"):
        _flow().fit(x, n_epochs=1)
    x[3, 5] = np.inf
    with pytest.raises(ValueError, match="infinite"):
        _flow().fit(x, n_epochs=1)
