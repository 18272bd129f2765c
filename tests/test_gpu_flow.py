"""Coupling-flow log_prob / sample on tcgen05 (asp_flow_*_bf16) against the numpy spec
(oracle/flow.py).  Stated tolerance: the kernel multiplies bf16 operands and accumulates in
fp32; against the bf16-emulating oracle log_q agrees to 5e-3 absolute for >= 99 % of the
particles and to 0.05 for all (an activation that sits on a bf16 rounding boundary may round
the other way under a different fp32 summation order), against the exact fp64 flow to 0.15
absolute (mean 0.03) at D = 64 - the bound quoted in DESIGN.md."""
import math

import numpy as np
import pytest

from oracle import flow as of

pytestmark = pytest.mark.gpu


def _t(a, dev):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a), device=dev)


def _flow(dims, n_layers, cuda, seed=0):
    from aspire_b200.flows import CouplingFlow
    from aspire_b200.rng import PhiloxRNG

    p = of.init_params(dims, n_layers=n_layers, seed=seed)
    f = CouplingFlow(dims, device=cuda, n_layers=n_layers, rng=PhiloxRNG(3))
    f.set_params(p)
    return f, p


@pytest.mark.parametrize("dims,n_layers,n", [(64, 3, 5000), (64, 1, 128), (64, 4, 77), (32, 3, 4097), (32, 2, 1), (16, 3, 3001), (16, 2, 130)])
def test_flow_logprob_matches_spec(cuda, dims, n_layers, n):
    f, p = _flow(dims, n_layers, cuda)
    rng = np.random.default_rng(n)
    x = rng.normal(size=(n, dims)) * p["std"] * 1.5 + p["mean"]
    z, logq = f.forward(_t(x, cuda))
    f._check_error()
    z_e, lq_e = of.forward(p, x, emulate_bf16=True)
    z_x, lq_x = of.forward(p, x)
    logq, z = logq.cpu().numpy(), z.cpu().numpy()
    assert np.isfinite(logq).all()
    d = np.abs(logq - lq_e)                                         # same arithmetic model
    assert (d < 5e-3).mean() >= 0.99 and d.max() < 0.05, (d.max(), (d < 5e-3).mean())
    dz = np.abs(z - z_e).max(axis=1)
    assert (dz < 2e-3).mean() >= 0.99 and dz.max() < 0.05
    err = np.abs(logq - lq_x)                                       # stated bf16 bound vs exact fp64
    assert err.max() < 0.15 and err.mean() < 0.03, (err.max(), err.mean())
    np.testing.assert_array_equal(f.log_prob(_t(x, cuda)).cpu().numpy(), logq)  # deterministic


@pytest.mark.parametrize("dims", [16, 32, 64])
def test_flow_sample_inverts_logprob(cuda, dims):
    from aspire_b200 import ops

    f, p = _flow(dims, 3, cuda, seed=5)
    n = 3000
    u = np.random.default_rng(1).normal(size=(n, dims))
    x, logq = ops.flow_sample(f.struct(), _t(u, cuda))
    f._check_error()
    x_e, lq_e = of.inverse(p, u, emulate_bf16=True)
    dx = np.abs(x.cpu().numpy() - x_e).max(axis=1)
    dq = np.abs(logq.cpu().numpy() - lq_e)
    assert (dx < 5e-3).mean() >= 0.99 and dx.max() < 0.05
    assert (dq < 5e-3).mean() >= 0.99 and dq.max() < 0.05
    # density consistency: log_prob of the drawn points equals the log_q returned with them
    back = f.log_prob(x).cpu().numpy()
    assert np.abs(back - logq.cpu().numpy()).max() < 0.05
    # sample_and_log_prob through the Flow API (Philox base noise)
    xs, lqs = f.sample_and_log_prob(20_000)
    assert xs.shape == (20_000, dims) and bool(lqs.isfinite().all())


def test_flow_as_smc_proposal_importance(cuda):
    """C4 shape at reduced N: a fitted coupling flow as the proposal of the importance sampler
    (GMM likelihood); the evidence estimate is consistent with the analytic value."""
    import math

    import torch

    from aspire_b200 import targets as tg
    from aspire_b200.aspire import Aspire
    from aspire_b200.flows import CouplingFlow
    from aspire_b200.rng import PhiloxRNG

    dims = 32
    like = tg.GaussianDensity(0.5, 0.7, dims, device=cuda)
    prior = tg.GaussianDensity(0.0, 1.0, dims, device=cuda)
    var = 1.0 / (1.0 / 0.49 + 1.0)
    post = PhiloxRNG(1).normal(4000, dims, cuda) * math.sqrt(var) + var * 0.5 / 0.49
    torch.manual_seed(0)
    flow = CouplingFlow(dims, device=cuda, rng=PhiloxRNG(2))
    hist = flow.fit(post, n_epochs=30, lr=2e-3)
    assert hist.training_loss[-1] < hist.training_loss[0]
    a = Aspire(log_likelihood=like, log_prior=prior, dims=dims, flow=flow)
    s = a.sample_posterior(200_000, sampler="importance")
    flow._check_error()
    logz = dims * (-0.5 * math.log(2 * math.pi * 1.49) - 0.5 * 0.25 / 1.49)
    assert s.efficiency > 0.002
    assert s.log_evidence == pytest.approx(logz, abs=6 * s.log_evidence_error + 0.05)


def test_c4_tempering_step_with_the_flow_as_proposal(cuda):
    """BASELINE configs[3] at reduced N: 64-D, 8-component mixture likelihood, N(0, 4^2) prior, the coupling flow
    (bf16 tcgen05) as the proposal density inside the tempering step - reweight, beta search, exact resample,
    pCN moves whose accept test needs log q(x') from the flow kernel - against numpy: oracle/smc_numpy.py +
    oracle/pcn.py with oracle/flow.py (bf16-emulating) as q, the same numpy Generator call order.
    Stated bound (bf16 flow): log q agrees to 5e-3 for 99 % of the particles, so beta may differ by a few
    tolerances and an accept decision may flip where |log u - log a| is inside the flow error."""
    import torch

    from aspire_b200 import ops
    from aspire_b200 import targets as tg
    from aspire_b200.samplers import MiniPCNSMC
    from aspire_b200.samples import SMCSamples
    from oracle import densities as dn
    from oracle import pcn as opcn
    from oracle import problems
    from oracle import smc_numpy as sn

    dims, n, moves, rho, seed = 64, 20_000, 5, 0.3, 11
    f, p = _flow(dims, 3, cuda, seed=2)
    pg, po = tg.gmm_problem(dims, device=cuda), problems.gmm_problem(dims)
    u = np.random.default_rng(1).normal(size=(n, dims))
    x_d, lq_d = ops.flow_sample(f.struct(), _t(u, cuda))
    x = x_d.cpu().numpy()
    lq_e = of.forward(p, x, emulate_bf16=True)[1]
    dq = np.abs(lq_d.cpu().numpy() - lq_e)
    assert (dq < 5e-3).mean() >= 0.99 and dq.max() < 0.05
    ll, lp = dn.evaluate(po["like"], x), dn.evaluate(po["prior"], x)
    np.testing.assert_allclose(pg["like"].log_prob(x_d).cpu().numpy(), ll, rtol=1e-12, atol=1e-10)

    s = MiniPCNSMC(pg["like"], pg["prior"], dims, f, rng=np.random.default_rng(seed))
    s._configure_kernel(dict(n_steps=moves, rho=rho, adapt_rho=False, step_fn="pcn"))
    s.target_efficiency, s.target_efficiency_rate, s.adaptive, s.adaptive_min_beta_step = 0.5, 1.0, True, False
    s.max_beta_step, s.resampling = 1.0, "multinomial"
    from aspire_b200.history import SMCHistory

    s.history = SMCHistory()
    start = SMCSamples(x_d, log_likelihood=pg["like"].log_prob(x_d), log_prior=pg["prior"].log_prob(x_d), log_q=lq_d,
                       beta=0.0, device=cuda)
    assert not s.fused  # the proposal is a flow: semi-fused moves (asp_pcn_propose_eval + flow + asp_pcn_accept_scalar)
    out, beta, _ = s.tempering_step(start, 0.0, math.nan, 0.0, 1e-6)
    f.check_error()

    # ---- numpy: the same step with the bf16-emulating flow spec, same generator call order
    rng = np.random.default_rng(seed)
    beta_ref, _ = sn.determine_beta(ll, lp, lq_e, 0.0, beta_tolerance=1e-6)
    assert abs(beta - beta_ref) <= 4e-6                      # a few tolerances: the last decisions see log q +- 5e-3
    rec, _ = sn.smc_step_reweight(ll, lp, lq_e, 0.0, beta_tolerance=1e-6)
    assert s.history.ess[-1] == pytest.approx(rec.ess, rel=2e-2)
    assert s.history.log_norm_ratio[-1] == pytest.approx(rec.log_norm_ratio, abs=2e-2)
    # resample with the GPU's beta and log q so that the same ancestors are compared move by move
    lq_g = lq_d.cpu().numpy()
    pw = sn.resample_weights(ll, lp, lq_g, 0.0, beta)
    idx = sn.choice_indices(pw, rng.random(n))
    xi = rng.standard_normal((moves, n, dims))
    log_u = np.log(1.0 - rng.random((moves, n)))
    mu, sigma = opcn.weighted_moments(x[idx], np.ones(n))
    q = dict(kind="callable", fn=lambda y: of.forward(p, y, emulate_bf16=True)[1])
    ref = opcn.sweep(x[idx], ll[idx], lp[idx], lq_g[idx], like=po["like"], prior=po["prior"], q=q, beta=beta, mu=mu,
                     sigma=sigma, rho=rho, n_moves=moves, xi=xi, log_u=log_u)
    got = out.x.cpu().numpy()
    same = np.isclose(got, ref[0], rtol=1e-9, atol=1e-9).all(axis=1)
    assert same.mean() > 0.97, same.mean()                   # flips only where |log u - log a| < flow error
    acc_ref = ref[4].sum() / (moves * n)
    assert s.history.mcmc_acceptance[-1] == pytest.approx(acc_ref, abs=5e-3)
    np.testing.assert_allclose(out.log_likelihood.cpu().numpy()[same], ref[1][same], rtol=1e-9, atol=1e-9)
    dlq = np.abs(out.log_q.cpu().numpy()[same] - ref[3][same])
    assert (dlq < 5e-3).mean() >= 0.99 and dlq.max() < 0.05


@pytest.mark.parametrize("dims", [4, 11, 20, 48])
def test_any_dimension_runs_on_padded_kernels(cuda, dims):
    """2 <= D <= 64 on the 16 / 32 / 64-dimensional tcgen05 kernels with inert padding coordinates: log_prob equals
    the torch restatement of the same (fitted) flow, samples come back in D dimensions with the log density
    log_prob gives them, and the density is normalised (importance estimate of its integral)."""
    import torch

    from aspire_b200.flows import CouplingFlow
    from aspire_b200.rng import PhiloxRNG

    rng = np.random.default_rng(dims)
    x = rng.normal(size=(3000, dims)) * np.linspace(0.6, 1.8, dims) + 0.5
    x[:, -1] += 0.6 * x[:, 0]
    torch.manual_seed(0)
    f = CouplingFlow(dims, device=cuda, rng=PhiloxRNG(4))
    f.fit(_t(x, cuda), n_epochs=6, batch_size=500, lr=3e-3)
    lq = f.log_prob(_t(x, cuda))
    f.check_error()
    ref = f._torch_forward(f._embed(_t(x, cuda)).float(), f.layers).double() + f._pad_const
    err = (lq - ref).abs()
    assert float(err.max()) < 0.15 and float(err.mean()) < 0.03           # the stated bf16 bound
    xs, lqs = f.sample_and_log_prob(50_000)
    assert xs.shape == (50_000, dims) and bool(lqs.isfinite().all())
    back = f.log_prob(xs)
    assert float((back - lqs).abs().max()) < 0.05
    z, lq2 = f.forward(_t(x[:100], cuda))
    assert z.shape == (100, dims) and float((lq2 - lq[:100]).abs().max()) == 0.0
    # normalisation: E_q[N(x; m, s) / q(x)] = 1 for any normalised density N
    m, s = torch.as_tensor(x.mean(0), device=cuda), torch.as_tensor(x.std(0) * 0.8, device=cuda)
    log_n = (-0.5 * ((xs - m) / s) ** 2 - torch.log(s) - 0.5 * math.log(2 * math.pi)).sum(dim=1)
    w = torch.exp(log_n - lqs)
    assert float(w.mean()) == pytest.approx(1.0, abs=6 * float(w.std()) / math.sqrt(len(w)) + 0.02)


def test_quickstart_with_a_learned_coupling_flow(cuda):
    """The reference's README quickstart shape (4-D, README.md:42-77) with a LEARNED flow as the proposal, like the
    reference's default: fit the coupling flow to initial samples, then SMC; posterior N(0, 1/2 I), log Z = 2 ln pi."""
    import torch

    from aspire_b200 import targets as tg
    from aspire_b200.aspire import Aspire
    from aspire_b200.rng import PhiloxRNG

    dims = 4
    like = tg.GaussianDensity(0.0, 1.0, dims, normalised=False, device=cuda)
    prior = tg.GaussianDensity(0.0, 1.0, dims, normalised=False, device=cuda)
    torch.manual_seed(1)
    a = Aspire(log_likelihood=like, log_prior=prior, dims=dims, flow_class="coupling", device=cuda)
    init = PhiloxRNG(3).normal(2000, dims, cuda) * 1.2
    a.fit(init, n_epochs=20)
    s = a.sample_posterior(20_000, sampler="smc", sampler_kwargs=dict(n_steps=20), rng=5)
    x = s.x.cpu().numpy()
    assert np.abs(x.mean(axis=0)).max() < 0.05
    np.testing.assert_allclose(x.var(axis=0), 0.5, rtol=0.08)
    assert s.log_evidence == pytest.approx(2 * math.log(math.pi), abs=0.1)


def _close_up_to_relu_gates(got, ref, rtol, atol):
    """Element-wise agreement except for the few entries a ReLU gate can move: a pre-activation within rounding of
    zero is open in one fp32 summation order and shut in the other, which changes one sample's contribution to one
    row of a weight gradient (seen at D = 64: 4 of 31,296 entries).  At most 0.1 % of the entries may miss the
    element-wise bound, and the tensors must agree to ``rtol`` in norm."""
    import torch

    bad = (got - ref).abs() > atol + rtol * ref.abs()
    assert int(bad.sum()) <= max(1, got.numel() // 1000), (int(bad.sum()), got.numel())
    assert float((got - ref).norm()) <= 5 * rtol * float(ref.norm()) + atol


@pytest.mark.parametrize("dims", [16, 5, 64])
def test_native_fit_gradients_and_updates_match_torch_autograd(cuda, dims):
    """asp_flow_fit_step (forward, backward, Adam kernels over the fp32 master weights) against torch autograd +
    torch.optim.Adam on the same weights and the same mini-batches: the loss before every update, the gradient of
    the first step and the weights after 12 steps agree; padding rows stay exactly zero."""
    import torch

    from aspire_b200.flows import CouplingFlow, _NativeTrainer

    rng = np.random.default_rng(dims)
    x = rng.normal(size=(1500, dims)) * np.linspace(0.7, 2.0, dims) - 0.3
    x[:, -1] += 0.7 * x[:, 0] ** 2
    f = CouplingFlow(dims, device=cuda, n_layers=3, seed=7)
    xt = f._embed(_t(x, cuda))
    f.mean, f.std = xt.mean(dim=0), xt.std(dim=0).clamp_min(1e-12)
    if f._pad:
        f.mean = torch.where(f._real, f.mean, torch.zeros_like(f.mean))
        f.std = torch.where(f._real, f.std, torch.ones_like(f.std))
    for layer in f.layers:  # leave the tiny-W3 initialisation: every path of the backward pass must carry signal
        layer["W3"].mul_(30.0)
        layer["b3"].add_(0.05)
    f._hold_padding()
    x32 = xt.float()
    ref_layers = [{k: v.detach().clone().requires_grad_(True) for k, v in layer.items()} for layer in f.layers]
    w0 = [{k: v.detach().clone() for k, v in layer.items()} for layer in f.layers]
    params = [p for layer in ref_layers for p in layer.values()]
    opt = torch.optim.Adam(params, lr=2e-3)
    tr = _NativeTrainer(f, 500)
    for step in range(12):
        batch = x32[(step % 3) * 500:(step % 3) * 500 + 500]
        loss_ref = -f._torch_forward(batch, ref_layers).mean()
        opt.zero_grad()
        loss_ref.backward()
        if step == 0:
            g_ref = torch.cat([torch.cat([layer[n].grad.reshape(-1) for n in ("W1", "b1", "W2", "b2", "W3", "b3")])
                               for layer in ref_layers])
            if f._pad:  # torch sees gradients on the padding rows; the optimiser side holds them at zero
                masks = [f._out_mask(t).float() for t in range(3)]
        clip = 5.0 if step >= 6 else None
        f._mask_padding_grads(ref_layers)
        if clip:
            torch.nn.utils.clip_grad_norm_(params, clip)
        opt.step()
        with torch.no_grad():
            if f._pad:
                for t, layer in enumerate(ref_layers):
                    m = f._out_mask(t).float()
                    layer["W3"].mul_(m[:, None])
                    layer["b3"].mul_(m)
        loss = tr.step(batch, 2e-3, clip)
        assert float(loss) == pytest.approx(float(loss_ref.detach()), rel=1e-3, abs=1e-3), step
        if step == 0:
            got, ref = tr.unflatten(tr.g.clone()), tr.unflatten(g_ref)
            for t in range(3):
                for n in ("W1", "b1", "W2", "b2", "W3", "b3"):
                    r = ref[t][n]
                    if f._pad and n in ("W3", "b3"):  # compare on the real rows only
                        r = r * (masks[t][:, None] if n == "W3" else masks[t])
                    _close_up_to_relu_gates(got[t][n], r, rtol=2e-3, atol=2e-5)
    got = tr.unflatten(tr.w)
    for t in range(3):
        for n in ("W1", "b1", "W2", "b2", "W3", "b3"):
            # Adam normalises every element's step to ~lr, so the comparison is on the update as a whole
            upd, upd_ref = got[t][n] - w0[t][n], ref_layers[t][n].detach() - w0[t][n]
            assert float((upd - upd_ref).norm()) <= 0.05 * float(upd_ref.norm()) + 1e-6, (t, n)
        if f._pad:
            dead = ~f._out_mask(t)
            assert float(got[t]["W3"][dead].abs().sum()) == 0.0 and float(got[t]["b3"][dead].abs().sum()) == 0.0
    # validation loss entry point == forward of the current weights
    f.layers = got
    val = float(tr.loss(x32[:400]))
    assert val == pytest.approx(float(-f._torch_forward(x32[:400], got).mean()), rel=2e-4)


def test_native_fit_trains_and_follows_the_protocol(cuda):
    """CouplingFlow.fit on the device (native kernels by default): history, best-validation restore, and the fitted
    flow is a better proposal than the untrained one; the torch route (native=False) lands on a comparable loss."""
    import torch

    from aspire_b200.flows import CouplingFlow
    from aspire_b200.rng import PhiloxRNG

    rng = np.random.default_rng(0)
    z = rng.normal(size=(6000, 16))
    z[:, 8:] = 0.9 * z[:, :8] + 0.3 * z[:, 8:]
    x = z * np.linspace(0.5, 3.0, 16) + 2.0
    a = CouplingFlow(16, device=cuda, rng=PhiloxRNG(1), seed=3)
    ha = a.fit(_t(x, cuda), n_epochs=15, batch_size=250, lr=2e-3, lr_annealing=True, clip_grad=10.0)
    b = CouplingFlow(16, device=cuda, rng=PhiloxRNG(1), seed=3)
    hb = b.fit(_t(x, cuda), n_epochs=15, batch_size=250, lr=2e-3, lr_annealing=True, clip_grad=10.0, native=False)
    assert len(ha.training_loss) == len(ha.validation_loss) == 15
    assert ha.training_loss[-1] < ha.training_loss[0] - 0.5
    assert min(ha.validation_loss) == pytest.approx(min(hb.validation_loss), abs=0.05)
    np.testing.assert_allclose(ha.training_loss[:3], hb.training_loss[:3], rtol=2e-3)
    lq = a.log_prob(_t(x[:2000], cuda))
    a.check_error()
    # the tcgen05 evaluation of the re-packed bf16 weights reproduces the trained fp32 model (these rows are mostly
    # training rows, so the value sits between the last training loss and the validation loss)
    assert ha.training_loss[-1] - 0.3 < float(-lq.mean()) < min(ha.validation_loss) + 0.1


# ---------------------------------------------------------------- the sweep with the flow inside (csrc/flow_move.cu)
@pytest.mark.parametrize("dims,like,n,n_comp,philox", [(64, "gmm", 6000, 8, False), (64, "gmm", 4097, 8, True),
                                                       (32, "gauss", 5000, 0, False), (16, "gmm", 3001, 11, True),
                                                       (32, "gmm", 130, 3, True), (64, "gauss", 1, 0, True)])
def test_flow_sweep_equals_the_semi_fused_moves(cuda, dims, like, n, n_comp, philox):
    """asp_pcn_sweep_flow (gather + all moves, log q(x') on tcgen05 inside the move, one kernel) against the per-move
    chain asp_pcn_propose_eval -> asp_flow_logprob_bf16 -> asp_pcn_accept_scalar on the same ancestors and the same
    noise (supplied arrays or the Philox streams): the same trajectory.  The two routes evaluate the same fp32 flow
    chain; the proposal's fp64 rounding may differ in the last bit, which can move log q by ~1e-6 and flip an accept
    decision that sits inside that margin - allowed for at most 0.1 % of the particles."""
    import torch

    from aspire_b200 import ops
    from aspire_b200 import targets as tg

    rng = np.random.default_rng(dims + n)
    f, _ = _flow(dims, 3, cuda, seed=4)
    f.mean = _t(rng.normal(0.0, 0.3, dims), cuda)
    f.std = _t(rng.uniform(1.5, 2.5, dims), cuda)
    f._struct = None
    prior = tg.GaussianDensity(0.0, 4.0, dims, device=cuda)
    if like == "gmm":
        means = rng.choice([-2.0, 2.0], size=(n_comp, dims))
        L = tg.GaussianMixtureDensity(means, 0.7, device=cuda)
    else:
        L = tg.GaussianDensity(rng.normal(0.5, 0.2, dims), rng.uniform(0.8, 1.2, dims), dims, device=cuda)
    moves, beta, rho = 6, 0.35, 0.3
    x = _t(rng.normal(0.0, 2.0, (n, dims)), cuda)
    ll, lp, lq = L.log_prob(x), prior.log_prob(x), f.log_prob(x)
    m = n + 17
    idx = _t(np.sort(rng.integers(0, n, size=m)), cuda)
    mu, sigma = _t(rng.normal(0.0, 0.2, dims), cuda), _t(rng.uniform(1.5, 2.2, dims), cuda)
    if philox:
        mk = lambda: ops.make_pcn_params(beta, rho, mu, sigma, moves, seed=17, stream0=3, index0=500)  # noqa: E731
    else:
        xi = _t(rng.standard_normal((moves, m, dims)), cuda)
        log_u = _t(np.log(1.0 - rng.random((moves, m))), cuda)
        mk = lambda: ops.make_pcn_params(beta, rho, mu, sigma, moves, xi=xi, log_u=log_u)  # noqa: E731
    assert ops.pcn_sweep_flow_supported(dims, L, prior)
    out = ops.pcn_sweep_flow(x, ll, lp, lq, idx, L, prior, f.struct(), mk())
    f.check_error()
    prm = mk()
    st = list(ops.gather_rows(x, idx, ll, lp, lq))
    acc = torch.zeros(moves, dtype=torch.int64, device=cuda)
    for t in range(moves):
        xp, ll2, lp2, dr = ops.pcn_propose_eval(st[0], L, prior, prm, t)
        ops.pcn_accept_scalar(st[0], st[1], st[2], st[3], xp, ll2, lp2, f.log_prob(xp), dr, prm, t, acc)
    f.check_error()
    got, ref = out[0].cpu().numpy(), st[0].cpu().numpy()
    same = np.isclose(got, ref, rtol=1e-12, atol=1e-12).all(axis=1)
    assert same.mean() >= 0.999, same.mean()
    assert abs(int(out[4].sum()) - int(acc.sum())) <= max(2, moves * m // 1000)
    if not philox:
        np.testing.assert_array_equal(out[4].cpu().numpy()[:1], acc.cpu().numpy()[:1])  # move 0 starts from equal bits
    for k in (1, 2):
        np.testing.assert_allclose(out[k].cpu().numpy()[same], st[k].cpu().numpy()[same], rtol=1e-11, atol=1e-10)
    np.testing.assert_allclose(out[3].cpu().numpy()[same], st[3].cpu().numpy()[same], rtol=0, atol=2e-5)
    rate = int(acc.sum()) / (moves * m)
    assert 0.02 < rate < 0.98 or n < 10


def test_flow_sweep_is_the_route_of_the_c4_mutation(cuda):
    """A sampler with built-in densities and an unpadded CouplingFlow proposal mutates through asp_pcn_sweep_flow; the
    ``flow_sweep=False`` kernel option keeps the per-move chain, and both give the same ensemble."""
    import torch

    from aspire_b200 import _lib
    from aspire_b200 import targets as tg
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import MiniPCNSMC
    from aspire_b200.samples import SMCSamples

    dims, n = 64, 9000
    f, _ = _flow(dims, 3, cuda, seed=5)
    f.std = torch.full((dims,), 2.5, dtype=torch.float64, device=cuda)
    f._struct = None
    pg = tg.gmm_problem(dims, device=cuda)
    x0, lq0 = f.sample_and_log_prob(n)
    res = []
    for flag in (True, False):
        s = MiniPCNSMC(pg["like"], pg["prior"], dims, f, rng=PhiloxRNG(9))
        s._configure_kernel(dict(n_steps=4, rho=0.3, adapt_rho=False, flow_sweep=flag, step_fn="pcn"))
        from aspire_b200.history import SMCHistory

        s.history = SMCHistory()
        p0 = SMCSamples(x0, log_likelihood=pg["like"].log_prob(x0), log_prior=pg["prior"].log_prob(x0), log_q=lq0,
                        beta=0.0, device=cuda)
        assert s._flow_sweep_ok(p0, 4, False) is flag
        before = _lib.launch_count
        out = s.mutate(p0, 0.2)
        res.append((out, s.history.mcmc_acceptance[-1], _lib.launch_count - before))
    (a, acc_a, launches_a), (b, acc_b, launches_b) = res
    assert launches_a < launches_b
    same = torch.isclose(a.x, b.x, rtol=1e-12, atol=1e-12).all(dim=1)
    assert float(same.double().mean()) >= 0.999
    assert acc_a == pytest.approx(acc_b, abs=2e-3)
    f.check_error()


def test_flow_sweep_is_deterministic_and_independent_of_the_split(cuda):
    """Size-independent properties of asp_pcn_sweep_flow at 2e6 x 64-D (C4 shape): two launches give identical bits;
    sweeping the two halves of the slot range separately (``index0`` = first slot of the half: the noise is keyed by the
    global output slot) gives the same rows as sweeping them together; a particle that never moved keeps its row
    and logs bit for bit."""
    import torch

    from aspire_b200 import ops
    from aspire_b200 import targets as tg

    dims, n, moves = 64, 2_000_000, 3
    f, _ = _flow(dims, 3, cuda, seed=6)
    f.std = torch.full((dims,), 2.5, dtype=torch.float64, device=cuda)
    f._struct = None
    pg = tg.gmm_problem(dims, device=cuda)
    x, lq = f.sample_and_log_prob(n)
    ll, lp = pg["like"].log_prob(x), pg["prior"].log_prob(x)
    idx = torch.sort((ops.philox_uniform(6, 0, n, cuda) * n).to(torch.int64).clamp_(0, n - 1)).values
    mu, sigma = ops.moments_to_mu_sigma(ops.weighted_moments(x[: 1 << 18]), dims)

    def sweep(lo, hi):
        prm = ops.make_pcn_params(0.02, 0.3, mu, sigma, moves, seed=5, stream0=7, index0=lo)
        return ops.pcn_sweep_flow(x, ll, lp, lq, idx[lo:hi].contiguous(), pg["like"], pg["prior"], f.struct(), prm)

    a, b = sweep(0, n), sweep(0, n)
    for u, v in zip(a[:4], b[:4]):
        assert torch.equal(u, v)
    h = n // 2 + 77  # an uneven split that is not a multiple of the 128-particle tile
    c, d = sweep(0, h), sweep(h, n)
    for k in range(4):
        assert torch.equal(a[k][:h], c[k]) and torch.equal(a[k][h:], d[k])
    assert int(a[4].sum()) == int(c[4].sum()) + int(d[4].sum())
    f.check_error()
    stay = (a[0] == x[idx]).all(dim=1)
    assert 0.0 < float(stay.double().mean()) < 0.9          # some never moved (acceptance < 1), most did
    for k, src in ((1, ll), (2, lp), (3, lq)):
        assert torch.equal(a[k][stay], src[idx][stay])
    rate = int(a[4].sum()) / (moves * n)
    assert 0.3 < rate < 0.99


def test_flow_proposal_example_runs_through_the_flow_sweep(cuda, capsys):
    """examples/flow_proposal.py: Aspire.fit on the device, then SMC with the learned flow as proposal on a bimodal
    16-D target; every mutation goes through asp_pcn_sweep_flow (checked on the call counter), evidence and mode
    masses come out right."""
    import importlib.util
    import os

    from aspire_b200 import _lib

    spec = importlib.util.spec_from_file_location("flow_proposal", os.path.join(os.path.dirname(__file__), "..", "examples",
                                                                             "flow_proposal.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    calls = []
    orig = _lib.check

    def counting(rc, what=""):
        calls.append(what)
        return orig(rc, what)

    from aspire_b200 import ops

    ops.check, old = counting, ops.check
    try:
        posterior, log_z, left = mod.main(60_000)
    finally:
        ops.check = old
    assert "asp_pcn_sweep_flow" in calls and "asp_pcn_propose_eval" not in calls and "asp_pcn_propose" not in calls
    assert posterior.log_evidence == pytest.approx(log_z, abs=5 * posterior.log_evidence_error + 0.15)
    assert 0.4 < left < 0.6
    assert "tempering steps" in capsys.readouterr().out


def test_flow_sweep_with_the_independence_proposal(cuda):
    """rho = 1 (where the step-size adaptation saturates): x' = mu + sigma xi does not depend on the current point;
    the flow sweep takes it (no division by sqrt(1 - rho^2) anywhere) and still equals the per-move chain."""
    import torch

    from aspire_b200 import ops
    from aspire_b200 import targets as tg

    dims, n, moves = 32, 4000, 4
    rng = np.random.default_rng(3)
    f, _ = _flow(dims, 3, cuda, seed=8)
    f.std = torch.full((dims,), 1.8, dtype=torch.float64, device=cuda)
    f._struct = None
    L = tg.GaussianDensity(rng.normal(0.3, 0.2, dims), 1.1, dims, device=cuda)
    P = tg.GaussianDensity(0.0, 4.0, dims, device=cuda)
    x = _t(rng.normal(0.0, 1.5, (n, dims)), cuda)
    ll, lp, lq = L.log_prob(x), P.log_prob(x), f.log_prob(x)
    mu, sigma = _t(np.full(dims, 0.2), cuda), _t(np.full(dims, 1.3), cuda)
    mk = lambda: ops.make_pcn_params(0.6, 1.0, mu, sigma, moves, seed=2, stream0=9)  # noqa: E731
    out = ops.pcn_sweep_flow(x, ll, lp, lq, None, L, P, f.struct(), mk())
    prm = mk()
    st = [x.clone(), ll.clone(), lp.clone(), lq.clone()]
    acc = torch.zeros(moves, dtype=torch.int64, device=cuda)
    for t in range(moves):
        xp, ll2, lp2, dr = ops.pcn_propose_eval(st[0], L, P, prm, t)
        ops.pcn_accept_scalar(st[0], st[1], st[2], st[3], xp, ll2, lp2, f.log_prob(xp), dr, prm, t, acc)
    f.check_error()
    same = torch.isclose(out[0], st[0], rtol=1e-12, atol=1e-12).all(dim=1)
    assert float(same.double().mean()) >= 0.998 and abs(int(out[4].sum()) - int(acc.sum())) <= 8
    assert 0.01 < int(acc.sum()) / (moves * n) < 0.9
