#!/usr/bin/env python
"""Micro-benchmarks of BASELINE.json configs[1], [3], [4] (bench.py carries the headline, configs[2]).

    python benchmarks/micro.py flow        # C4: coupling-flow log_prob, N = 1e7, D = 64 (tcgen05)
    python benchmarks/micro.py resample    # C5: exact CDF + ancestor search, N = 1e4 .. 1e9
    python benchmarks/micro.py importance  # C2: importance reweighting, N = 1e6, D = 16
    python benchmarks/micro.py sweep       # K8: fused gather + 10 pCN moves per benchmark problem
    python benchmarks/micro.py full        # whole sample_posterior(sampler="smc") run through the drop-in API
    python benchmarks/micro.py c4          # C4: mixture SMC step with the tcgen05 flow as proposal (unfused moves)
    python benchmarks/micro.py transform   # 8f-4: preconditioning transforms, N = 2e7, D = 32 (HBM-bound pass)
    python benchmarks/micro.py precond     # 8f-4: SMC step with bounded parameters (logit + affine preconditioning)

Each prints JSON lines; times are CUDA events over `--iters` launches after 3 warm-ups, on inputs
larger than L2 where the size allows (noted per line otherwise).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from aspire_b200 import ops  # noqa: E402
from aspire_b200 import targets as tg  # noqa: E402

PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def flow(args):
    from aspire_b200.flows import CouplingFlow

    dev = torch.device("cuda")
    for dims in (64, 32):
        n = args.n or 10_000_000
        f = CouplingFlow(dims, device=dev, n_layers=3)
        x = ops.philox_normal(1, 0, n, dims, dev)
        ms = timed(lambda: ops.flow_logprob(f.struct(), x), args.iters)
        f._check_error()
        half, h = dims // 2, 64
        flops = 3 * 2 * (half * h + h * h + h * dims) * n
        bytes_ = n * (8 * dims + 8)
        print(json.dumps({"bench": "flow_logprob_bf16", "config": "C4" if dims == 64 else "C4 (D=32)", "particles": n,
                          "dims": dims, "ms": ms, "particles_per_s": n / ms * 1e3, "tflops": flops / ms / 1e9,
                          "tensor_frac_of_measured_bf16": flops / ms / 1e9 / PEAKS["bf16_tflops"],
                          "hbm_gbs": bytes_ / ms / 1e6, "hbm_frac": bytes_ / ms / 1e6 / PEAKS["hbm_gbs"]}))
        u = ops.philox_normal(2, 0, n, dims, dev)
        ms = timed(lambda: ops.flow_sample(f.struct(), u), args.iters)
        print(json.dumps({"bench": "flow_sample_bf16", "particles": n, "dims": dims, "ms": ms,
                          "particles_per_s": n / ms * 1e3, "tflops": flops / ms / 1e9}))


def _numpy_resample(n, sigma=2.0):
    """The reference's resampling on the host cores: numpy ``Generator.choice`` = cumsum + normalise + searchsorted
    of n iid uniforms (src/aspire/samples.py:1251-1287), one core (numpy's sort / search are single-threaded)."""
    import time

    import numpy as np

    rng = np.random.default_rng(7)
    lw = rng.normal(0.0, sigma, n)
    p = np.exp(lw - lw.max())
    p /= p.sum()
    t0 = time.perf_counter()
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    idx = cdf.searchsorted(rng.random(n), side="right")
    t_mult = time.perf_counter() - t0
    t0 = time.perf_counter()
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    idx = cdf.searchsorted((np.arange(n) + 0.37) / n, side="right")
    t_sys = time.perf_counter() - t0
    del idx
    return t_mult * 1e3, t_sys * 1e3


def resample(args):
    """C5 (BASELINE.json configs[4]): exact CDF + ancestors for systematic and multinomial resampling, 1e4 .. 1e9
    particles, log-weights ~ N(0, 2^2), beside numpy on the host cores (n <= --numpy-max).  Under torchrun
    (``--gpus N``) the particles are sharded: exact CDF with the carry chain over the ranks, owner-computes
    ancestors at sorted uniforms, device-timed, max over ranks.  Multinomial is timed in both modes: sorted uniforms
    generated on the device (throughput mode) and supplied iid uniforms through the guide table (parity mode)."""
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return _resample_sharded(args, world)
    dev = torch.device("cuda")
    sizes = [args.n] if args.n else [10**4, 10**5, 10**6, 10**7, 10**8, 10**9]
    for n in sizes:
        lw = ops.philox_normal(7, 1, n, 1, dev).reshape(-1) * 2.0
        p = torch.exp(lw - lw.max())
        p /= p.sum()
        del lw
        iters = args.iters if n <= 10**8 else 3
        t_cdf = timed(lambda: ops.cdf_exact(p), iters)
        cdf = ops.cdf_exact(p)
        seq = [0]

        def sorted_mult():
            seq[0] += 1
            return ops.search_multinomial_sorted(cdf, 7, seq[0], n)

        t_sorted = timed(sorted_mult, iters)
        t_sys = timed(lambda: ops.search_systematic(cdf, 0.37, n), iters)
        u = ops.philox_uniform(7, 0, n, dev)
        t_iid = timed(lambda: ops.search_multinomial(cdf, u), iters)
        del u
        np_mult = np_sys = None
        if n <= args.numpy_max:
            np_mult, np_sys = _numpy_resample(n)
        for name, t, t_np in (("multinomial_sorted_uniforms", t_cdf + t_sorted, np_mult),
                              ("multinomial_supplied_iid_uniforms", t_cdf + t_iid, np_mult),
                              ("systematic", t_cdf + t_sys, np_sys)):
            print(json.dumps({"bench": f"resample_{name}", "config": "C5", "n_gpus": 1, "particles": n, "ms_cdf": t_cdf,
                              "ms_total": t, "indices_per_s": n / t * 1e3, "algorithmic_gbs": 32 * n / t / 1e6,
                              "hbm_frac": 32 * n / t / 1e6 / PEAKS["hbm_gbs"],
                              "numpy_host_ms": t_np, "numpy_indices_per_s": None if t_np is None else n / t_np * 1e3,
                              "numpy_cores": 1, "vs_numpy": None if t_np is None else t_np / t,
                              "note": "inputs L2-resident" if 24 * n < 126e6 else "inputs exceed L2"}))
        del p, cdf


def _resample_sharded(args, world):
    import os

    import torch.distributed as dist

    from aspire_b200.distributed import Comm

    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    sizes = [args.n] if args.n else [10**5, 10**6, 10**7, 10**8, 10**9]
    for line in resample_sharded_lines(comm, dev, sizes, args.iters):
        print(json.dumps(line))
    dist.destroy_process_group()


def resample_sharded_lines(comm, dev, sizes, iters_small, peer=False):
    """The sharded C5 measurement on an existing communicator (every rank calls it; rank 0 gets the lines).
    ``peer``: the engine's mailboxes (False: create them here)."""
    from aspire_b200.distributed import PeerMailboxes, ShardedSMC

    world, r = comm.world, comm.rank
    eng = ShardedSMC(comm, None, None, None, seed=7)
    eng._peer = PeerMailboxes.create(comm, dev) if peer is False else peer
    lines = []
    for n in sizes:
        base, rem = divmod(n, world)
        n_local, start = base + (1 if r < rem else 0), r * base + min(r, rem)
        lw = ops.philox_normal(7, 1, n_local, 1, dev, index0=start).reshape(-1) * 2.0
        mx = lw.max().reshape(1)
        comm.all_reduce_max(mx)
        p = torch.exp(lw - mx)
        tot = p.sum().reshape(1)
        comm.all_reduce_sum(tot)
        p /= tot
        del lw

        def cdf_only():
            info = torch.stack([p.sum(), p.max(), p.sum() * 0, p.sum() * 0 + n_local]).contiguous()
            return eng.global_cdf(p, None, info, n)

        def multinomial():
            cdf, chain, _ = cdf_only()
            return eng.sorted_ancestors(cdf, chain, n)

        def systematic():
            cdf, chain, _ = cdf_only()
            lo, hi = eng.slice_bounds(chain, cdf, r, world)
            # slots whose position (j + u0) / n falls into this rank's CDF slice (lo, hi]
            j0 = 0 if r == 0 else int(torch.ceil(torch.tensor(lo * n - 0.37)).item())
            j1 = n if r == world - 1 else int(torch.ceil(torch.tensor(hi * n - 0.37)).item())
            j0, j1 = max(min(j0, n), 0), max(min(j1, n), 0)
            return ops.search_systematic(cdf, 0.37, max(j1 - j0, 0), j0=j0, m_total=n)

        iters = iters_small if n <= 10**8 else 3
        res = {}
        for name, fn in (("cdf", cdf_only), ("multinomial_sorted_uniforms", multinomial), ("systematic", systematic)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            comm.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device=dev)
            comm.all_reduce_max(t)
            res[name] = float(t[0])
        if r == 0:
            for name in ("multinomial_sorted_uniforms", "systematic"):
                t = res[name]
                lines.append({"bench": f"resample_{name}", "config": "C5", "n_gpus": world, "particles": n,
                              "ms_cdf": res["cdf"], "ms_total": t, "indices_per_s": n / t * 1e3,
                              "algorithmic_gbs_per_gpu": 32 * n / world / t / 1e6,
                              "hbm_frac": 32 * n / world / t / 1e6 / PEAKS["hbm_gbs"],
                              "note": "sharded: exact CDF carry chain over " + ("NVLink peer memory" if eng._peer else "NCCL")
                                      + ", owner-computes ancestors, device-timed, max over ranks"})
        del p
    return lines


def importance(args):
    """C2: 16-D Gaussian posterior, importance reweighting of 1e6 flow-proposed particles: the whole
    ``ImportanceSampler.sample`` (flow sample + log q on tcgen05, prior / likelihood kernels, weights and
    evidence) and its reweighting half alone."""
    from aspire_b200.flows import CouplingFlow
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import ImportanceSampler

    dev = torch.device("cuda")
    n, dims = args.n or 1_000_000, 16
    prob = tg.gaussian_problem(dims, device=dev)
    x = ops.philox_normal(3, 0, n, dims, dev, mean=prob["q"].mean, sigma=prob["q"].sigma)

    def step():
        lq = prob["q"].log_prob(x)
        lp = prob["prior"].log_prob(x)
        ll = prob["like"].log_prob(x)
        return ops.importance_weights(ll, lp, lq)

    ms = timed(step, args.iters)
    print(json.dumps({"bench": "importance_reweight", "config": "C2", "particles": n, "dims": dims, "ms": ms,
                      "particles_per_s": n / ms * 1e3, "algorithmic_gbs": (16 * dims + 64) * n / ms / 1e6,
                      "note": "launch-latency dominated at this size (6 kernels, 0.3 GB)"}))
    flow = CouplingFlow(dims, device=dev, n_layers=3, rng=PhiloxRNG(5), seed=2)
    flow.mean = torch.full((dims,), 0.3, dtype=torch.float64, device=dev)
    flow.std = torch.full((dims,), 1.2, dtype=torch.float64, device=dev)
    flow._struct = None
    sampler = ImportanceSampler(prob["like"], prob["prior"], dims, flow)
    res = {}

    def full():
        res["s"] = sampler.sample(n)

    ms = timed(full, args.iters)
    flow._check_error()
    s = res["s"]
    print(json.dumps({"bench": "importance_sampler_flow_proposal", "config": "C2", "particles": n, "dims": dims,
                      "ms": ms, "particles_per_s": n / ms * 1e3, "log_evidence": float(s.log_evidence),
                      "efficiency": float(s.effective_sample_size) / n,
                      "note": "ImportanceSampler.sample: Philox normals -> coupling flow sample + log q (tcgen05) -> "
                              "prior, likelihood -> weights / evidence / ESS"}))


def sweep(args):
    """K8 alone: fused resampling gather + 10 pCN moves (Philox noise), per benchmark problem.
    --dump PATH also saves the result of a small run (A/B checks between kernel variants)."""
    import numpy as np

    dev = torch.device("cuda")
    cases = [("rosenbrock", 32, 20_000_000), ("gmm", 64, 4_000_000), ("gauss", 16, 20_000_000)]
    mk = dict(rosenbrock=tg.rosenbrock_problem, gmm=tg.gmm_problem, gauss=tg.gaussian_problem)
    dump = {}
    for name, dims, n in cases:
        if args.only and args.only != name:
            continue
        n = args.n or n
        dims = args.dims or dims
        pg = mk[name](dims, device=dev)
        x = ops.philox_normal(5, 0, n, dims, dev, mean=pg["q"].mean, sigma=pg["q"].sigma)
        lq, lp, ll = pg["q"].log_prob(x), pg["prior"].log_prob(x), pg["like"].log_prob(x)
        idx = (ops.philox_uniform(6, 0, n, dev) * n).to(torch.int64).clamp_(0, n - 1)
        if args.sorted_idx:  # what the sorted-uniform resampler hands to the sweep
            idx = torch.sort(idx).values
        mom = ops.weighted_moments(x[: 1 << 20])
        mu, sigma = ops.moments_to_mu_sigma(mom, dims)
        prm = ops.make_pcn_params(0.1, 0.3, mu, sigma, args.moves, seed=99, stream0=3, total_accepts_only=True,
                                  nu=args.nu)  # --nu > 0: tpCN moves (Student-t reference) in the same fused kernel
        out = tuple(torch.empty_like(t) for t in (x, ll, lp, lq))
        res = {}

        def run():
            res["o"] = ops.pcn_sweep(x, ll, lp, lq, idx, pg["like"], pg["prior"], pg["q"], prm, out=out)

        ms = timed(run, args.iters)
        acc = float(res["o"][4].sum()) / (n * args.moves)
        w = (16 * dims + 48) * n
        print(json.dumps({"bench": "tpcn_sweep" if args.nu > 0 else "pcn_sweep", "nu": args.nu, "problem": name, "particles": n, "dims": dims, "moves": args.moves,
                          "ms": ms, "particle_moves_per_s": n * args.moves / ms * 1e3, "normals_per_s":
                          n * args.moves * dims / ms * 1e3, "algorithmic_gbs": w / ms / 1e6,
                          "hbm_frac": w / ms / 1e6 / PEAKS["hbm_gbs"], "acceptance": acc,
                          "legacy": os.environ.get("ASP_SWEEP_LEGACY", "0"), "lib": os.environ.get("ASPIRE_B200_LIB", "")}))
        if args.dump:
            m = min(n, 20_000)
            dump[name + "_x"] = res["o"][0][:m].cpu().numpy()
            dump[name + "_ll"] = res["o"][1][:m].cpu().numpy()
            dump[name + "_acc"] = res["o"][4].cpu().numpy()
        del x, lq, lp, ll, idx, out, res
    if args.dump:
        np.savez(args.dump, **dump)


def c4(args):
    """C4: 64-D 8-component mixture, coupling-flow proposal evaluated on tcgen05 inside the loop.  One tempering step
    = reweight + beta bisection + exact resample + `moves` pCN moves; the proposal density is the flow, so every move
    is asp_pcn_propose -> flow log_prob (tcgen05) -> prior / likelihood kernels -> asp_pcn_accept."""
    from aspire_b200.flows import CouplingFlow
    from aspire_b200.history import SMCHistory
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import SMCSampler
    from aspire_b200.samples import SMCSamples

    dev = torch.device("cuda")
    n, dims, moves = args.n or 10_000_000, 64, args.moves
    prob = tg.gmm_problem(dims, device=dev)
    flow = CouplingFlow(dims, device=dev, n_layers=3, rng=PhiloxRNG(3), seed=1)
    flow.std = torch.full((dims,), 2.5, dtype=torch.float64, device=dev)  # broad proposal over the mixture
    flow._struct = None
    s = SMCSampler(prob["like"], prob["prior"], dims, flow, rng=PhiloxRNG(4))
    assert not s.fused
    s._configure_kernel(dict(n_steps=moves, rho=0.3, adapt_rho=False))
    s.resampling, s.target_efficiency, s.target_efficiency_rate = "multinomial", 0.5, 1.0
    s.adaptive, s.adaptive_min_beta_step, s.max_beta_step = True, False, 1.0
    s.history = SMCHistory()
    x0, lq0 = flow.sample_and_log_prob(n)
    p0 = SMCSamples(x0, log_likelihood=prob["like"].log_prob(x0), log_prior=prob["prior"].log_prob(x0), log_q=lq0,
                    beta=0.0, device=dev)
    t_flow = timed(lambda: flow.log_prob(x0), 3)

    def step():
        p0.__dict__.pop("_slope_cache", None)
        return s.tempering_step(p0, 0.0, float("nan"), 0.0, 1e-6)

    ms = timed(step, max(2, args.iters // 3))
    flow._check_error()
    half, h = dims // 2, 64
    flops = 3 * 2 * (half * h + h * h + h * dims) * n * moves
    print(json.dumps({"bench": "c4_smc_step_flow_proposal", "config": "C4", "particles": n, "dims": dims, "moves": moves,
                      "ms_per_step": ms, "particles_per_s": n / ms * 1e3, "ms_flow_logprob": t_flow,
                      "flow_share": moves * t_flow / ms, "flow_tflops_in_step": flops / ms / 1e9,
                      "beta": s.history.beta[-1], "acceptance": s.history.mcmc_acceptance[-1]}))


def flowsweep(args):
    """asp_pcn_sweep_flow alone (C4 shape: 64-D, 8-component mixture, coupling-flow proposal): gather + `moves` pCN
    moves with log q(x') on tcgen05 inside the move, Philox noise, sorted ancestors."""
    from aspire_b200.flows import CouplingFlow
    from aspire_b200.rng import PhiloxRNG

    dev = torch.device("cuda")
    n, dims, moves = args.n or 10_000_000, args.dims or 64, args.moves
    prob = tg.gmm_problem(dims, device=dev) if not args.only else tg.gaussian_problem(dims, device=dev)
    flow = CouplingFlow(dims, device=dev, n_layers=3, rng=PhiloxRNG(3), seed=1)
    flow.std = torch.full((dims,), 2.5, dtype=torch.float64, device=dev)
    flow._struct = None
    x, lq = flow.sample_and_log_prob(n)
    ll, lp = prob["like"].log_prob(x), prob["prior"].log_prob(x)
    idx = torch.sort((ops.philox_uniform(6, 0, n, dev) * n).to(torch.int64).clamp_(0, n - 1)).values
    mu, sigma = ops.moments_to_mu_sigma(ops.weighted_moments(x[: 1 << 20]), dims)
    prm = ops.make_pcn_params(0.01, 0.3, mu, sigma, moves, seed=99, stream0=3, total_accepts_only=True)
    out = tuple(torch.empty_like(t) for t in (x, ll, lp, lq))
    res = {}

    def run():
        res["o"] = ops.pcn_sweep_flow(x, ll, lp, lq, idx, prob["like"], prob["prior"], flow.struct(), prm, out=out)

    ms = timed(run, args.iters)
    flow._check_error()
    half, h = dims // 2, 64
    flops = 3 * 2 * (half * h + h * h + h * dims) * n * moves
    print(json.dumps({"bench": "pcn_sweep_flow", "particles": n, "dims": dims, "moves": moves, "ms": ms,
                      "particle_moves_per_s": n * moves / ms * 1e3, "flow_tflops": flops / ms / 1e9,
                      "algorithmic_gbs": (16 * dims + 48) * n / ms / 1e6,
                      "acceptance": float(res["o"][4].sum()) / (n * moves), "lib": os.environ.get("ASPIRE_B200_LIB", "")}))


def full(args):
    """A whole ``Aspire.sample_posterior(sampler="smc")`` run (all tempering steps to beta = 1) through the
    drop-in API: 32-D Rosenbrock, N particles, 10 pCN moves per step, adaptive beta at efficiency 0.5."""
    import time

    from aspire_b200.aspire import Aspire
    from aspire_b200.flows import DiagGaussianFlow
    from aspire_b200.rng import PhiloxRNG

    dev = torch.device("cuda")
    n, dims = args.n or 10_000_000, 32
    prob = tg.rosenbrock_problem(dims, device=dev)
    flow = DiagGaussianFlow(dims, device=dev, mean=prob["q"].mean_np, sigma=prob["q"].sigma_np, rng=PhiloxRNG(1))
    a = Aspire(log_likelihood=prob["like"], log_prior=prob["prior"], dims=dims, flow=flow, device="cuda")
    for rep in range(2):  # the first run warms the allocator and the kernels
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out, hist = a.sample_posterior(n, sampler="smc", return_history=True, rng=PhiloxRNG(2 + rep),
                                       sampler_kwargs=dict(n_steps=args.moves, **({"step_fn": args.only} if args.only else {})),
                                       store_sample_history=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    steps = len(hist.beta)
    print(json.dumps({"bench": "smc_full_run", "step_fn": args.only or "tpcn (MiniPCNSMC default)", "particles": n, "dims": dims, "moves": args.moves, "tempering_steps": steps,
                      "seconds": dt, "ms_per_step": dt / steps * 1e3, "particle_steps_per_s": n * steps / dt,
                      "log_evidence": float(out.log_evidence), "log_evidence_error": float(out.log_evidence_error),
                      "mean_acceptance": float(sum(hist.mcmc_acceptance) / steps),
                      "likelihood_evaluations": a.n_likelihood_evaluations}))


def precond(args):
    """Scope row 8f-4 end to end: one tempering step of a 32-D problem with bounded parameters - Gaussian likelihood
    inside the unit box, flat prior, the proposal a Gaussian on the logit image of the box (`TransformedFlow`) and the
    pCN moves in the logit + affine preconditioned space (`SMCSampler._mutate_preconditioned`: per move one
    propose, one transform, three density evaluations, one accept - unfused).  The same step without bounds runs
    the fused sweep; both are timed."""
    from aspire_b200 import transforms as T
    from aspire_b200.flows import DiagGaussianFlow, TransformedFlow
    from aspire_b200.history import SMCHistory
    from aspire_b200.rng import PhiloxRNG
    from aspire_b200.samplers import SMCSampler
    from aspire_b200.samples import SMCSamples

    dev = torch.device("cuda")
    n, dims, moves = args.n or 10_000_000, 32, args.moves
    params = [f"x{j}" for j in range(dims)]
    bounds = {p: [0.0, 1.0] for p in params}
    like = tg.GaussianDensity([0.3 + 0.4 * j / dims for j in range(dims)], 0.15, dims, normalised=False, device=dev)
    prior = tg.UniformDensity(0.0, 1.0, dims, device=dev)
    ft = T.FlowTransform(params, prior_bounds=bounds, bounded_transform="logit", affine_transform=False, device=dev)
    flow = TransformedFlow(DiagGaussianFlow(dims, device=dev, mean=[0.0] * dims, sigma=[1.2] * dims, rng=PhiloxRNG(3)), ft)
    pt = T.CompositeTransform(params, prior_bounds=bounds, bounded_to_unbounded=True, bounded_transform="logit",
                              affine_transform=True, device=dev)
    out = {}
    for name, transform in (("preconditioned", pt), ("unbounded_space_reference", None)):
        s = SMCSampler(like, prior, dims, flow, rng=PhiloxRNG(4), preconditioning_transform=transform)
        s._configure_kernel(dict(n_steps=moves, rho=0.3, adapt_rho=False))
        s.resampling, s.target_efficiency, s.target_efficiency_rate = "multinomial", 0.5, 1.0
        s.adaptive, s.adaptive_min_beta_step, s.max_beta_step = True, False, 1.0
        s.history = SMCHistory()
        x0, lq0 = flow.sample_and_log_prob(n)
        p0 = SMCSamples(x0, log_likelihood=like.log_prob(x0), log_prior=prior.log_prob(x0), log_q=lq0, beta=0.0,
                        device=dev)

        def step():
            p0.__dict__.pop("_slope_cache", None)
            return s.tempering_step(p0, 0.0, float("nan"), 0.0, 1e-6)

        ms = timed(step, max(2, args.iters // 3))
        out[name] = dict(ms_per_step=ms, particles_per_s=n / ms * 1e3, beta=s.history.beta[-1],
                         acceptance=s.history.mcmc_acceptance[-1])
        del p0, x0, lq0
    print(json.dumps({"bench": "smc_step_bounded_parameters", "particles": n, "dims": dims, "moves": moves,
                      "route": "preconditioned: unfused moves in the logit + affine space (propose, transform, 3 density "
                               "evaluations, accept per move); unbounded_space_reference: the same problem and proposal "
                               "without preconditioning, semi-fused moves on x (asp_pcn_propose_eval + accept_scalar)",
                      **out}))


def transform(args):
    """Preconditioning transforms (scope row 8f-4): forward and inverse pass over N x 32 rows, for an all-logit, an
    all-probit, a periodic + affine and a mixed description.  Algorithmic bytes: 8 D read + 8 D + 8 written."""
    import math

    from aspire_b200 import transforms as T

    dev = torch.device("cuda")
    n, d = args.n or 20_000_000, 32
    params = [f"x{j}" for j in range(d)]
    z = ops.philox_normal(1, 0, n, d, dev)
    cases = {
        "logit": dict(bounds={p: [0.0, 1.0] for p in params}, periodic=[], bounded="logit", affine=False),
        "probit": dict(bounds={p: [0.0, 1.0] for p in params}, periodic=[], bounded="probit", affine=False),
        "periodic+affine": dict(bounds={p: [0.0, 2 * math.pi] for p in params}, periodic=params, bounded="logit",
                                affine=True),
        "mixed": dict(bounds={p: ([0.0, 1.0] if j < 12 else [0.0, 2 * math.pi] if j < 16 else [-math.inf, math.inf])
                              for j, p in enumerate(params)}, periodic=params[12:16], bounded="logit", affine=True),
    }
    nbytes = n * (16 * d + 8)
    for name, c in cases.items():
        t = T.CompositeTransform(params, periodic_parameters=c["periodic"], prior_bounds=c["bounds"],
                                 bounded_transform=c["bounded"], affine_transform=c["affine"], device=dev)
        if c["affine"]:
            t.fit(ops.philox_uniform(2, 0, 4096 * d, dev).reshape(4096, d).contiguous())
        x = t.inverse(z)[0]
        for direction, fn in (("inverse", lambda: t.inverse(z)), ("forward", lambda: t.forward(x))):
            ms = timed(fn, args.iters)
            print(json.dumps({"bench": "transform", "case": name, "direction": direction, "rows": n, "dims": d, "ms": ms,
                              "rows_per_s": n / ms * 1e3, "algorithmic_gbs": nbytes / ms / 1e6,
                              "frac_of_hbm_peak": nbytes / ms / 1e6 / PEAKS["hbm_gbs"],
                              "note": "includes the allocation of the two outputs by the caching allocator"}))
        del x


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["flow", "resample", "importance", "sweep", "c4", "flowsweep", "full", "transform", "precond"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--moves", type=int, default=10)
    ap.add_argument("--dims", type=int, default=0)
    ap.add_argument("--sorted-idx", action="store_true")
    ap.add_argument("--nu", type=float, default=0.0, help="sweep: > 0 runs tpCN moves with this many degrees of freedom")
    ap.add_argument("--numpy-max", type=int, default=10**7, help="largest n timed with numpy on the host (resample)")
    ap.add_argument("--gpus", type=int, default=1, help="resample: launch with torch.distributed.run for N > 1")
    ap.add_argument("--only", default="")
    ap.add_argument("--dump", default="")
    a = ap.parse_args()
    ops.device_check()
    {"flow": flow, "resample": resample, "importance": importance, "sweep": sweep, "c4": c4, "flowsweep": flowsweep, "full": full, "transform": transform, "precond": precond}[a.what](a)
