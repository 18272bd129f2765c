"""world_size-2 gloo test of the sharded SMC step (host-side orchestration of
aspire_b200.distributed) on CPU: the union over ranks equals the single-process step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import pcn as opcn
from oracle import philox as oph
from oracle import problems
from oracle import smc_numpy as sn

N, DIMS, MOVES, SEED = 6000, 6, 3, 99


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, splits, headroom, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_backend

    from aspire_b200.distributed import Comm, ShardedSMC, Shard

    prob = problems.rosenbrock_problem(DIMS)
    x, ll, lp, lq = problems.draw(prob, N, 4)
    a, b = splits[rank], splits[rank + 1]
    t = lambda v: torch.as_tensor(np.ascontiguousarray(v))  # noqa: E731
    shard = Shard(t(x[a:b]), t(ll[a:b]), t(lp[a:b]), t(lq[a:b]))
    eng = ShardedSMC(Comm(), prob["like"], prob["prior"], prob["q"], backend=cpu_backend, seed=SEED, n_moves=MOVES,
                     rho=0.4, headroom=headroom)
    new, rec = eng.step(shard, 0.0)
    ret[rank] = dict(slot0=new.slot0, x=new.x.numpy(), ll=new.log_l.numpy(), rec=rec)
    dist.destroy_process_group()


@pytest.mark.parametrize("splits,headroom", [([0, 3000, 6000], 1.04), ([0, 1, 6000], 1.04), ([0, 4097, 6000], 1.04),
                                             ([0, 3000, 6000], 0.0), ([0, 1, 6000], 0.0)])
def test_two_rank_step_equals_single_process(splits, headroom):
    """``headroom = 0`` shrinks the margins to the 4096-row floor: with the [0, 1, 6000] split rank 1 produces slots
    far outside them, sweeps into scratch and copies the rows it keeps."""
    import sys

    from oracle import sorted_resample as sr

    sys.path.insert(0, os.path.dirname(__file__))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), splits, headroom, ret), nprocs=2, join=True)
    # single-process oracle of the same step (same Philox streams: spacings stream 0, noise from stream 1 + 2^30)
    prob = problems.rosenbrock_problem(DIMS)
    x, ll, lp, lq = problems.draw(prob, N, 4)
    beta, _ = sn.determine_beta(ll, lp, lq, 0.0, beta_tolerance=1e-6)
    p = sn.resample_weights(ll, lp, lq, 0.0, beta)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    idx = sr.resample(cdf, SEED, 0, N)
    mu, sigma = opcn.weighted_moments(x[idx], np.ones(N))
    ref = opcn.sweep(x[idx], ll[idx], lp[idx], lq[idx], like=prob["like"], prior=prob["prior"], q=prob["q"], beta=beta,
                     mu=mu, sigma=sigma, rho=0.4, n_moves=MOVES, seed=SEED, stream0=1 + 0x40000000)
    # rebalanced: every rank holds its N / G slots, in slot order
    assert ret[0]["slot0"] == 0 and ret[1]["slot0"] == N // 2
    assert len(ret[0]["x"]) == N // 2 and len(ret[1]["x"]) == N - N // 2
    xs = np.concatenate([ret[0]["x"], ret[1]["x"]])
    lls = np.concatenate([ret[0]["ll"], ret[1]["ll"]])
    # ancestors live where their offspring were born: rank 0 produced exactly the slots with idx < split
    assert ret[0]["rec"].n_local_out == int((idx < splits[1]).sum())
    assert ret[1]["rec"].n_local_out == int((idx >= splits[1]).sum())
    np.testing.assert_allclose(xs, ref[0], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(lls, ref[1], rtol=1e-8, atol=1e-8)
    for r in (0, 1):
        rec = ret[r]["rec"]
        assert rec.beta == beta and rec.n_total == N
        assert rec.ess == pytest.approx(float(sn.effective_sample_size(sn.log_weights(ll, lp, lq, 0.0, beta))), rel=1e-10)
        assert rec.log_norm_ratio == pytest.approx(sn.log_evidence_ratio(ll, lp, lq, 0.0, beta), rel=1e-10)
        assert rec.log_norm_ratio_var == pytest.approx(sn.log_evidence_ratio_variance(ll, lp, lq, 0.0, beta), rel=1e-9)
        assert rec.acceptance == pytest.approx(ref[4].sum() / (MOVES * N), abs=1e-12)


def _loop_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_backend

    from aspire_b200.distributed import Comm, ShardedSMC, Shard, run_tempering

    n, dims = 3000, 4
    prob = problems.gaussian_problem(dims)
    x, ll, lp, lq = problems.draw(prob, n, 8)
    a, b = rank * n // world, (rank + 1) * n // world
    t = lambda v: torch.as_tensor(np.ascontiguousarray(v))  # noqa: E731
    eng = ShardedSMC(Comm(), prob["like"], prob["prior"], prob["q"], backend=cpu_backend, seed=7, n_moves=2, rho=0.5)
    shard, recs, log_z, err = run_tempering(eng, Shard(t(x[a:b]), t(ll[a:b]), t(lp[a:b]), t(lq[a:b])))
    ret[(world, rank)] = dict(beta=[r.beta for r in recs], log_z=log_z, err=err, slot0=shard.slot0,
                             x=shard.x.numpy())
    dist.destroy_process_group()


def test_two_rank_tempering_loop_equals_single_rank_loop():
    """The whole loop to beta = 1 on two ranks against one rank: the shards are rebalanced into slot order after
    every step, so the two runs hold the same particle in the same global row - same beta schedule, same evidence
    increments, same ensemble row by row (the all-reduced sums differ in the last bits only)."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_loop_worker, args=(1, _free_port(), ret), nprocs=1, join=True)
    mp.spawn(_loop_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    one, r0, r1 = ret[(1, 0)], ret[(2, 0)], ret[(2, 1)]
    assert one["beta"][-1] == 1.0 and len(one["beta"]) >= 2
    assert r0["beta"] == one["beta"] and r1["beta"] == one["beta"]
    assert r0["log_z"] == pytest.approx(one["log_z"], rel=1e-10) and r1["log_z"] == pytest.approx(one["log_z"], rel=1e-10)
    assert r0["err"] == pytest.approx(one["err"], rel=1e-8)
    assert (r0["slot0"], r1["slot0"]) == (0, len(r0["x"])) and len(r0["x"]) + len(r1["x"]) == len(one["x"])
    xs = np.concatenate([r0["x"], r1["x"]])
    np.testing.assert_allclose(xs, one["x"], rtol=1e-8, atol=1e-8)
    var = 1.0 / (1.0 / 0.49 + 1.0)
    for ens in (xs, one["x"]):
        assert np.abs(ens.mean(axis=0) - var * 0.5 / 0.49).max() < 0.06
        np.testing.assert_allclose(ens.var(axis=0), var, rtol=0.15)
    # analytic evidence of the Gaussian problem (C2): prior N(0, I) x likelihood N(0.5, 0.7^2 I)
    d = 4
    exact = d * (-0.5 * np.log(2 * np.pi * (1.0 + 0.49)) - 0.5 * 0.25 / (1.0 + 0.49))
    assert one["log_z"] == pytest.approx(exact, abs=0.15)


def _peer_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aspire_b200.distributed import Comm, PeerMailboxes

    # no CUDA device: every rank must agree to keep the NCCL / gloo collectives (create is itself collective)
    ret[rank] = PeerMailboxes.create(Comm(), "cpu") is None
    dist.destroy_process_group()


def test_peer_mailboxes_fall_back_together_without_gpus():
    """The in-kernel collectives are an all-or-nothing upgrade: ``PeerMailboxes.create`` returns None on every rank
    when any rank cannot map the mailboxes (here: no GPU at all), and the sharded step keeps its collectives."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_peer_descriptor_layout_matches_the_header():
    """``asp_peer_t`` / ``ASP_PEER_BOX_BYTES`` in include/aspire_b200.h against the ctypes mirror."""
    import ctypes as C
    import re

    from aspire_b200 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = open(os.path.join(root, "include", "aspire_b200.h")).read()
    max_peers = int(re.search(r"#define ASP_MAX_PEERS (\d+)", h).group(1))
    box = re.search(r"#define ASP_PEER_BOX_BYTES \((.+)\)", h).group(1).replace("ASP_MAX_PEERS", str(max_peers))
    assert _lib.MAX_PEERS == max_peers
    assert re.fullmatch(r"[0-9+*() ]+", box), box  # plain integer arithmetic only
    assert _lib.PEER_BOX_BYTES == eval(box, {"__builtins__": {}})  # noqa: S307
    assert C.sizeof(_lib.PeerDesc) == 8 * max_peers + 8
    assert _lib.PeerDesc.rank.offset == 8 * max_peers and _lib.PeerDesc.world.offset == 8 * max_peers + 4
    # the chain region (csrc/scan_exact.cu: kChainBase = 2 * 8 * 16 + 8, counter 8 doubles later) fits in the box
    assert (2 * max_peers * 16 + 8 + 8 + 1) * 8 <= _lib.PEER_BOX_BYTES


def _long_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cpu_backend

    from aspire_b200.distributed import Comm, ShardedSMC, Shard

    n, dims, steps = 2000, 2, 200
    prob = problems.gaussian_problem(dims)
    x, ll, lp, lq = problems.draw(prob, n, 21)
    order = np.argsort(-ll)  # rank 0 starts with the high-likelihood half: the first steps are as skewed as it gets
    x, ll, lp, lq = x[order], ll[order], lp[order], lq[order]
    a, b = rank * n // world, (rank + 1) * n // world
    t = lambda v: torch.as_tensor(np.ascontiguousarray(v))  # noqa: E731
    # a low efficiency target makes every resampling step degenerate (few distinct ancestors): the ranks' shares of
    # the offspring fluctuate strongly from step to step
    eng = ShardedSMC(Comm(), prob["like"], prob["prior"], prob["q"], backend=cpu_backend, seed=5, n_moves=1, rho=0.5,
                     target_efficiency=0.05, headroom=0.0)
    shard = Shard(t(x[a:b]), t(ll[a:b]), t(lp[a:b]), t(lq[a:b]))
    sizes, produced, slots = [], [], []
    for _ in range(steps):
        shard, rec = eng.step(shard, 0.0)  # re-temper the current ensemble from beta = 0 every time
        sizes.append(len(shard))
        produced.append(rec.n_local_out)
        slots.append(shard.slot0)
    ret[rank] = dict(sizes=sizes, produced=produced, slots=slots, finite=bool(torch.isfinite(shard.x).all()))
    dist.destroy_process_group()


def test_two_hundred_skewed_steps_keep_the_shards_balanced():
    """200 consecutive steps on two ranks with strongly fluctuating offspring shares (efficiency target 0.05, the
    ensemble initially sorted by likelihood): no step raises, every shard holds exactly N / G particles after every
    step (the surplus is shipped peer to peer, nothing depends on head-room), and the exchange really happens."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_long_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    r0, r1 = ret[0], ret[1]
    assert r0["sizes"] == [1000] * 200 and r1["sizes"] == [1000] * 200
    assert set(r0["slots"]) == {0} and set(r1["slots"]) == {1000}
    assert all(p0 + p1 == 2000 for p0, p1 in zip(r0["produced"], r1["produced"]))
    surplus = np.abs(np.array(r0["produced"]) - 1000)
    assert surplus.max() >= 50 and (surplus > 0).mean() > 0.9  # rows crossed the shard boundary in almost every step
    assert r0["finite"] and r1["finite"]
