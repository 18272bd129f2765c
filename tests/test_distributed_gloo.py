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


def _worker(rank, world, port, splits, ret):
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
                     rho=0.4)
    new, rec = eng.step(shard, 0.0)
    ret[rank] = dict(slot=new.slot.numpy(), x=new.x.numpy(), ll=new.log_l.numpy(), rec=rec)
    dist.destroy_process_group()


@pytest.mark.parametrize("splits", [[0, 3000, 6000], [0, 1, 6000], [0, 4097, 6000]])
def test_two_rank_step_equals_single_process(splits):
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), splits, ret), nprocs=2, join=True)
    # single-process oracle of the same step (same Philox streams: uniforms stream 0, noise from stream 1 + 2^30)
    prob = problems.rosenbrock_problem(DIMS)
    x, ll, lp, lq = problems.draw(prob, N, 4)
    beta, _ = sn.determine_beta(ll, lp, lq, 0.0, beta_tolerance=1e-6)
    p = sn.resample_weights(ll, lp, lq, 0.0, beta)
    u = oph.uniform_f64(SEED, 0, np.arange(N))
    idx = sn.choice_indices(p, u)
    mu, sigma = opcn.weighted_moments(x[idx], np.ones(N))
    ref = opcn.sweep(x[idx], ll[idx], lp[idx], lq[idx], like=prob["like"], prior=prob["prior"], q=prob["q"], beta=beta,
                     mu=mu, sigma=sigma, rho=0.4, n_moves=MOVES, seed=SEED, stream0=1 + 0x40000000)
    slot = np.concatenate([ret[0]["slot"], ret[1]["slot"]])
    xs = np.concatenate([ret[0]["x"], ret[1]["x"]])
    lls = np.concatenate([ret[0]["ll"], ret[1]["ll"]])
    assert sorted(slot.tolist()) == list(range(N))          # every output slot born exactly once
    order = np.argsort(slot)
    # ancestors live where their offspring were born: rank 0 holds exactly the slots with idx < split
    assert set(ret[0]["slot"].tolist()) == set(np.nonzero(idx < splits[1])[0].tolist())
    np.testing.assert_allclose(xs[order], ref[0], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(lls[order], ref[1], rtol=1e-8, atol=1e-8)
    for r in (0, 1):
        rec = ret[r]["rec"]
        assert rec.beta == beta and rec.n_total == N
        assert rec.ess == pytest.approx(float(sn.effective_sample_size(sn.log_weights(ll, lp, lq, 0.0, beta))), rel=1e-10)
        assert rec.log_norm_ratio == pytest.approx(sn.log_evidence_ratio(ll, lp, lq, 0.0, beta), rel=1e-10)
        assert rec.log_norm_ratio_var == pytest.approx(sn.log_evidence_ratio_variance(ll, lp, lq, 0.0, beta), rel=1e-9)
        assert rec.acceptance == pytest.approx(ref[4].sum() / (MOVES * N), abs=1e-12)
