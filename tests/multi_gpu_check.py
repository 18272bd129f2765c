"""Launched by torchrun (one rank per GPU) from tests/test_gpu_multi.py: the sharded
step on G GPUs against the single-GPU step on rank 0 (same global particles, same
Philox streams).  Exits non-zero on mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from aspire_b200 import ops  # noqa: E402
from aspire_b200 import targets as tg  # noqa: E402
from aspire_b200.distributed import Comm, Shard, ShardedSMC  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_003
    dims, moves, seed = 32, 4, 4242
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    r, g = comm.rank, comm.world
    prob = tg.rosenbrock_problem(dims, device=dev)
    base = n_total // g
    n_local = base + (1 if r < n_total % g else 0)
    start = r * base + min(r, n_total % g)
    x0 = ops.philox_normal(seed, 0, n_local, dims, dev, mean=prob["q"].mean, sigma=prob["q"].sigma, index0=start)
    shard = Shard(x0, prob["like"].log_prob(x0), prob["prior"].log_prob(x0), prob["q"].log_prob(x0))
    eng = ShardedSMC(comm, prob["like"], prob["prior"], prob["q"], seed=seed + 1, n_moves=moves, rho=0.5)
    new, rec = eng.step(shard, 0.0)
    # the bisection with its all-reduce inside the pass kernel (NVLink peer memory) against the NCCL route:
    # the same state, bit for bit, on every rank
    d, stats_local = ops.reweight(shard.log_l, shard.log_p, shard.log_q)
    stats = eng.global_stats(stats_local)
    via_nccl = ops.beta_search(d, stats, 0.0, 0.5, 1e-6, n_total=n_total, allreduce=comm.all_reduce_sum)
    p2p_note = "in-kernel all-reduce: not available (NCCL route used)"
    if eng._peer is not None:
        def timed(fn, reps=10):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            return out, e0.elapsed_time(e1) / reps
        via_p2p, ms_p2p = timed(lambda: ops.beta_search(d, stats, 0.0, 0.5, 1e-6, n_total=n_total,
                                                        peer=eng._peer.desc))
        _, ms_nccl = timed(lambda: ops.beta_search(d, stats, 0.0, 0.5, 1e-6, n_total=n_total,
                                                   allreduce=comm.all_reduce_sum))
        # root-guided search over the mailboxes (default) against the plain bisection over NCCL: the same bracket,
        # evaluation count and sums; the plain bisection over the mailboxes: the same state bit for bit
        assert (via_p2p.beta_min, via_p2p.beta_max, via_p2p.n_evals, via_p2p.min_valid, via_p2p.eff_beta_max) == (
            via_nccl.beta_min, via_nccl.beta_max, via_nccl.n_evals, via_nccl.min_valid, via_nccl.eff_beta_max)
        assert abs(via_p2p.s1_min - via_nccl.s1_min) <= 1e-12 * via_nccl.s1_min
        assert abs(via_p2p.s2_min - via_nccl.s2_min) <= 1e-12 * via_nccl.s2_min
        plain_p2p = ops.beta_search(d, stats, 0.0, 0.5, 1e-6, n_total=n_total, peer=eng._peer.desc, root_guided=False)
        assert bytes(plain_p2p) == bytes(via_nccl), (plain_p2p.beta_min, via_nccl.beta_min)
        p2p_note = f"in-kernel all-reduce over {eng._peer.how}: {ms_p2p:.3f} ms per search, NCCL {ms_nccl:.3f} ms"
    if eng._peer is not None:
        # the exact-CDF carry chain through the mailboxes against send / recv / broadcast: same slice, bit for bit
        dbeta = via_nccl.beta_min
        pw, result, tiles = ops.smc_weights(d, dbeta, dbeta * float(stats[1]), via_nccl.s1_min, n_total=n_total)
        info = torch.stack([result[1], tiles[1::2].max(), result[0], result[0] * 0 + len(shard)]).contiguous()
        cdf_a, chain_a, _ = eng.global_cdf(pw, tiles, info, n_total)
        keep, eng._peer = eng._peer, None
        cdf_b, chain_b, _ = eng.global_cdf(pw, tiles, info, n_total)
        eng._peer = keep
        lo_a, hi_a = eng.slice_bounds(chain_a, cdf_a, r, g)
        lo_b, hi_b = eng.slice_bounds(chain_b, cdf_b, r, g)
        assert torch.equal(cdf_a, cdf_b) and (lo_a, hi_a) == (lo_b, hi_b), (lo_a, lo_b, hi_a, hi_b)
        _, ms_a = timed(lambda: eng.global_cdf(pw, tiles, info, n_total))
        eng._peer = None
        _, ms_b = timed(lambda: eng.global_cdf(pw, tiles, info, n_total))
        eng._peer = keep
        p2p_note += f"; CDF chain over peer memory == NCCL chain ({ms_a:.3f} ms vs {ms_b:.3f} ms)"
    if r == 0:
        print(p2p_note)
    # gather (x, logL) on rank 0: the shards are rebalanced, rank r holds the slots [r N/G, (r+1) N/G) in order
    cnt = torch.tensor([len(new), new.slot0, rec.n_local_out], dtype=torch.int64, device=dev)
    info = comm.all_gather(cnt).tolist()
    counts = [c[0] for c in info]
    cap = max(counts)
    pad = lambda t: torch.cat([t, t.new_zeros((cap - t.shape[0],) + tuple(t.shape[1:]))])  # noqa: E731
    xs = comm.all_gather(pad(new.x))
    lls = comm.all_gather(pad(new.log_l))
    ok = True
    if r == 0:
        x = torch.cat([xs[i, :c] for i, c in enumerate(counts)])
        ll = torch.cat([lls[i, :c] for i, c in enumerate(counts)])
        assert sum(counts) == n_total and max(counts) - min(counts) <= 1, counts
        assert [c[1] for c in info] == [sum(counts[:i]) for i in range(g)], "shards are not in slot order"
        assert sum(c[2] for c in info) == n_total, "the produced slot ranges are not a partition"
        order = torch.arange(n_total, device=dev)
        # single-GPU reference: the same engine with a trivial communicator
        class Solo(Comm):
            def __init__(self):
                self.rank, self.world = 0, 1
            def all_reduce_sum(self, t): return t
            def all_reduce_max(self, t): return t
            def all_gather(self, t): return t.unsqueeze(0)
            def broadcast(self, t, src): return t
            def send(self, t, dst): pass
            def recv(self, t, src): pass
            def barrier(self): pass
        xg = ops.philox_normal(seed, 0, n_total, dims, dev, mean=prob["q"].mean, sigma=prob["q"].sigma)
        full = Shard(xg, prob["like"].log_prob(xg), prob["prior"].log_prob(xg), prob["q"].log_prob(xg))
        solo = ShardedSMC(Solo(), prob["like"], prob["prior"], prob["q"], seed=seed + 1, n_moves=moves, rho=0.5)
        ref, rrec = solo.step(full, 0.0)
        assert ref.slot0 == 0 and len(ref) == n_total
        assert rrec.beta == rec.beta, (rrec.beta, rec.beta)
        np.testing.assert_allclose(rec.ess, rrec.ess, rtol=1e-12)
        np.testing.assert_allclose(rec.log_norm_ratio, rrec.log_norm_ratio, rtol=1e-12)
        np.testing.assert_allclose(rec.log_norm_ratio_var, rrec.log_norm_ratio_var, rtol=1e-10)
        dx = (x[order] - ref.x).abs().max(dim=1).values
        frac_bad = float((dx > 1e-9).double().mean())
        print(f"G={g} N={n_total} beta={rec.beta:.6g} acc={rec.acceptance:.4f} produced per rank={[c[2] for c in info]} "
              f"max|dx|={float(dx.max()):.3e} frac(|dx|>1e-9)={frac_bad:.2e}")
        ok = frac_bad < 1e-5 and abs(rec.acceptance - rrec.acceptance) < 1e-5
        np.testing.assert_allclose(ll[order][dx <= 1e-9].cpu().numpy(), ref.log_l[dx <= 1e-9].cpu().numpy(),
                                   rtol=1e-9, atol=1e-9)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    comm.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if float(flag[0]) == 1.0 else 1)


if __name__ == "__main__":
    main()
