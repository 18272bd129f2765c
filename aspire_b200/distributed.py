"""Particles sharded over the GPUs of one box: one process per GPU, ``torch.distributed``
(NCCL over NVLink) for the plumbing.  No reference site - the reference is a single
process; the per-rank arithmetic is the same C-ABI as on one GPU.

Rank r owns a contiguous slice of the particle set (and of the weights / CDF).  Per
tempering step the ranks exchange only scalars:

* ``max``  all-reduce of {-min d, max d, #NaN, #inf}                       (1 x 4 doubles)
* ``sum``  all-reduce of (S1, S2) per bisection pass                      (~22 x 2 doubles)
* all-gather of {sum p, max p, var-sum, n} per rank                      (1 x 4 doubles)
* exact-CDF carry chain: rank r-1 -> r, {exact running sum, running max} (G-1 hops x 2 doubles)
* broadcast of the exact total                                           (1 double)
* ``sum``  all-reduce of the pCN reference-measure moments                (1 + 2 D doubles)
* ``sum``  all-reduce of the accept counters                              (n_moves int64)

* all-gather of the integer tile sums of the sorted-uniform spacings       (N / 2048 int64 in total)
* all-gather of the CDF value in front of every slice                    (1 double per rank)

Resampling is OWNER-COMPUTES on SORTED uniforms (``asp_sorted_spacings`` / ``asp_search_sorted``,
``oracle/sorted_resample.py``): output slot j is the j-th smallest uniform, so the rank that holds the CDF slice
[k0, k1) owns the CONTIGUOUS slot range {j : T_{k0-1} <= S_j < T_{k1-1}}; it regenerates those spacings itself
(counter-based), resolves its ancestors locally and mutates the offspring where their ancestors live.  The union
over ranks of (slot, ancestor) is exactly the single-GPU result, slot for slot.  Afterwards the shards are
REBALANCED: rank r keeps the slots of its target range [r N/G, (r+1) N/G) and only the slivers that cross a shard
boundary - the weight imbalance of the step - travel peer to peer (one grouped NCCL send/recv over NVLink),
straight into their final rows; the sweep has already written the rows that stay into place.  The global array
stays in slot order, every shard keeps exactly N/G particles, and nothing depends on head-room.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch


class Comm:
    """Thin wrapper over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def all_gather(self, t):
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group) \
            if t.is_cuda else self.dist.all_gather(list(out.unbind(0)), t.contiguous(), group=self.group)
        return out

    def send(self, t, dst):
        self.dist.send(t, dst, group=self.group)

    def recv(self, t, src):
        self.dist.recv(t, src, group=self.group)

    def broadcast(self, t, src):
        self.dist.broadcast(t, src, group=self.group)
        return t

    def barrier(self):
        self.dist.barrier(group=self.group)


class SoloComm(Comm):
    """A one-rank communicator: the sharded step on a single process (reference side of the parity checks)."""

    def __init__(self):
        import torch.distributed as dist

        self.dist, self.group, self.rank, self.world = dist, None, 0, 1

    def all_reduce_sum(self, t):
        return t

    def all_reduce_max(self, t):
        return t

    def all_gather(self, t):
        return t.unsqueeze(0)

    def send(self, t, dst):
        pass

    def recv(self, t, src):
        pass

    def broadcast(self, t, src):
        return t

    def barrier(self):
        pass


class PeerMailboxes:
    """Mailboxes of the in-kernel all-reduce (``asp_beta_search_pass_p2p``): every rank owns
    ``ASP_PEER_BOX_BYTES`` of zeroed device memory that all the other ranks of the box can store into over
    NVLink.  Mapped with torch's symmetric memory (CUDA VMM handles exchanged through the group's store; plain
    CUDA IPC handles of a caching-allocator block were tried and fault without explicit peer enabling, so they
    are not used).  ``create`` returns None - and the caller keeps NCCL - unless EVERY rank mapped every
    mailbox.  ``ASPIRE_B200_P2P=0`` forces NCCL.  No reference site (single process)."""

    def __init__(self, desc, keep, how):
        self.desc, self._keep, self.how = desc, keep, how

    @staticmethod
    def _map_symmetric(comm: "Comm", device, n):
        import torch.distributed._symmetric_memory as symm
        box = symm.empty(n, dtype=torch.float64, device=device)
        box.zero_()
        torch.cuda.synchronize(device)
        hdl = symm.rendezvous(box, comm.group if comm.group is not None else comm.dist.group.WORLD)
        ptrs = [int(p) + int(getattr(hdl, "offset", 0)) for p in hdl.buffer_ptrs]
        if len(ptrs) != comm.world or not all(ptrs) or ptrs[comm.rank] != box.data_ptr():
            raise RuntimeError("symmetric memory: unexpected buffer pointers")
        return ptrs, (box, hdl), "symmetric memory"

    @classmethod
    def _map(cls, comm: "Comm", device):
        from . import _lib
        return cls._map_symmetric(comm, device, _lib.PEER_BOX_BYTES // 8)

    @classmethod
    def create(cls, comm: "Comm", device):
        import os
        from . import _lib
        device = torch.device(device)
        ok, mapped = 0.0, None
        if (device.type == "cuda" and 2 <= comm.world <= _lib.MAX_PEERS and comm.world <= torch.cuda.device_count()
                and os.environ.get("ASPIRE_B200_P2P", "1") != "0"):
            try:
                mapped = cls._map(comm, device)
                ok = 1.0
            except Exception:  # noqa: BLE001
                ok = 0.0
        flag = torch.tensor([-ok], dtype=torch.float64, device=device)
        comm.all_reduce_max(flag)  # every rank or none
        if float(flag[0]) != -1.0:
            return None
        ptrs, keep, how = mapped
        desc = _lib.PeerDesc()
        for r, ptr in enumerate(ptrs):
            desc.box[r] = ptr
        desc.rank, desc.world = comm.rank, comm.world
        torch.cuda.synchronize(device)
        comm.dist.barrier(group=comm.group)
        return cls(desc, keep, how)


@dataclass
class Shard:
    """This rank's particles: rows + logs of the global slots [slot0, slot0 + len) (slot order: the Philox
    counter of a particle's noise is its global slot)."""

    x: torch.Tensor
    log_l: torch.Tensor
    log_p: torch.Tensor
    log_q: torch.Tensor
    slot0: int = 0

    def __len__(self):
        return self.x.shape[0]


@dataclass
class StepRecord:
    beta: float
    ess: float
    ess_target: float
    log_norm_ratio: float
    log_norm_ratio_var: float
    acceptance: float
    n_total: int
    n_local_out: int
    n_beta_evals: int


class ShardedSMC:
    """One tempering step over sharded particles.  ``backend`` is ``aspire_b200.ops``
    (the CPU tests inject an oracle-backed stand-in with the same functions)."""

    def __init__(self, comm: Comm, like, prior, q, backend=None, seed: int = 0, n_moves: int = 10, rho: float = 0.5,
                 target_efficiency: float = 0.5, beta_tolerance: float = 1e-6, headroom: float = 1.04,
                 moment_particles: int = 1 << 22):
        if backend is None:
            from . import ops as backend
        self.ops = backend
        self.comm = comm
        self.like, self.prior, self.q = like, prior, q
        self.seed = int(seed)
        self.n_moves, self.rho = int(n_moves), float(rho)
        self.target_efficiency, self.beta_tolerance = float(target_efficiency), float(beta_tolerance)
        self.headroom = headroom
        self.moment_particles = int(moment_particles)
        self._stream = 0
        self._peer = False  # mailboxes of the in-kernel all-reduce: False = not tried yet, None = NCCL

    def reset(self):
        """Forget the cached particle count (a different particle set is about to be stepped)."""
        self._n_total = None

    def next_stream(self, count=1):
        s = self._stream
        self._stream += count
        return s

    # -- global reweighting statistics ------------------------------------------
    def global_stats(self, stats_local):
        g = torch.stack([-stats_local[0], stats_local[1], stats_local[2], stats_local[3]])
        self.comm.all_reduce_max(g)
        return torch.stack([-g[0], g[1], g[2], g[3]]).contiguous()

    # -- exact CDF over all ranks -------------------------------------------------
    def global_cdf(self, p, tiles, info, n_total, mark=None):
        """-> (local slice of the global normalised CDF, chain, gathered): ``chain`` = device
        [exact running sum in front of this slice, exact total] (this slice covers (chain[0] / chain[1], ...]),
        ``gathered`` = device [G, 4] {sum p, max p, var-sum, n} of every rank.  No device -> host read."""
        ops, comm = self.ops, self.comm
        r, g = comm.rank, comm.world
        dev = p.device
        gathered = comm.all_gather(info)                     # [G, 4] = {sum p, max p, var-sum, n}
        carry = torch.zeros(4, dtype=torch.float64, device=dev)
        carry[3] = float(n_total)
        if r > 0:                                            # approximate offset / running max of the ranks in front
            carry[0] = gathered[:r, 0].sum()
            carry[1] = gathered[:r, 1].max()
        ops.cdf_plan_aggregate(p, tiles, carry)
        if mark:
            mark("cdf_local")
        if self._peer:                                       # carry chain + total through the peers' mailboxes
            cdf, chain = ops.cdf_resolve_write_p2p(p, self._peer.desc)
            return cdf, chain, gathered
        if r > 0:                                            # exact carry chain: rank r-1 -> r
            inc = torch.empty(2, dtype=torch.float64, device=dev)
            comm.recv(inc, r - 1)
            carry[2] = inc[0]
            carry[1] = torch.maximum(carry[1], inc[1])
        carry_out = ops.cdf_resolve(p, carry)
        if r < g - 1:
            comm.send(carry_out, r + 1)
        total = carry_out[:1].clone()
        comm.broadcast(total, g - 1)
        cdf = ops.cdf_write(p, total)
        return cdf, torch.stack([carry[2], total[0]]), gathered

    # -- owner-computes multinomial resampling at sorted uniforms ----------------------
    def sorted_ancestors(self, cdf, chain, n_total, mark=None):
        """-> (idx, bounds): the ancestors (local indices, ascending) of the slots this rank owns, and the slot
        boundaries J_0 .. J_G of all ranks (host ints; the one device -> host read of the resampling)."""
        ops, comm = self.ops, self.comm
        mark = mark or (lambda name: None)
        r, g = comm.rank, comm.world
        dev = cdf.device
        stream_u = self.next_stream()
        tiles_total = ops.sorted_tiles(n_total)
        per = -(-tiles_total // g)                               # tiles generated per rank (ceil)
        t_lo = min(r * per, tiles_total)
        sp = ops.SortedSpacings(self.seed, stream_u, n_total, dev, j0=t_lo * ops.SPACING_TILE,
                                n_tiles=min(per, tiles_total - t_lo))
        if g > 1:
            sums = torch.zeros(per, dtype=torch.int64, device=dev)
            sums[:sp.n_tiles] = sp.tile_sums
            sp.scan(comm.all_gather(sums).reshape(-1)[:tiles_total].contiguous())
        else:
            sp.scan()
        lo_dev = (chain[0] / chain[1]).reshape(1) if r > 0 else torch.zeros(1, dtype=torch.float64, device=dev)
        lo_all = comm.all_gather(lo_dev).reshape(-1) if g > 1 else lo_dev
        bounds = ops.sorted_bounds(sp, lo_all, n_total)          # host ints J_0 .. J_G: rank s owns slots [J_s, J_s+1)
        if bounds is None:
            raise RuntimeError("exact CDF: a peer GPU did not hand over its running sum within 20 s")
        mark("select")
        j_lo, j_hi = bounds[r], bounds[r + 1]
        idx = sp.window(j_lo, j_hi).search(cdf, j_lo, j_hi)
        mark("resample")
        return idx, bounds

    @staticmethod
    def slice_bounds(chain, cdf, rank, world):
        """(lo, hi) host floats of the slice (one device -> host read; tests and diagnostics)."""
        c_in, th, c_last = torch.stack([chain[0], chain[1], cdf[-1]]).tolist()
        return (c_in / th if rank > 0 else 0.0), (1.0 if rank == world - 1 else c_last)

    # -- one tempering step ---------------------------------------------------------
    def step(self, shard: Shard, beta_prev: float, u_all=None, mark=None) -> tuple[Shard, StepRecord]:
        """Three device -> host reads per step: the search state (beta), the slot boundaries (launch sizes) and the
        step record at the end."""
        ops, comm = self.ops, self.comm
        mark = mark or (lambda name: None)
        dev = shard.x.device
        n_local = len(shard)
        d, stats_local = ops.reweight(shard.log_l, shard.log_p, shard.log_q)
        stats = self.global_stats(stats_local)
        if self.__dict__.get("_n_total") is None:            # rebalanced shards keep their sizes: counted once per
            counts = torch.tensor([float(n_local)], dtype=torch.float64, device=dev)  # engine (``reset`` recounts);
            comm.all_reduce_sum(counts)                      # every step re-checks the total in its last read
            self._n_total = int(round(float(counts[0])))
        n_total = self._n_total
        if self._peer is False:
            self._peer = PeerMailboxes.create(comm, dev) if dev.type == "cuda" and comm.world > 1 else None
        state = ops.beta_search(d, stats, beta_prev, self.target_efficiency, self.beta_tolerance, n_total=n_total,
                                allreduce=comm.all_reduce_sum,
                                **({} if self._peer is None else {"peer": self._peer.desc}))
        mark("beta_search")
        if state.n_nan > 0 or state.n_inf > 0:               # the all-reduced counters travel with the state
            raise ValueError(f"Log weights contain NaN values for beta={beta_prev}")
        beta = min(max(state.beta_min, beta_prev), 1.0)
        if beta == beta_prev:
            raise RuntimeError(f"Beta did not increase from previous value {beta:.6g}.")
        dbeta = beta - beta_prev
        c = dbeta * state.d_max
        if state.min_valid:
            s1, s2 = state.s1_min, state.s2_min
        else:  # not reachable with max_beta_step = 1: beta* is always an evaluated beta
            sums = ops.ess_sums(d, [dbeta], stats).reshape(-1).clone()
            comm.all_reduce_sum(sums)
            s1, s2 = float(sums[0]), float(sums[1])
        p, result, tiles = ops.smc_weights(d, dbeta, c, s1, n_total=n_total)
        info = torch.stack([result[1], tiles[1::2].max(), result[0], result[0] * 0 + n_local]).contiguous()
        cdf, chain, gathered = self.global_cdf(p, tiles, info, n_total, mark)
        mean_u = s1 / n_total
        mark("cdf")
        idx, bounds = self.sorted_ancestors(cdf, chain, n_total, mark)
        r, g = comm.rank, comm.world
        j_lo, j_hi = bounds[r], bounds[r + 1]
        n_out = j_hi - j_lo
        dims = shard.x.shape[1]
        # reference measure: moments of every (n_total // moment_particles)-th slot - the same set whatever the
        # sharding, spread over the whole genealogy (the ancestors are ascending)
        stride = max(n_total // max(self.moment_particles, 1), 1)
        first = (-j_lo) % stride
        if first < n_out:
            mom = ops.weighted_moments(shard.x, idx=idx, stride=stride, first=first)
        else:
            mom = torch.zeros(1 + 2 * dims, dtype=torch.float64, device=dev)
        comm.all_reduce_sum(mom)
        mu, sigma = ops.moments_to_mu_sigma(mom, dims, floor=1e-150)
        mark("moments")
        prm = ops.make_pcn_params(beta, self.rho, mu, sigma, self.n_moves, seed=self.seed,
                                  stream0=self.next_stream(self.n_moves) + 0x40000000, index0=j_lo,
                                  total_accepts_only=True)
        # target partition: slot order, N/G slots per rank.  The sweep writes slot j to row margin + (j - A_r) of
        # the new buffers, so the slots that stay are already in place; the slivers sit in the margins.
        base, rem = divmod(n_total, g)
        target = [s_ * base + min(s_, rem) for s_ in range(g + 1)]
        a_lo, a_hi = target[r], target[r + 1]
        n_keep = a_hi - a_lo
        margin = 0 if g == 1 else max(int(n_keep * (self.headroom - 1.0)), 0) + 4096
        off = margin + (j_lo - a_lo)
        in_place = off >= 0 and off + n_out <= n_keep + 2 * margin
        bufs = self._output_buffers(n_keep + 2 * margin, dims, dev, shard)
        if in_place:
            src, src_off = bufs, off
        else:  # an imbalance beyond the margins: sweep into a scratch set, every kept row is copied once
            src = (torch.empty(max(n_out, 1), dims, dtype=torch.float64, device=dev),
                   *(torch.empty(max(n_out, 1), dtype=torch.float64, device=dev) for _ in range(3)))
            src_off = 0
        if n_out > 0:
            out = tuple(t[src_off:src_off + n_out] for t in src)
            _, _, _, _, accepts = ops.pcn_sweep(shard.x, shard.log_l, shard.log_p, shard.log_q, idx, self.like,
                                                self.prior, self.q, prm, out=out)
            acc = accepts.sum().reshape(1).to(torch.float64)
        else:
            acc = torch.zeros(1, dtype=torch.float64, device=dev)
        mark("pcn_sweep")
        self.exchange(src, src_off, j_lo, j_hi, bufs, margin, bounds, target)
        mark("exchange")
        comm.all_reduce_sum(acc)
        acc_sum, var_sum, n_seen = torch.stack([acc[0], gathered[:, 2].sum(), gathered[:, 3].sum()]).tolist()  # the step's device -> host read
        if int(round(n_seen)) != n_total:                    # the same on every rank: everybody raises
            self._n_total = None
            raise RuntimeError(f"the shards hold {int(round(n_seen))} particles, the engine counted {n_total}: "
                               "call reset() before stepping a different particle set")
        rec = StepRecord(
            beta=beta, ess=s1 * s1 / s2, ess_target=state.eff_beta_max * n_total,
            log_norm_ratio=c + math.log(s1) - math.log(n_total),
            log_norm_ratio_var=var_sum / n_total / (n_total * mean_u * mean_u) if mean_u != 0 else math.nan,
            acceptance=acc_sum / max(self.n_moves * n_total, 1), n_total=n_total, n_local_out=n_out,
            n_beta_evals=state.n_evals)
        x, ll, lp, lq = (t[margin:margin + n_keep] for t in bufs)
        return Shard(x, ll, lp, lq, a_lo), rec

    def _output_buffers(self, rows, dims, dev, shard):
        """Two persistent buffer sets, used alternately (the set that holds the input shard is never the one
        written): a Shard returned by ``step`` stays valid until the second next step.  Fresh allocations per
        step would be tied to the NCCL stream by the exchange and come back late, so the caching allocator keeps
        falling back to cudaMalloc / cudaFree of ~14 GB blocks (seen as 40 - 290 ms stalls at 2 GPUs)."""
        sets = self.__dict__.setdefault("_buffer_sets", {})
        key = (rows, dims, str(dev))
        pair = sets.get(key)
        if pair is None:
            sets.clear()
            pair = sets[key] = [tuple([torch.empty(rows, dims, dtype=torch.float64, device=dev)] +
                                      [torch.empty(rows, dtype=torch.float64, device=dev) for _ in range(3)])
                                for _ in range(2)]
            self._buffer_turn = 0
            self._warm_exchange(pair[0], dev)
        held = shard.x.untyped_storage().data_ptr()
        for _ in range(2):
            bufs = pair[self._buffer_turn]
            self._buffer_turn ^= 1
            if bufs[0].untyped_storage().data_ptr() != held:
                return bufs
        raise RuntimeError("unreachable: both buffer sets hold the input shard")

    def _warm_exchange(self, bufs, dev, rows: int = 1 << 16):
        """NCCL connects its send / recv channels lazily, per peer and per message-size class (seen: a 110 ms stall
        the first time a sliver is larger than every earlier one).  One exchange of `rows` rows with every peer,
        before the first step, connects them all."""
        comm = self.comm
        if comm.world == 1 or dev.type != "cuda":
            return
        n = min(rows, bufs[0].shape[0] // 2)
        ops_ = []
        for s in range(comm.world):
            if s != comm.rank:
                ops_ += [comm.dist.P2POp(comm.dist.isend, bufs[0][:n], s, comm.group),
                         comm.dist.P2POp(comm.dist.irecv, bufs[0][n:2 * n], s, comm.group)]
        for w in comm.dist.batch_isend_irecv(ops_):
            w.wait()

    def exchange(self, src, src_off, j_lo, j_hi, dst, margin, bounds, target):
        """Shard rebalancing: this rank produced the slots [j_lo, j_hi) (rows ``src[src_off:]``) and keeps the slots
        [target[r], target[r+1]) (rows ``dst[margin:]``).  Slot ranges that cross a shard boundary are sent to
        their target rank and received straight into their final rows: one grouped send/recv, nothing else moves
        (``src is dst``: the rows that stay were written in place by the sweep)."""
        comm = self.comm
        r, g = comm.rank, comm.world
        a_lo, a_hi = target[r], target[r + 1]
        k_lo, k_hi = max(j_lo, a_lo), min(j_hi, a_hi)
        if src is not dst and k_hi > k_lo:
            for s_t, d_t in zip(src, dst):
                d_t[margin + k_lo - a_lo:margin + k_hi - a_lo].copy_(s_t[src_off + k_lo - j_lo:src_off + k_hi - j_lo])
        if g == 1:
            return
        ops_ = []
        for s in range(g):
            if s == r:
                continue
            lo_s, hi_s = max(j_lo, target[s]), min(j_hi, target[s + 1])          # mine, theirs to keep
            if hi_s > lo_s:
                ops_ += [comm.dist.P2POp(comm.dist.isend, t[src_off + lo_s - j_lo:src_off + hi_s - j_lo], s, comm.group)
                         for t in src]
            lo_r, hi_r = max(bounds[s], a_lo), min(bounds[s + 1], a_hi)          # theirs, mine to keep
            if hi_r > lo_r:
                ops_ += [comm.dist.P2POp(comm.dist.irecv, t[margin + lo_r - a_lo:margin + hi_r - a_lo], s, comm.group)
                         for t in dst]
        if ops_:
            for w in comm.dist.batch_isend_irecv(ops_):
                w.wait()


def run_tempering(engine: "ShardedSMC", shard: "Shard", beta: float = 0.0, max_steps: int = 10_000):
    """The tempering loop (reference ``smc/base.py:401-454``) over sharded particles: ``ShardedSMC.step`` until
    beta reaches 1.  Every rank takes the same decisions (they all see the all-reduced sums), so no rank can
    leave the loop early.  -> (final shard, [StepRecord], log evidence, log evidence error)"""
    records = []
    while beta < 1.0 and len(records) < max_steps:
        shard, rec = engine.step(shard, beta)
        beta = rec.beta
        records.append(rec)
    log_z = float(sum(r.log_norm_ratio for r in records))
    err = math.sqrt(float(sum(r.log_norm_ratio_var for r in records)))
    return shard, records, log_z, err


def parity_check(comm: Comm, dev, n_total: int = 1_000_003, dims: int = 32, moves: int = 4, seed: int = 4242) -> dict:
    """One sharded tempering step on all ranks against the same step on rank 0 alone (same global particles, same
    Philox streams): beta identical, every particle of the rebalanced global array compared with the single-GPU
    array row by row.  Collective: every rank must call it.  -> dict (filled on rank 0)."""
    from . import ops
    from . import targets as tg

    r, g = comm.rank, comm.world
    prob = tg.rosenbrock_problem(dims, device=dev)
    base, rem = divmod(n_total, g)
    n_local, start = base + (1 if r < rem else 0), r * base + min(r, rem)
    x0 = ops.philox_normal(seed, 0, n_local, dims, dev, mean=prob["q"].mean, sigma=prob["q"].sigma, index0=start)
    shard = Shard(x0, prob["like"].log_prob(x0), prob["prior"].log_prob(x0), prob["q"].log_prob(x0), start)
    eng = ShardedSMC(comm, prob["like"], prob["prior"], prob["q"], seed=seed + 1, n_moves=moves, rho=0.5)
    new, rec = eng.step(shard, 0.0)
    cap = base + (1 if rem else 0)
    pad = lambda t: torch.cat([t, t.new_zeros((cap - t.shape[0],) + tuple(t.shape[1:]))])  # noqa: E731
    xs, lls = comm.all_gather(pad(new.x)), comm.all_gather(pad(new.log_l))
    produced = comm.all_gather(torch.tensor([rec.n_local_out], dtype=torch.int64, device=dev)).reshape(-1).tolist()
    out = {}
    if r == 0:
        counts = [base + (1 if i < rem else 0) for i in range(g)]
        x = torch.cat([xs[i, :c] for i, c in enumerate(counts)])
        ll = torch.cat([lls[i, :c] for i, c in enumerate(counts)])
        xg = ops.philox_normal(seed, 0, n_total, dims, dev, mean=prob["q"].mean, sigma=prob["q"].sigma)
        full = Shard(xg, prob["like"].log_prob(xg), prob["prior"].log_prob(xg), prob["q"].log_prob(xg))
        solo = ShardedSMC(SoloComm(), prob["like"], prob["prior"], prob["q"], seed=seed + 1, n_moves=moves, rho=0.5)
        ref, rrec = solo.step(full, 0.0)
        dx = (x - ref.x).abs().max(dim=1).values
        out = {"ranks": g, "particles": n_total, "dims": dims, "moves": moves, "beta_identical": rec.beta == rrec.beta,
               "beta": rec.beta, "ess_rel_diff": abs(rec.ess - rrec.ess) / rrec.ess,
               "log_norm_ratio_rel_diff": abs(rec.log_norm_ratio - rrec.log_norm_ratio) / abs(rrec.log_norm_ratio),
               "acceptance_diff": abs(rec.acceptance - rrec.acceptance), "max_abs_dx": float(dx.max()),
               "max_abs_dlogl": float((ll - ref.log_l).abs().max()), "produced_per_rank": produced,
               "ok": bool(rec.beta == rrec.beta and float(dx.max()) < 1e-9 and sum(produced) == n_total)}
    return out


def _host_buffer_leg(args, comm, engine, shard0, n_total, dev, unit):
    """The sharded step with HOST buffers: every rank uploads its slice from pinned host memory over its own
    PCIe link, runs ``ShardedSMC.step`` and downloads its offspring, consecutive steps software-pipelined as in
    ``bench.py::run_e2e``; device-timed, max over ranks."""
    import psutil

    dims = shard0.x.shape[1]
    row_bytes = 8 * (dims + 3)
    n_local = len(shard0)
    budget = int(0.2 * psutil.virtual_memory().available / comm.world / (2 * row_bytes))
    free_dev, _ = torch.cuda.mem_get_info(dev)
    m = max(1, min(n_local, budget, int(0.6 * free_dev / (4.6 * row_bytes))))  # two input sets + two output sets
    neg = torch.tensor([-m / n_local], dtype=torch.float64, device=dev)
    comm.all_reduce_max(neg)  # every rank uses the smallest fraction, so slices stay comparable
    m = max(1, int(n_local * float(-neg[0])))
    src = (shard0.x[:m], shard0.log_l[:m], shard0.log_p[:m], shard0.log_q[:m])
    host_in = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in src]
    for h, t in zip(host_in, src):
        h.copy_(t)
    cap = int(m * 1.2) + 4096
    host_out = [torch.empty((cap, dims), dtype=torch.float64).pin_memory()] + [
        torch.empty(cap, dtype=torch.float64).pin_memory() for _ in range(3)]
    dev_in = [[torch.empty_like(h, device=dev) for h in host_in] for _ in range(2)]
    stream = torch.cuda.current_stream()
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    comp_done, d2h_done = [None, None], [None, None]
    moved = [0, 0]
    n_chunks = 16  # row chunks per copy: the step's small device -> host reads do not queue behind a multi-GB copy

    def chunked_copy(dst, src_):
        rows = dst.shape[0]
        step_rows = max((rows + n_chunks - 1) // n_chunks, 1) if dst.ndim == 2 else max(rows, 1)
        for r0 in range(0, rows, step_rows):
            dst[r0:r0 + step_rows].copy_(src_[r0:r0 + step_rows], non_blocking=True)

    def upload(k):
        b_ = k % 2
        with torch.cuda.stream(s_h2d):
            if comp_done[b_] is not None:
                s_h2d.wait_event(comp_done[b_])  # step k-2 has finished reading this input set
            for d, h in zip(dev_in[b_], host_in):
                chunked_copy(d, h)
            e = torch.cuda.Event()
            e.record(s_h2d)
        moved[0] += m * row_bytes
        return e

    def run(n_steps):
        """Software pipeline over three streams: the upload of step k+1 and the download of step k-1 overlap step k
        (PCIe is full duplex); the engine's two output buffer sets alternate, so step k+2 waits for download k."""
        up = upload(0)
        for k in range(n_steps):
            b_ = k % 2
            nxt = upload(k + 1) if k + 1 < n_steps else None
            stream.wait_event(up)
            if d2h_done[b_] is not None:
                stream.wait_event(d2h_done[b_])
            new, _ = engine.step(Shard(*dev_in[b_]), 0.0)
            comp_done[b_] = torch.cuda.Event()
            comp_done[b_].record(stream)
            kk = min(len(new), cap)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(comp_done[b_])
                for h, t in zip(host_out, (new.x, new.log_l, new.log_p, new.log_q)):
                    chunked_copy(h[:kk], t[:kk])
                d2h_done[b_] = torch.cuda.Event()
                d2h_done[b_].record(s_d2h)
            moved[1] += kk * row_bytes
            up = nxt
        stream.wait_stream(s_d2h)

    steps = max(2, min(args.steps, 10))
    engine.reset()  # a different particle count than the device-resident leg
    run(1)
    torch.cuda.synchronize()
    comm.barrier()
    moved[0] = moved[1] = 0
    comp_done[:] = [None, None]
    d2h_done[:] = [None, None]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    s_h2d.wait_event(a)
    run(steps)
    b.record(stream)
    torch.cuda.synchronize()
    comm.barrier()
    t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
    comm.all_reduce_max(t)
    tot = torch.tensor([float(m), moved[0] / steps, moved[1] / steps], dtype=torch.float64, device=dev)
    comm.all_reduce_sum(tot)
    ms = float(t[0])
    engine.reset()
    return {"value": float(tot[0]) / (ms * 1e-3), "unit": unit, "h2d_bytes_per_step": int(tot[1]),
            "d2h_bytes_per_step": int(tot[2]), "particles": int(tot[0]), "ms_per_step": ms, "steps": steps,
            "api": "ShardedSMC.step on pinned host buffers, one PCIe link per rank; upload of step k+1 and download of "
                   "step k-1 overlap step k (three streams, two buffer sets, chunked copies); timed from the first upload "
                   "to the last download, max over ranks"}


# ---------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): launched by torch.distributed.run, one rank per GPU
# ---------------------------------------------------------------------------------
def run_bench(args, rank, world, dev, metric, unit, workload_config, bytes_per_particle, measured_peaks, ClockSampler,
              cpu_baseline, other_configs=None):
    import json

    import torch.distributed as dist

    from . import _lib, ops
    from . import targets as tg

    dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    dims, moves = args.dims, args.moves
    n_total = args.particles if args.scaling == "strong" else args.particles * world
    base = n_total // world
    n_local = base + (1 if rank < n_total % world else 0)
    start = rank * base + min(rank, n_total % world)
    prob = tg.rosenbrock_problem(dims, device=dev)
    seed = 20251104
    # P0: the same global particle set as the single-GPU run, this rank's contiguous slice
    x0 = ops.philox_normal(seed, 0, n_local, dims, dev, mean=prob["q"].mean, sigma=prob["q"].sigma, index0=start)
    shard0 = Shard(x0, prob["like"].log_prob(x0), prob["prior"].log_prob(x0), prob["q"].log_prob(x0))
    engine = ShardedSMC(comm, prob["like"], prob["prior"], prob["q"], seed=seed + 1, n_moves=moves, rho=0.5)
    phases = ["beta_search", "cdf_local", "cdf", "select", "resample", "moments", "pcn_sweep", "exchange"]
    stream = torch.cuda.current_stream()

    def one_step(record=None):
        evs = []

        def mark(name):
            if record is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                evs.append((name, e))

        mark("start")
        new, rec = engine.step(shard0, 0.0, mark=mark)
        if record is not None:
            record.append(evs)
        return new, rec

    parity = parity_check(comm, dev)  # the N-rank step against the 1-GPU step, before anything is timed
    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    comm.barrier()
    clocks = ClockSampler(dev.index) if rank == 0 else None  # one sampler thread per box, not per rank
    if clocks:
        clocks.start()
    records = []
    launches0 = _lib.launch_count
    import gc

    gc.collect()
    gc.disable()  # a cyclic-GC pause on one rank stalls every rank at the next collective
    torch.cuda.synchronize()
    comm.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(args.steps):
        new, rec = one_step(records)
    b.record(stream)
    torch.cuda.synchronize()
    comm.barrier()
    gc.enable()
    launches = _lib.launch_count - launches0
    clock_info = clocks.stop() if clocks else None
    ms = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device=dev)
    comm.all_reduce_max(ms)                                  # max over ranks, device-timed
    ms_per_step = float(ms[0])
    phase_ms = {k: 0.0 for k in phases}
    for evs in records:
        for (_, e0), (name, e1) in zip(evs[:-1], evs[1:]):
            phase_ms[name] += e0.elapsed_time(e1) / len(records)
    pm = torch.tensor([phase_ms[k] for k in phases], dtype=torch.float64, device=dev)
    per_rank = comm.all_gather(pm).cpu().tolist()  # [rank][phase]: imbalance shows up here
    sm = torch.tensor([evs[0][1].elapsed_time(evs[-1][1]) for evs in records], dtype=torch.float64, device=dev)
    step_ms_per_rank = comm.all_gather(sm).cpu().tolist()  # [rank][step]: a straggling step shows up here
    comm.all_reduce_max(pm)
    phase_ms = {k: float(v) for k, v in zip(phases, pm)}
    e2e = None
    if not getattr(args, "no_e2e", False):
        e2e = _host_buffer_leg(args, comm, engine, shard0, n_total, dev, unit)
    others = None
    if other_configs is not None:  # collective: every rank runs it, rank 0 reports (bench.py: the sharded C5 lines)
        del shard0
        torch.cuda.empty_cache()
        others = other_configs(comm, dev, engine._peer)
    if rank == 0:
        peaks, kind = measured_peaks()
        value = n_total / (ms_per_step * 1e-3)
        sweep_bytes = (n_total / world) * (16 * dims + 48)
        sweep_ms = phase_ms["pcn_sweep"]
        ach = sweep_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
        step_bytes = (n_total / world) * bytes_per_particle(dims)
        step_ach = step_bytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {**workload_config(args, n_total), "parallelism": f"particle-sharded x{world}, owner-computes "
                       "resampling on sorted uniforms, scalar collectives + boundary-crossing slivers peer to peer",
                       "beta_search_allreduce": ("inside the pass kernel over NVLink peer memory (" + engine._peer.how + ")"
                                                 if engine._peer else "NCCL")},
            "clocks": clock_info, "gpu_launches": launches,
            "e2e": e2e,
            "roofline": {"kernel": "pcn_sweep_tpp_kernel<32,1,rosenbrock,gaussian> (per GPU)", "bound": "hbm", "achieved": ach,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                         "peak_source": f"MEASURED_PEAKS.json ({kind})",
                         "note": "issue-bound kernel (Philox, Box-Muller, fp64); HBM fraction reported as the contract asks"},
            "step_roofline": {"achieved": step_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s per GPU",
                              "frac": step_ach / peaks["hbm_gbs"]},
            "cpu_baseline": None, "phase_ms": phase_ms,
            "phase_ms_per_rank": [[round(v, 3) for v in row] for row in per_rank],
            "step_ms_per_rank": [[round(v, 2) for v in row] for row in step_ms_per_rank], "beta_after_step": rec.beta, "acceptance": rec.acceptance,
            "n_local_out": rec.n_local_out, "multi_gpu_parity": parity, "other_configs": others,
        }
        print(json.dumps(line))
    dist.destroy_process_group()
