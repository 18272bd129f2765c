"""Diagnostic for the in-kernel all-reduce (asp_beta_search_pass_p2p): run under torchrun with 2+ GPUs.
Maps the mailboxes, runs ONE single-level pass and prints what every rank saw."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aspire_b200 import _lib, ops  # noqa: E402
from aspire_b200.distributed import Comm, PeerMailboxes  # noqa: E402


def main():
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    comm = Comm()
    r = comm.rank
    lib = _lib.load()
    n = 100_000
    g = torch.Generator(device="cpu").manual_seed(7 + r)
    d = torch.randn(n, generator=g, dtype=torch.float64).to(dev)
    stats = torch.tensor([-6.0, 6.0, 0.0, 0.0], dtype=torch.float64, device=dev)
    for route in ["symm"]:
        try:
            pm = PeerMailboxes.create(comm, dev)
        except Exception as e:  # noqa: BLE001
            print(f"[{r}] route {route}: create raised {e!r}", flush=True)
            continue
        if pm is None:
            print(f"[{r}] route {route}: not available", flush=True)
            continue
        box = pm._keep[0]
        print(f"[{r}] route {route}: {pm.how} ptrs={[hex(pm.desc.box[i] or 0) for i in range(comm.world)]} "
              f"own={hex(box.data_ptr())}", flush=True)
        state = torch.zeros(16, dtype=torch.float64, device=dev)
        ws = ops.reduce_ws(dev).data_ptr()
        st = torch.cuda.current_stream().cuda_stream
        for it in range(3):
            lib.asp_beta_search_init(state.data_ptr(), 0.0, 0.5, 1e-6, float(n * comm.world), stats.data_ptr(), st)
            rc = lib.asp_beta_search_pass_p2p(d.data_ptr(), n, state.data_ptr(), 0, None, C.byref(pm.desc), ws, st)
            torch.cuda.synchronize()
            out = _lib.BetaState.from_buffer_copy(state.cpu().numpy().tobytes())
            # the NCCL route
            lib.asp_beta_search_init(state.data_ptr(), 0.0, 0.5, 1e-6, float(n * comm.world), stats.data_ptr(), st)
            lib.asp_beta_search_pass(d.data_ptr(), n, state.data_ptr(), 0, ws, st)
            comm.all_reduce_sum(state[7:9])
            ref = state.cpu().tolist()[7:9]
            b = box.cpu().tolist()
            print(f"[{r}] {route} it{it} rc={rc} reserved={out.reserved} evals={out.n_evals} s1={out.s1!r} nccl s1={ref[0]!r} "
                  f"counter={b[256]} flags p0={[b[w * 16 + 15] for w in range(comm.world)]} "
                  f"p1={[b[128 + w * 16 + 15] for w in range(comm.world)]}", flush=True)
        # a straggler: rank 1 launches 2 s late, the others must wait for it inside the kernel
        import time
        lib.asp_beta_search_init(state.data_ptr(), 0.0, 0.5, 1e-6, float(n * comm.world), stats.data_ptr(), st)
        torch.cuda.synchronize()
        dist.barrier()
        if r == 1:
            time.sleep(2.0)
        t0 = time.time()
        lib.asp_beta_search_pass_p2p(d.data_ptr(), n, state.data_ptr(), 0, None, C.byref(pm.desc), ws, st)
        torch.cuda.synchronize()
        out = _lib.BetaState.from_buffer_copy(state.cpu().numpy().tobytes())
        print(f"[{r}] straggler pass: waited {time.time() - t0:.3f} s reserved={out.reserved} s1={out.s1!r}", flush=True)
        # the whole search, both routes, on sizes like the multi-GPU test's
        for n2 in (100_000, 500_001 + r):
            d2 = (torch.randn(n2, generator=g, dtype=torch.float64) * 30.0).to(dev)
            lo = torch.stack([-d2.min(), d2.max()])
            comm.all_reduce_max(lo)
            stats2 = torch.stack([-lo[0], lo[1], lo[0] * 0, lo[0] * 0]).contiguous()
            tot = n2 * comm.world if n2 == 100_000 else sum(500_001 + k for k in range(comm.world))
            a = ops.beta_search(d2, stats2, 0.0, 0.5, 1e-6, n_total=tot, allreduce=comm.all_reduce_sum)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.time()
            try:
                b = ops.beta_search(d2, stats2, 0.0, 0.5, 1e-6, n_total=tot, peer=pm.desc)
                print(f"[{r}] search n={n2}: p2p beta={b.beta_min!r} evals={b.n_evals} nccl beta={a.beta_min!r} "
                      f"evals={a.n_evals} same={bytes(a) == bytes(b)} {time.time() - t0:.3f} s", flush=True)
            except Exception as e:  # noqa: BLE001
                bx = box.cpu().tolist()
                print(f"[{r}] search n={n2}: {e!r} after {time.time() - t0:.3f} s counter={bx[256]} "
                      f"flags p0={[bx[w * 16 + 15] for w in range(comm.world)]} "
                      f"p1={[bx[128 + w * 16 + 15] for w in range(comm.world)]}", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
