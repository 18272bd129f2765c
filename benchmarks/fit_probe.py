#!/usr/bin/env python
"""Setup-path kernels of round 2b under a timer: the device flow fit (one Adam step = ~50 small launches) and the
Student-t EM fit of the tpCN reference measure.  Prints one JSON line each (CUDA events after warm-up)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from aspire_b200 import ops
from aspire_b200.flows import CouplingFlow, _NativeTrainer
from aspire_b200.rng import PhiloxRNG


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


def main():
    dev = torch.device("cuda")
    for dims in (16, 64):
        f = CouplingFlow(dims, device=dev, n_layers=3, rng=PhiloxRNG(1), seed=1)
        x = PhiloxRNG(2).normal(4000, dims, dev).float()
        tr = _NativeTrainer(f, 500)
        ms = timed(lambda: tr.step(x[:500], 1e-3, 5.0), 50)
        ms_loss = timed(lambda: tr.loss(x[:500]), 50)
        print(json.dumps({"bench": "flow_fit_step", "dims": dims, "batch": 500, "layers": 3, "ms_per_step": ms,
                          "ms_validation_loss": ms_loss, "samples_per_s": 500 / ms * 1e3}))
    for n, dims in ((1 << 20, 32), (1 << 20, 64)):
        xs = PhiloxRNG(3).normal(n, dims, dev)
        mu, sg = xs.mean(dim=0), xs.std(dim=0)
        ms = timed(lambda: ops.student_t_fit(xs, mu, sg, n_iter=30), 5)
        print(json.dumps({"bench": "student_t_fit", "rows": n, "dims": dims, "em_iterations": 30, "ms": ms,
                          "gbs": 30 * n * dims * 8 / ms / 1e6}))


if __name__ == "__main__":
    main()
