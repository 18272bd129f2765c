#!/usr/bin/env python
"""Host <-> device copy bandwidth of the box with pinned memory: H2D alone, D2H alone, both at once on two streams.
What bounds bench.py's host-buffer leg (`e2e`: 21 GB each way per step).  Prints one JSON line."""
import json

import torch


def timed(fn, iters=3):
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
    gb = 4
    n = gb * (1 << 30) // 8
    h_in = torch.empty(n, dtype=torch.float64, pin_memory=True).fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float64, device="cuda")
    d_out = torch.ones(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        d_in.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_out, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    t1, t2, t3 = timed(h2d), timed(d2h), timed(both)
    size = n * 8 / 1e9
    print(json.dumps({"bench": "pcie_probe", "gb_per_copy": size, "h2d_gbs": size / t1 * 1e3, "d2h_gbs": size / t2 * 1e3,
                      "both_each_gbs": size / t3 * 1e3, "both_total_gbs": 2 * size / t3 * 1e3}))


if __name__ == "__main__":
    main()
