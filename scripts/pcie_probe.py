#!/usr/bin/env python
"""PCIe link probe: pinned host <-> device copy bandwidth (what bounds bench.py's e2e), alone and with
both directions busy, for a few transfer sizes.  Prints one JSON object."""
import json

import torch


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 1000.0)
    return best


def main():
    out = {}
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    for mb in (16, 64, 256, 1024):
        n = mb << 20
        h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
        d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
        up = timed(lambda: d_a.copy_(h_a, non_blocking=True))
        dn = timed(lambda: h_b.copy_(d_b, non_blocking=True))

        def both():
            cur = torch.cuda.current_stream()
            s_up.wait_stream(cur)
            s_dn.wait_stream(cur)
            with torch.cuda.stream(s_up):
                d_a.copy_(h_a, non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_b.copy_(d_b, non_blocking=True)
            cur.wait_stream(s_up)
            cur.wait_stream(s_dn)
        bi = timed(both)
        out["%dMiB" % mb] = {"h2d_gbs": n / up / 1e9, "d2h_gbs": n / dn / 1e9, "both_each_gbs": n / bi / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
