#!/usr/bin/env python
"""Do small copies on other streams wait behind a large device-to-host copy?  (development helper: how the copy
engines of this GPU take work from different streams)"""
import time

import torch

big_d = torch.ones(512 * 1024 * 1024 // 8 * 4, dtype=torch.float64, device="cuda")      # 2 GiB
big_h = torch.empty(big_d.numel(), dtype=torch.float64).pin_memory()
small_h = torch.ones(64, dtype=torch.float64).pin_memory()
small_d = torch.zeros(64, dtype=torch.float64, device="cuda")
sa, sb, sc, sk = (torch.cuda.Stream() for _ in range(4))
x = torch.zeros(1 << 20, device="cuda")


def probe(label, fn, stream):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(sa):
        big_h.copy_(big_d, non_blocking=True)           # ~37 ms at 55 GB/s
    time.sleep(0.002)
    t1 = time.perf_counter()
    with torch.cuda.stream(stream):
        fn()
    stream.synchronize()
    t2 = time.perf_counter()
    sa.synchronize()
    t3 = time.perf_counter()
    print(f"{label:34s} done {1e3*(t2-t1):7.2f} ms after it was issued; the large copy took {1e3*(t3-t0):6.1f} ms")


for _ in range(2):
    probe("small H2D on another stream", lambda: small_d.copy_(small_h, non_blocking=True), sb)
    probe("small D2H on another stream", lambda: small_h.copy_(small_d, non_blocking=True), sc)
    probe("small kernel on another stream", lambda: x.add_(1.0), sk)
    # the large copy cut into 32 MB pieces: does a small D2H get in between?
    torch.cuda.synchronize()
    n = big_d.numel(); piece = (32 << 20) // 8
    t0 = time.perf_counter()
    with torch.cuda.stream(sa):
        for o in range(0, n, piece):
            big_h[o:o + piece].copy_(big_d[o:o + piece], non_blocking=True)
    t1 = time.perf_counter()
    with torch.cuda.stream(sc):
        small_h.copy_(small_d, non_blocking=True)
    sc.synchronize(); t2 = time.perf_counter(); sa.synchronize(); t3 = time.perf_counter()
    print(f"{'small D2H behind 64 queued pieces':34s} done {1e3*(t2-t1):7.2f} ms after it was issued; the pieces took {1e3*(t3-t0):6.1f} ms")
