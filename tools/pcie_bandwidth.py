#!/usr/bin/env python
"""Pinned-memory copy bandwidth of this box for the sizes the host-buffer step moves (MB per direction, GB/s)."""
import torch
dev = torch.device("cuda", 0)
for mb in (1, 3, 6, 12, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, src, dst in (("D2H", d, h), ("H2D", h, d)):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print("{} {:3d} MB: {:.3f} ms  {:.1f} GB/s".format(name, mb, ms, n / ms / 1e6), flush=True)
