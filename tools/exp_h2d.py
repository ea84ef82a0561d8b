#!/usr/bin/env python3
"""Experiment (GPU box): pinned host -> device copy bandwidth for the bench's 315 MB frame batch, alone and with a concurrent
D2H stream of the output size; the ceiling of the end-to-end path."""
import time, torch
dev = torch.device("cuda", 0)
h = torch.empty(1024 * 640 * 480, dtype=torch.uint8, pin_memory=True); d = torch.empty_like(h, device=dev)
ho = torch.empty(63410176, dtype=torch.uint8, pin_memory=True); do = torch.empty_like(ho, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(both, chunks):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        n = h.numel() // chunks
        for c in range(chunks):
            with torch.cuda.stream(s1): d[c * n:(c + 1) * n].copy_(h[c * n:(c + 1) * n], non_blocking=True)
        if both:
            with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    return h.numel() / dt / 1e9, dt * 1e3
for both in (False, True):
    for chunks in (1, 16):
        gbs, ms = run(both, chunks)
        print(f"H2D 315 MB in {chunks} copies{' + concurrent 63 MB D2H' if both else ''}: {gbs:.1f} GB/s ({ms:.2f} ms) -> at most {1024 / ms * 1e3:.0f} frames/s", flush=True)
