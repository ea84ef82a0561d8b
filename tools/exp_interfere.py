#!/usr/bin/env python3
"""Experiment (GPU box): does a continuous pinned H2D / D2H stream slow the resident extraction kernels down?"""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_commit_b200 import ORBextractor, synth
c = synth.CONFIGS["tum1"]; W, H = c["width"], c["height"]
dev = torch.device("cuda", 0); B = 1024
frames = np.stack([synth.synth_image(W, H, 1 + i) for i in range(32)])
d_imgs = torch.from_numpy(np.ascontiguousarray(frames[np.arange(B) % 32])).to(dev)
ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], device=0)
cap = ex.reserve(W, H, B)
kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
nkp = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(device=dev)
h = torch.empty(64 * W * H, dtype=torch.uint8, pin_memory=True); d = torch.empty((8, 64 * W * H), dtype=torch.uint8, device=dev)
ho = torch.empty(4_000_000, dtype=torch.uint8, pin_memory=True); do = torch.empty_like(ho, device=dev)
cs, cs2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
def step(): ex.extract_device(d_imgs.data_ptr(), B, W, H, W, W * H, kps.data_ptr(), cap, nkp.data_ptr(), desc.data_ptr(), st.cuda_stream)
def measure(bg):
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    copied = 0
    with torch.cuda.stream(st): e0.record()
    for i in range(20):
        step()
        if bg:                                    # ~ one step's worth of traffic per step: 16 x 19.7 MB up, 16 x 4 MB down
            for k in range(16):
                with torch.cuda.stream(cs): d[k % 8].copy_(h, non_blocking=True)
                if bg > 1:
                    with torch.cuda.stream(cs2): ho.copy_(do, non_blocking=True)
                copied += h.numel()
    with torch.cuda.stream(st): e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    return ms
for bg, name in ((0, "kernels alone"), (1, "with concurrent H2D (315 MB per step)"), (2, "with concurrent H2D + D2H")):
    ms = measure(bg)
    print(f"{name}: {ms:.3f} ms per 1024 frames ({B / ms * 1e3:.0f} frames/s)", flush=True)
