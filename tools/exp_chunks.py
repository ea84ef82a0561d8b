#!/usr/bin/env python3
"""Experiment (GPU box): end-to-end frames/s of orbx_extract_batch (1024 pinned VGA frames per call) against ORBX_CHUNK / ORBX_SLOTS."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_commit_b200 import ORBextractor, api, synth
c = synth.CONFIGS["tum1"]; W, H = c["width"], c["height"]; B = 1024
frames = np.stack([synth.synth_image(W, H, 1 + i) for i in range(32)])
hb_t = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True); hb = hb_t.numpy(); hb[:] = frames[np.arange(B) % 32]
ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], device=0)
cap = ex.reserve(W, H, B)
k_t = torch.empty((B, cap, 28), dtype=torch.uint8, pin_memory=True); d_t = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
n_t = torch.zeros(B, dtype=torch.int32, pin_memory=True)
kp = k_t.numpy().view(api.KP_DTYPE).reshape(B, cap); de = d_t.numpy(); nk = n_t.numpy()
for chunk, slots in ((64, 8), (32, 8), (128, 8), (128, 4), (256, 4), (64, 4), (16, 8), (512, 2)):
    os.environ["ORBX_CHUNK"] = str(chunk); os.environ["ORBX_SLOTS"] = str(slots)
    for _ in range(2): ex.extract_host(hb, kp, de, nk)
    t0 = time.perf_counter()
    for _ in range(8): ex.extract_host(hb, kp, de, nk)
    dt = (time.perf_counter() - t0) / 8
    print(f"chunk {chunk:4d} x {slots} slots: {B / dt:9.0f} frames/s ({dt * 1e3:.2f} ms per call)", flush=True)
