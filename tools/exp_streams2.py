#!/usr/bin/env python3
"""Experiment (GPU box): the host path's schedule without its copies — 16 launches of 64 resident frames round-robin over 8
extractor instances / streams — against one 1024-frame launch."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_commit_b200 import ORBextractor, synth
c = synth.CONFIGS["tum1"]; W, H = c["width"], c["height"]
dev = torch.device("cuda", 0); B = 1024
frames = np.stack([synth.synth_image(W, H, 1 + i) for i in range(32)])
d_imgs = torch.from_numpy(np.ascontiguousarray(frames[np.arange(B) % 32])).to(dev)
for chunk, S in ((1024, 1), (64, 8), (64, 1), (128, 8), (32, 8)):
    exs = [ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], device=0) for _ in range(S)]
    cap = [e.reserve(W, H, chunk) for e in exs][0]
    kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    nkp = torch.zeros(B, dtype=torch.int32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    def step():
        for k in range(B // chunk):
            s = k % S; o = k * chunk
            exs[s].extract_device(d_imgs[o:].data_ptr(), chunk, W, H, W, W * H, kps[o:].data_ptr(), cap, nkp[o:].data_ptr(), desc[o:].data_ptr(), streams[s].cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"{B // chunk:3d} launches of {chunk:4d} frames over {S} streams: {B / dt:9.0f} frames/s ({dt * 1e3:.3f} ms per {B} frames)", flush=True)
    del exs
