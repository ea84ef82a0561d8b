#!/usr/bin/env python3
"""Experiment (GPU box): resident throughput of the TUM1 step when the batch is split over S extractor instances on S streams."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orb_slam2_commit_b200 import ORBextractor, synth
c = synth.CONFIGS["tum1"]; W, H = c["width"], c["height"]
dev = torch.device("cuda", 0); B = 1024
frames = np.stack([synth.synth_image(W, H, 1 + i) for i in range(32)])
d_imgs = torch.from_numpy(np.ascontiguousarray(frames[np.arange(B) % 32])).to(dev)
for S in (1, 2, 4, 8):
    b = B // S
    exs = [ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], device=0) for _ in range(S)]
    cap = [e.reserve(W, H, b) for e in exs][0]
    kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    nkp = torch.zeros(B, dtype=torch.int32, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    def step():
        for s in range(S):
            exs[s].extract_device(d_imgs[s * b:].data_ptr(), b, W, H, W, W * H, kps[s * b:].data_ptr(), cap, nkp[s * b:].data_ptr(), desc[s * b:].data_ptr(), streams[s].cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(30): step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"S={S}: {B * 30 / dt:.0f} frames/s ({dt / 30 * 1e3:.3f} ms per {B} frames), kp sum {int(nkp.sum())}", flush=True)
    del exs
