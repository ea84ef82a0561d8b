#!/usr/bin/env python3
"""Experiment (GPU box): config 4 on one GPU — 2048 queries x 1,000,000 train rows through orbx_hamming_top2_device, CUDA-event
time per search and a parity check of the packed result against a second run (determinism) and the host one-shot call on a
slice. Usage: python tools/exp_hamming.py [reps]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam2_commit_b200 import dist as od, synth, hamming_top2
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
train, query = synth.synth_descriptors(1_000_000, 2048, seed=42)
dq = torch.from_numpy(query).cuda(); dt = torch.from_numpy(train).cuda()
for _ in range(3):
    r = od.hamming_top2_single(dq, dt)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    r = od.hamming_top2_single(dq, dt)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
i1, d1, d2 = [x.cpu().numpy() for x in r]
j1, f1, f2 = hamming_top2(query[:64], train[:200_000])
k1, g1, g2 = [x.cpu().numpy() for x in od.hamming_top2_single(dq[:64].contiguous(), dt[:200_000].contiguous())]
print(json.dumps({"ms": round(ms, 4), "pair_distances_per_s": 2048 * 1e6 / (ms * 1e-3), "matched_exact": int((d1 == 0).sum()),
                  "slice_equal_to_host_call": bool(np.array_equal(j1, k1) and np.array_equal(f1, g1) and np.array_equal(f2, g2))}))
