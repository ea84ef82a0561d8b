#!/usr/bin/env python3
"""Experiment (GPU box): per-stage CUDA-event times of the resident TUM1 step (orbx_enable_timing) plus a quick parity
check of the same library against the committed golden fixtures of the verbatim reference (tests/golden/ref_*.npz).
Usage: python tools/exp_step.py [batch] [steps] [config]   — prints one line per run; ORBX_* env switches apply."""
import json, os, sys, zlib
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam2_commit_b200 import ORBextractor, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfgname = sys.argv[3] if len(sys.argv) > 3 else "tum1"


def parity():
    bad = []
    for name, seed in (("small", 11), ("tum1", 1), ("kitti", 2), ("euroc", 1000)):
        g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
        c = json.loads(str(g["cfg"]))
        img = synth.synth_image(c["width"], c["height"], seed)
        ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
        kps, desc = ex(img)
        gk, gd = g["keypoints"], g["descriptors"]
        ok = len(kps) == len(gk) and all(np.array_equal(kps[f], gk[f]) for f in ("x", "y", "response", "octave")) and \
            np.array_equal(desc, gd) and np.abs(kps["angle"] - gk["angle"]).max() <= 1e-3
        for l in range(c["nlevels"]):
            ok = ok and zlib.crc32(ex.pyramid_level(l, with_apron=True).tobytes()) == int(g["level_crc"][l])
        if not ok:
            bad.append(f"{name}({len(kps)} vs {len(gk)})")
    return bad


c = synth.CONFIGS[cfgname]; W, H = c["width"], c["height"]
dev = torch.device("cuda", 0)
frames = np.stack([synth.synth_image(W, H, 1 + i) for i in range(32)])
d_imgs = torch.from_numpy(np.ascontiguousarray(frames[np.arange(B) % 32])).to(dev)
ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], device=0)
cap = ex.reserve(W, H, B)
kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=dev); desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
nkp = torch.zeros(B, dtype=torch.int32, device=dev)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)


def step():
    ex.extract_device(d_imgs.data_ptr(), B, W, H, W, W * H, kps.data_ptr(), cap, nkp.data_ptr(), desc.data_ptr(), st.cuda_stream)


for _ in range(3):
    step()
torch.cuda.synchronize()
ex.enable_timing(True)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
sm, n = ex.stage_ms()
bad = parity()
print(json.dumps({"cfg": cfgname, "batch": B, "ms_per_step": round(ms, 4), "frames_per_s": round(B / ms * 1e3),
                  "stage_ms": {k: round(float(v), 4) for k, v in zip(("pyramid", "fast", "quadtree", "describe"), sm)},
                  "kp_mean": float(nkp.float().mean()), "parity_bad": bad,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("ORBX_")}}), flush=True)
