#!/usr/bin/env python3
"""Randomised parity sweep (GPU box): N seeded (image size, pyramid settings, thresholds, quota) configurations, each
extracted through the C ABI on cuda:0 and through the oracle on the host; keypoints (order included), responses, octaves,
angles (1e-3 deg), descriptors, every pyramid level with its apron and every blurred level are compared. Every fourth
configuration goes through the batch call with several distinct frames. Settings the reference itself cannot run (a level
below 62 px, a portrait level) are skipped. Usage: python tools/sweep_parity.py [trials] [seed]  — prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam2_commit_b200 import ORBextractor, OrbxError, synth
from oracle import binding as ob

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 77
rng = np.random.default_rng(seed)
ran = skipped = frames = kps_total = 0
bad = []
t0 = time.time()


def compare(kps, desc, ko, do, what):
    if len(kps) != len(ko):
        return f"{what}: {len(kps)} vs {len(ko)} keypoints"
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        if not np.array_equal(kps[f], ko[f]):
            return f"{what}: field {f}"
    if len(kps):
        da = np.abs(kps["angle"].astype(np.float64) - ko["angle"].astype(np.float64))
        if np.minimum(da, 360.0 - da).max() > 1e-3:
            return f"{what}: angle"
    if not np.array_equal(desc, do):
        return f"{what}: descriptors"
    return None


for trial in range(trials):
    w = int(rng.integers(90, 1500)); h = int(rng.integers(70, 900))
    sc = float(rng.choice([1.1, 1.15, 1.2, 1.25, 1.3, 1.5, 1.7, 2.0, 2.5]))
    nl_max = max(1, int(np.log(min(w, h) / 70.0) / np.log(sc)) + 1)          # mostly settings the reference can run (smallest level >= 62 px)
    nl = int(rng.integers(1, min(12, nl_max + 1) + 1))
    nf = int(rng.choice([30, 200, 500, 1000, 2000, 5000])); ini = int(rng.integers(6, 45)); mn = int(rng.integers(1, ini + 1))
    nb = int(rng.integers(2, 6)) if trial % 4 == 3 else 1
    imgs = [synth.synth_image(w, h, 9000 + 7 * trial + i) for i in range(nb)]
    what = f"trial {trial}: {w}x{h} nf={nf} sc={sc} nl={nl} th={ini}/{mn} batch={nb}"
    try:
        ex = ORBextractor(nf, sc, nl, ini, mn)
        if nb == 1:
            res = [ex(imgs[0])]
        else:
            kl, dl = ex.extract_batch(imgs)
            res = list(zip(kl, dl))
    except OrbxError as e:
        if e.code == 2:
            skipped += 1
            continue
        bad.append(f"{what}: {e}")
        continue
    orc = ob.Extractor(nf, sc, nl, ini, mn)
    for i, img in enumerate(imgs):
        ko, do = orc.extract(img)
        kps, desc = res[i]
        err = compare(kps, desc, ko, do, what + f" frame {i}")
        if err is None and nb == 1:
            for l in range(nl):
                if not np.array_equal(ex.pyramid_level(l, with_apron=True), orc.level(l)):
                    err = what + f": pyramid level {l}"; break
                bo = orc.level(l, blurred=True)
                if bo is not None and not np.array_equal(ex.blurred_level(l), bo):
                    err = what + f": blurred level {l}"; break
        if err:
            bad.append(err)
        frames += 1; kps_total += len(ko)
    ran += 1
print(json.dumps({"trials": trials, "seed": seed, "ran": ran, "skipped_unsupported_by_reference": skipped, "frames": frames,
                  "keypoints_compared": kps_total, "mismatches": bad[:20], "n_mismatches": len(bad), "seconds": round(time.time() - t0, 1)}))
sys.exit(1 if bad else 0)
