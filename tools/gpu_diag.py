#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA path with the oracle; prints instead of asserting (GPU box only)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob
from orb_slam2_commit_b200 import ORBextractor, hamming_top2, synth

def main():
    names = sys.argv[1:] or ["small", "tum1", "kitti"]
    for name in names:
        c = dict(width=200, height=150, nfeatures=150, scale=1.2, nlevels=4, ini_th=20, min_th=7) if name == "small" else synth.CONFIGS[name]
        img = synth.synth_image(c["width"], c["height"], 5)
        ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
        orc = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
        t = time.time(); kps, desc = ex(img); tg = time.time() - t
        t = time.time(); kps, desc = ex(img); tg2 = time.time() - t
        ko, do = orc.extract(img)
        print(f"== {name}: gpu {len(kps)} kp ({tg*1e3:.1f} ms first, {tg2*1e3:.2f} ms second), oracle {len(ko)} kp")
        for l in range(c["nlevels"]):
            pg, po = ex.pyramid_level(l, with_apron=True), orc.level(l)
            pe = "ok" if np.array_equal(pg, po) else f"DIFF {np.count_nonzero(pg != po)} px (payload {np.count_nonzero(pg[19:-19,19:-19] != po[19:-19,19:-19])})"
            cg, co = ex.debug_candidates(l), orc.candidates(l)
            ce = "ok" if cg.tobytes() == co.tobytes() else f"DIFF n={len(cg)} vs {len(co)}"
            if ce != "ok":
                sg = set(zip(cg["x"], cg["y"], cg["response"])); so = set(zip(co["x"], co["y"], co["response"]))
                ce += f" set-only-gpu {len(sg - so)} set-only-oracle {len(so - sg)}"
            print(f"   L{l}: pyramid {pe}; candidates {ce}; kp {ex.debug_level_counts()[l]} vs {orc.level_counts()[l]}")
        n = min(len(kps), len(ko))
        if len(kps) == len(ko):
            for f in ("x", "y", "size", "response", "octave", "class_id"):
                bad = np.count_nonzero(kps[f] != ko[f])
                if bad: print(f"   field {f}: {bad} differ (ordered compare)")
            sg = set(zip(kps["octave"], kps["x"], kps["y"])); so = set(zip(ko["octave"], ko["x"], ko["y"]))
            print(f"   keypoint set diff: {len(sg - so)} / {len(so - sg)}")
            da = np.abs(kps["angle"] - ko["angle"]); print(f"   angle max diff {da.max() if n else 0}, bit-identical {np.count_nonzero(kps['angle'].view(np.uint32) == ko['angle'].view(np.uint32))}/{n}")
            if kps.tobytes() == ko.tobytes():
                print(f"   descriptor rows differing: {np.count_nonzero((desc != do).any(axis=1))}")
        else:
            sg = set(zip(kps["octave"], kps["x"], kps["y"])); so = set(zip(ko["octave"], ko["x"], ko["y"]))
            print(f"   COUNT MISMATCH; set diff {len(sg - so)} / {len(so - sg)}")
    train, query = synth.synth_descriptors(100000, 1000)
    i1, d1, d2 = hamming_top2(query, train); j1, e1, e2 = ob.hamming_top2(query, train, nthreads=8)
    print("hamming: idx diff", np.count_nonzero(i1 != j1), "d1 diff", np.count_nonzero(d1 != e1), "d2 diff", np.count_nonzero(d2 != e2))

if __name__ == "__main__":
    main()
