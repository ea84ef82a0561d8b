#!/usr/bin/env python3
"""Writes tests/golden/ref_matchers.npz: outputs of the UNMODIFIED reference matcher code (oracle/_ref/libmatcher_ref.so,
libbow_ref.so; see oracle/matcher_glue.cc, oracle/bow_glue.cc) on seeded synthetic scenes. Run in the container that has
/root/reference; the fixtures travel with the repo and pin oracle and CUDA path where the reference is absent.
Scenes are regenerated from orb_slam2_commit_b200/synth.py by the tests (tests/test_oracle_golden.py, tests/test_gpu_matchers.py);
only outputs, and the quantities the reference derives with cv::Mat arithmetic (Rcw / tcw / Ow from Scw, sR21 / t21), are stored."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob
from orb_slam2_commit_b200 import synth
from orb_slam2_commit_b200.synth import golden_matcher_scenes as scenes


def main():
    S = scenes(); out = {}
    n, m = ob.ref_search_local_points(**S["local"], th=3.0, nnratio=0.8); out["local_n"] = n; out["local_match"] = m
    for mono, z in ((1, 0.0), (0, 5.0), (0, -5.0)):
        n, m, mode = ob.ref_search_by_projection_frame(**S["track"], th=7.0, mono=mono, tlw_z=z)
        out[f"track{mode}_n"] = n; out[f"track{mode}_match"] = m
    f = S["kf"]; inv_s2 = (np.float32(1.0) / (f["scale_factors"] * f["scale_factors"])).astype(np.float32)
    for mode in (0, 1):
        T12 = f["Tcw12"] if mode == 0 else (np.float32(1.3) * f["Tcw12"]).astype(np.float32)
        n, bi, T, Ow, d3 = ob.ref_fuse(f["kps"], f["desc"], None, T12, f["Ow3"], f["cam9"], f["scale_factors"], inv_s2, f["log_scale_factor"],
                                       f["pt_xyz"], f["pt_normal"], S["kf_raw"], f["pt_desc"], f["pt_flags"], 4.0, mode)
        out[f"fuse{mode}_n"] = n; out[f"fuse{mode}_best"] = bi; out[f"fuse{mode}_T"] = T; out[f"fuse{mode}_Ow"] = Ow; out["kf_dist3"] = d3
        n, m, T, Ow, d3 = ob.ref_search_by_projection_kf(f["kps"], f["desc"], f["occupied"], T12, f["cam9"], f["scale_factors"], f["log_scale_factor"],
                                                         f["pt_xyz"], f["pt_normal"], S["kf_raw"], f["pt_desc"], f["pt_flags"], f["pt_angle"], 10.0,
                                                         100 if mode == 0 else 50, mode, True)
        out[f"seq{mode}_n"] = n; out[f"seq{mode}_match"] = m; out[f"seq{mode}_T"] = T; out[f"seq{mode}_Ow"] = Ow
    q, v, _ = ob.ref_is_in_frustum(f["Tcw12"], f["Ow3"], f["cam9"], 8, f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], S["kf_raw"], 0.5)
    out["frustum_q"] = q; out["frustum_in_view"] = v
    Vo = ob.Vocabulary(10, 4, *S["voc"])
    t = S["tri"]; t1, t2 = Vo.transform(t["desc1"], 2), Vo.transform(t["desc2"], 2)
    n, m = ob.ref_search_for_triangulation(t1, t2, **t); out["tri_n"] = n; out["tri_match"] = m
    n, m, p = ob.ref_search_for_initialization(**S["init"], window_size=100, nnratio=0.9); out["init_n"] = n; out["init_match"] = m; out["init_prev"] = p
    k1, k2, S12, S21, cam, sf, lsf = S["sim3"]
    for k in (k1, k2):
        k["mp_dist_raw"] = np.stack([k["mp_dist"][:, 2] / sf[-1], k["mp_dist"][:, 2]], 1).astype(np.float32)
    R12 = (S12[:9] / np.float32(1.03)).astype(np.float32)
    n, m, S12r, S21r, d1, d2 = ob.ref_search_by_sim3(k1, k2, 1.03, R12, S12[9:], cam, sf, lsf, 7.5)
    out.update(sim3_n=n, sim3_match=m, sim3_S12=S12r, sim3_S21=S21r, sim3_dist1=d1, sim3_dist2=d2)
    path = os.path.join(tempfile.mkdtemp(), "voc.txt"); ob.write_vocabulary_text(path, 10, 4, *S["voc"])
    R = ob.RefVocabulary(path); a = R.transform(t["desc1"], 2)
    for kk in ("bow_id", "bow_val", "fv_node", "fv_off", "fv_feat"): out["bow_" + kk] = a[kk]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_matchers.npz"), **out)
    print({k: (v if np.ndim(v) == 0 else np.shape(v)) for k, v in out.items()})


if __name__ == "__main__":
    main()
