"""Shared by tests/test_oracle_golden.py (CPU: the oracle) and tests/test_gpu_matchers.py (GPU: the CUDA path through the C ABI):
checks an implementation of the matcher rows against tests/golden/ref_matchers.npz, the outputs of the UNMODIFIED reference
code on seeded scenes (tools/gen_golden_matchers.py). `impl` is a module-like object with the functions of
orb_slam2_commit_b200 / oracle.binding; `transform(desc, levelsup)` returns the BoW transform of `impl`."""
import os

import numpy as np

from orb_slam2_commit_b200.synth import golden_matcher_scenes as scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_matchers.npz")


def check(impl, transform, search_for_triangulation):
    G = np.load(GOLDEN)
    S = scenes()
    n, m = impl.search_local_points(**S["local"], th=3.0, nnratio=0.8)
    assert n == int(G["local_n"]) and np.array_equal(m, G["local_match"])
    for mode in (0, 1, 2):
        n, m = impl.search_by_projection_frame(**S["track"], th=7.0, mode=mode)
        assert n == int(G[f"track{mode}_n"]) and np.array_equal(m, G[f"track{mode}_match"]), mode
    f = S["kf"]; inv_s2 = (np.float32(1.0) / (f["scale_factors"] * f["scale_factors"])).astype(np.float32)
    for mode in (0, 1):
        n, bi, _ = impl.fuse_search(f["kps"], f["desc"], None, G[f"fuse{mode}_T"], G[f"fuse{mode}_Ow"], f["cam9"], f["scale_factors"], inv_s2,
                                    f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], G["kf_dist3"], f["pt_desc"], f["pt_flags"], 4.0, mode)
        assert n == int(G[f"fuse{mode}_n"]) and np.array_equal(bi, G[f"fuse{mode}_best"]), mode
        n, m = impl.search_by_projection_kf(f["kps"], f["desc"], f["occupied"], G[f"seq{mode}_T"], G[f"seq{mode}_Ow"], f["cam9"], f["scale_factors"],
                                            f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], G["kf_dist3"], f["pt_desc"], f["pt_flags"],
                                            f["pt_angle"], 10.0, 100 if mode == 0 else 50, mode, True)
        assert n == int(G[f"seq{mode}_n"]) and np.array_equal(m, G[f"seq{mode}_match"]), mode
    q, v = impl.is_in_frustum(f["Tcw12"], f["Ow3"], f["cam9"], 8, f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], G["kf_dist3"], 0.5)
    assert np.array_equal(v, G["frustum_in_view"]) and q[v != 0].tobytes() == G["frustum_q"][G["frustum_in_view"] != 0].tobytes()
    t = S["tri"]
    a = transform(t["desc1"], 2)
    for k in ("bow_id", "fv_node", "fv_off", "fv_feat"):
        assert np.array_equal(a[k], G["bow_" + k]), k
    assert np.array_equal(a["bow_val"].view(np.uint64), G["bow_bow_val"].view(np.uint64))
    n, m = search_for_triangulation(t)
    assert n == int(G["tri_n"]) and np.array_equal(m, G["tri_match"])
    n, m, p = impl.search_for_initialization(**S["init"], window_size=100, nnratio=0.9)
    assert n == int(G["init_n"]) and np.array_equal(m, G["init_match"]) and np.array_equal(p.view(np.uint32), G["init_prev"].view(np.uint32))
    k1, k2, _, _, cam, sf, lsf = S["sim3"]
    a1 = dict(k1); a1["mp_dist"] = G["sim3_dist1"]
    a2 = dict(k2); a2["mp_dist"] = G["sim3_dist2"]
    n, m = impl.search_by_sim3(a1, a2, G["sim3_S12"], G["sim3_S21"], cam, sf, lsf, 7.5)
    assert n == int(G["sim3_n"]) and np.array_equal(m, G["sim3_match"])
    assert min(int(G[k]) for k in G.files if k.endswith("_n")) > 50
