"""CPU-only: the oracle's matcher restatements (oracle/orb_oracle.c) against the UNMODIFIED reference ORBmatcher.cc, compiled
here from /root/reference by oracle/Makefile (target ref_matcher) against the stub classes of oracle/slamshim — every public
function of ORBmatcher, on the same seeded scenes the GPU parity tests use. This is what pins the control flow of the
restatements (and through them the CUDA kernels) to the reference itself. Skipped where oracle/_ref/libmatcher_ref.so is
absent and cannot be built (no reference tree)."""
import numpy as np
import pytest

from oracle import binding as ob
from orb_slam2_commit_b200 import synth

pytestmark = pytest.mark.skipif(ob.matcher_ref() is None, reason="oracle/_ref/libmatcher_ref.so not built (reference tree absent)")

ODD_BOUNDS = (-3.6, 643.2, -2.7, 482.9)     # undistorted-image bounds are not integers for a distorted camera (Frame.cc:508-538)


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False)])
def test_search_local_points_restatement_equals_reference(seed, stereo):
    s = synth.synth_local_points_scene(seed, stereo=stereo)
    for bounds in (None, ODD_BOUNDS):
        if bounds:
            s = dict(s); s["bounds4"] = np.array(bounds, np.float32)
        for th, nnratio in ((1.0, 0.8), (3.0, 0.8), (5.0, 0.6)):
            n, m = ob.search_local_points(**s, th=th, nnratio=nnratio)
            nr, mr = ob.ref_search_local_points(**s, th=th, nnratio=nnratio)
            assert n == nr and np.array_equal(m, mr), (seed, th, n, nr)
            assert n > 200
    for bit in (2, 0):
        s2 = dict(s); s2["query_flags"] = (s["query_flags"] & 1) | bit
        assert ob.search_local_points(**s2, th=5.0)[0] == ob.ref_search_local_points(**s2, th=5.0)[0]
        assert np.array_equal(ob.search_local_points(**s2, th=5.0)[1], ob.ref_search_local_points(**s2, th=5.0)[1])


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False)])
def test_search_by_projection_frame_restatement_equals_reference(seed, stereo):
    s = synth.synth_tracking_scene(seed, stereo=stereo)
    seen = set()
    for mono, tlw_z in ((1, 0.0), (0, 0.0), (0, 5.0), (0, -5.0)):
        for th, ori in ((7.0, True), (15.0, False)):
            nr, mr, mode = ob.ref_search_by_projection_frame(**s, th=th, mono=mono, tlw_z=tlw_z, check_orientation=ori)
            n, m = ob.search_by_projection_frame(**s, th=th, mode=mode, check_orientation=ori)
            assert n == nr and np.array_equal(m, mr), (seed, mono, tlw_z, th, n, nr)
            seen.add(mode)
    assert seen == {0, 1, 2}


def _raw_dist(s, rng):
    """(mfMinDistance, mfMaxDistance) per point for a fuse-type scene; a few points get a range that excludes them"""
    mx = s["pt_dist"][:, 2].copy()
    mn = (mx / s["scale_factors"][-1]).astype(np.float32)
    far = rng.random(len(mx)) < 0.06
    mx[far] *= np.float32(0.4)
    return np.stack([mn, mx], 1).astype(np.float32)


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False)])
def test_fuse_restatement_equals_reference(seed, stereo):
    s = synth.synth_fuse_scene(seed, stereo=stereo)
    raw = _raw_dist(s, np.random.default_rng(seed))
    for bounds in (None, ODD_BOUNDS):
        cam = s["cam9"].copy()
        if bounds:
            cam[5:9] = bounds
        for th in (3.0, 10.0):
            # Fuse(pKF, vpMapPoints, th)
            nr, bir, T, Ow, d3 = ob.ref_fuse(s["kps"], s["desc"], s["u_right"], s["Tcw12"], s["Ow3"], cam, s["scale_factors"], s["inv_level_sigma2"],
                                             s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], raw, s["pt_desc"], s["pt_flags"], th, 0)
            n, bi, _ = ob.fuse_search(s["kps"], s["desc"], s["u_right"], T, Ow, cam, s["scale_factors"], s["inv_level_sigma2"],
                                      s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], d3, s["pt_desc"], s["pt_flags"], th, 0)
            assert n == nr and np.array_equal(bi, bir), (seed, th, n, nr, np.count_nonzero(bi != bir))
            assert n > 100
            # Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) with Scw = 1.7 * [R | t]
            Scw = (np.float32(1.7) * s["Tcw12"]).astype(np.float32)
            nr, bir, T, Ow, d3 = ob.ref_fuse(s["kps"], s["desc"], s["u_right"], Scw, None, cam, s["scale_factors"], s["inv_level_sigma2"],
                                             s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], raw, s["pt_desc"], s["pt_flags"], th, 1)
            n, bi, _ = ob.fuse_search(s["kps"], s["desc"], s["u_right"], T, Ow, cam, s["scale_factors"], s["inv_level_sigma2"],
                                      s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], d3, s["pt_desc"], s["pt_flags"], th, 1)
            assert n == nr and np.array_equal(bi, bir), (seed, th, n, nr, np.count_nonzero(bi != bir))
            assert n > 100


def test_predict_scale_restatement_equals_reference_across_level_boundaries():
    """ratios swept finely across every level boundary: the level decides radius and admissible octaves, so an off-by-one in
    ceil(logf(ratio)/mfLogScaleFactor) changes the picks"""
    s = synth.synth_fuse_scene(7, n_points=4000, n_extra=100)
    dist = np.linalg.norm(s["pt_xyz"].astype(np.float64) - s["Ow3"].astype(np.float64), axis=1)
    md = (dist * np.float64(np.float32(1.2)) ** (np.arange(len(dist)) / 500.0)).astype(np.float32)
    raw = np.stack([np.zeros_like(md), md], 1)
    nr, bir, T, Ow, d3 = ob.ref_fuse(s["kps"], s["desc"], s["u_right"], s["Tcw12"], s["Ow3"], s["cam9"], s["scale_factors"], s["inv_level_sigma2"],
                                     s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], raw, s["pt_desc"], s["pt_flags"], 3.0, 0)
    d3[:, 1] = md * 100                        # the restatement takes the invariance bounds as given: open the upper one as the sweep needs
    raw2 = np.stack([np.zeros_like(md), md], 1)
    n, bi, _ = ob.fuse_search(s["kps"], s["desc"], s["u_right"], T, Ow, s["cam9"], s["scale_factors"], s["inv_level_sigma2"], s["log_scale_factor"],
                              s["pt_xyz"], s["pt_normal"], np.stack([d3[:, 0], np.float32(1.2) * raw2[:, 1], d3[:, 2]], 1), s["pt_desc"], s["pt_flags"], 3.0, 0)
    assert n == nr and np.array_equal(bi, bir)
    assert n > 300


@pytest.mark.parametrize("seed", [1, 2])
def test_search_by_projection_kf_restatement_equals_reference(seed):
    s = synth.synth_kf_projection_scene(seed)
    raw = _raw_dist(s, np.random.default_rng(seed))
    for bounds in (None, ODD_BOUNDS):
        cam = s["cam9"].copy()
        if bounds:
            cam[5:9] = bounds
        for mode, th, md, ori in ((0, 10.0, 100, True), (0, 3.0, 64, False), (1, 10.0, 50, True), (1, 4.0, 50, True)):
            T12 = s["Tcw12"] if mode == 0 else (np.float32(0.8) * s["Tcw12"]).astype(np.float32)
            nr, mr, T, Ow, d3 = ob.ref_search_by_projection_kf(s["kps"], s["desc"], s["occupied"], T12, cam, s["scale_factors"], s["log_scale_factor"],
                                                               s["pt_xyz"], s["pt_normal"], raw, s["pt_desc"], s["pt_flags"], s["pt_angle"], th, md, mode, ori)
            n, m = ob.search_by_projection_kf(s["kps"], s["desc"], s["occupied"], T, Ow, cam, s["scale_factors"], s["log_scale_factor"], s["pt_xyz"],
                                              s["pt_normal"], d3, s["pt_desc"], s["pt_flags"], s["pt_angle"], th, md, mode, ori)
            assert n == nr and np.array_equal(m, mr), (seed, mode, th, n, nr, np.count_nonzero(m != mr))
            assert n > 80


@pytest.mark.parametrize("seed", [1, 2])
def test_search_by_sim3_restatement_equals_reference(seed):
    k1, k2, S12, S21, cam, sf, lsf = synth.synth_sim3_scene(seed)
    for k in (k1, k2):
        k["mp_dist_raw"] = np.stack([k["mp_dist"][:, 2] / sf[-1], k["mp_dist"][:, 2]], 1).astype(np.float32)
    R12 = (S12[:9] / np.float32(1.03)).astype(np.float32); t12 = S12[9:]
    for bounds in (None, ODD_BOUNDS):
        c = cam.copy()
        if bounds:
            c[5:9] = bounds
        for th in (7.5, 3.0):
            nr, mr, S12r, S21r, d1, d2 = ob.ref_search_by_sim3(k1, k2, 1.03, R12, t12, c, sf, lsf, th)
            a = dict(k1); a["mp_dist"] = d1
            b = dict(k2); b["mp_dist"] = d2
            n, m = ob.search_by_sim3(a, b, S12r, S21r, c, sf, lsf, th)
            assert n == nr and np.array_equal(m, mr), (seed, th, n, nr)
    assert n > 30


@pytest.mark.parametrize("seed,stereo", [(1, False), (2, True)])
def test_search_for_triangulation_restatement_equals_reference(seed, stereo):
    voc = synth.synth_vocabulary(10, 4, 5)
    Vo = ob.Vocabulary(10, 4, *voc)
    s = synth.synth_triangulation_scene(voc, seed, stereo=stereo)
    t1, t2 = Vo.transform(s["desc1"], 2), Vo.transform(s["desc2"], 2)
    for only_stereo, ori in ((False, True), (False, False), (True, True)):
        n, m = ob.search_for_triangulation(t1, t2, **s, only_stereo=only_stereo, check_orientation=ori)
        nr, mr = ob.ref_search_for_triangulation(t1, t2, **s, only_stereo=only_stereo, check_orientation=ori)
        assert n == nr and np.array_equal(m, mr), (seed, only_stereo, ori, n, nr)
    # equal distances inside a node (the LAST equal candidate wins in the reference)
    s2 = dict(s); d2 = s["desc2"].copy(); d2[1::2] = d2[0::2][:len(d2[1::2])]; s2["desc2"] = d2
    k2 = s["kps2"].copy()
    for f in ("x", "y", "octave"):
        k2[f][1::2] = k2[f][0::2][:len(k2[f][1::2])]
    s2["kps2"] = k2
    t2b = Vo.transform(d2, 2)
    n, m = ob.search_for_triangulation(t1, t2b, **s2)
    nr, mr = ob.ref_search_for_triangulation(t1, t2b, **s2)
    assert n == nr and np.array_equal(m, mr) and n > 50


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_search_for_initialization_restatement_equals_reference(seed):
    s = synth.synth_initialization_scene(seed)
    for win, nnr, ori in ((100, 0.9, True), (100, 0.9, False), (30, 0.7, True), (1000, 0.99, True)):
        n, m, p = ob.search_for_initialization(**s, window_size=win, nnratio=nnr, check_orientation=ori)
        nr, mr, pr = ob.ref_search_for_initialization(**s, window_size=win, nnratio=nnr, check_orientation=ori)
        assert n == nr and np.array_equal(m, mr) and np.array_equal(p.view(np.uint32), pr.view(np.uint32)), (seed, win, n, nr)
        assert n > 100


@pytest.mark.parametrize("check_ori", [True, False])
def test_search_by_bow_restatements_equal_reference(check_ori):
    voc = synth.synth_vocabulary(10, 4, 11)
    Vo = ob.Vocabulary(10, 4, *voc)
    rng = np.random.default_rng(3)
    n = 1200
    d1 = synth.synth_features_near_words(voc, n, 1, max_flips=10)
    d2 = d1[rng.permutation(n)].copy()
    for i in range(n):
        for b in rng.integers(0, 256, rng.integers(0, 20)):
            d2[i, b >> 3] ^= np.uint8(1 << (b & 7))
    k1 = np.zeros(n, ob.KP_DTYPE); k2 = np.zeros(n, ob.KP_DTYPE)
    k1["angle"] = rng.uniform(0, 360, n); k2["angle"] = (k1["angle"] + rng.normal(8, 6, n)) % 360
    v1 = (rng.random(n) < 0.8).astype(np.uint8); v2 = (rng.random(n) < 0.8).astype(np.uint8)
    t1, t2 = Vo.transform(d1, 2), Vo.transform(d2, 2)
    for nnr in (0.7, 0.9):
        no, mo = ob.search_by_bow(t1, t2, d1, k1["angle"], v1, d2, k2["angle"], nnr, check_ori)
        nr, mr = ob.ref_search_by_bow(t1, t2, k1, d1, v1, k2, d2, None, nnr, check_ori, 0)
        assert no == nr and np.array_equal(mo, mr), (nnr, no, nr)
        no, mo = ob.search_by_bow_kf(t1, t2, d1, k1["angle"], v1, d2, k2["angle"], v2, nnr, check_ori)
        nr, mr = ob.ref_search_by_bow(t1, t2, k1, d1, v1, k2, d2, v2, nnr, check_ori, 1)
        assert no == nr and np.array_equal(mo, mr), (nnr, no, nr)
    assert no > 50


@pytest.mark.parametrize("name,seed,mbf,fx", [("kitti", 2, 386.1448, 718.856), ("euroc", 1000, 47.9064, 435.2047), ("small", 5, 40.0, 200.0)])
def test_compute_stereo_matches_restatement_equals_reference(name, seed, mbf, fx):
    """oc_stereo_match against the verbatim Frame::ComputeStereoMatches (Frame.cc:547-788) on the oracle's own pyramids:
    row bands, Hamming stage, 11x11 SAD slide, parabola, disparity / depth, median cut — mvuRight / mvDepth bit for bit."""
    c = dict(width=320, height=240, nfeatures=500, scale=1.2, nlevels=6, ini_th=20, min_th=7) if name == "small" else synth.CONFIGS[name]
    left, right = synth.synth_stereo_pair(c["width"], c["height"], seed)
    args = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    oL, oR = ob.Extractor(*args), ob.Extractor(*args)
    kl, dl = oL.extract(left); kr, dr = oR.extract(right)
    ur, dp = ob.stereo_match(oL, oR, kl, dl, kr, dr, mbf, fx)
    urr, dpr = ob.ref_stereo_match(oL, oR, kl, dl, kr, dr, mbf, fx)
    assert np.array_equal(ur.view(np.uint32), urr.view(np.uint32)), np.count_nonzero(ur != urr)
    assert np.array_equal(dp.view(np.uint32), dpr.view(np.uint32))
    assert np.count_nonzero(ur >= 0) > 50


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_is_in_frustum_restatement_equals_reference(seed):
    """oc_is_in_frustum against the verbatim Frame::isInFrustum (Frame.cc:315-378): return value and the five mTrack* fields"""
    s = synth.synth_fuse_scene(seed, n_points=3000)
    raw = _raw_dist(s, np.random.default_rng(seed))
    for bounds, limit in ((None, 0.5), (ODD_BOUNDS, 0.5), (None, 0.9)):
        cam = s["cam9"].copy()
        if bounds:
            cam[5:9] = bounds
        qr, vr, d3 = ob.ref_is_in_frustum(s["Tcw12"], s["Ow3"], cam, 8, s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], raw, limit)
        q, v = ob.is_in_frustum(s["Tcw12"], s["Ow3"], cam, 8, s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], d3, limit)
        assert np.array_equal(v, vr)
        assert q[v != 0].tobytes() == qr[vr != 0].tobytes()
        assert 300 < np.count_nonzero(v) < len(v)


@pytest.mark.skipif(ob.bow_ref() is None, reason="oracle/_ref/libbow_ref.so not built (reference tree absent)")
@pytest.mark.parametrize("k,L,seed,early", [(10, 4, 11, 0.03), (10, 4, 12, 0.0), (6, 5, 13, 0.0), (3, 6, 14, 0.02)])
def test_bow_restatement_equals_the_reference_dbow2_header(tmp_path, k, L, seed, early):
    """oc_vocab_transform / oc_bow_score_l1 against the reference's own DBoW2 code: TemplatedVocabulary.h (a header template,
    compiled verbatim as ORBVocabulary.h instantiates it) loads the vocabulary with its own text loader and runs transform()
    and score(); only the leaf functions of DBoW2's absent .cpp files are restated (oracle/bow_glue.cc).
    A leaf above level L - levelsup leaves `nid` uninitialised in the reference (TemplatedVocabulary.h:1151-1156, :1226-1258),
    so the FeatureVector is compared where every leaf lies at or below that level; the BowVector always."""
    voc = synth.synth_vocabulary(k, L, seed, early_leaf=early)
    path = str(tmp_path / "voc.txt")
    ob.write_vocabulary_text(path, k, L, *voc)
    R = ob.RefVocabulary(path); O = ob.Vocabulary(k, L, *voc)
    assert R.nwords == O.nwords > 100
    parent = np.asarray(voc[0]); is_leaf = np.asarray(voc[1])
    depth = np.zeros(len(parent) + 1, np.int64)
    for i, p in enumerate(parent):
        depth[i + 1] = depth[p] + 1
    min_leaf_depth = int(depth[1:][is_leaf != 0].min())
    f = synth.synth_features_near_words(voc, 1500, seed + 1, max_flips=12)
    g = synth.synth_features_near_words(voc, 1300, seed + 2, max_flips=12)
    checked_fv = 0
    for levelsup in range(0, L + 1):
        a, b = R.transform(f, levelsup), O.transform(f, levelsup)
        assert np.array_equal(a["bow_id"], b["bow_id"]) and np.array_equal(a["bow_val"].view(np.uint64), b["bow_val"].view(np.uint64))
        if L - levelsup <= min_leaf_depth:
            for key in ("fv_node", "fv_off", "fv_feat"):
                assert np.array_equal(a[key], b[key]), (levelsup, key)
            checked_fv += 1
    assert checked_fv >= 2
    ta, tb = O.transform(f, 4), O.transform(g, 4)
    assert R.score(ta, tb) == ob.bow_score_l1(ta["bow_id"], ta["bow_val"], tb["bow_id"], tb["bow_val"])
    assert R.score(ta, ta) == ob.bow_score_l1(ta["bow_id"], ta["bow_val"], ta["bow_id"], ta["bow_val"])


@pytest.mark.parametrize("seed", list(range(100, 124)))
def test_small_random_scenes_restatements_equal_reference(seed):
    """Many small scenes (tens to a few hundred features, random thresholds): sparse grids, empty windows, single candidates
    and heavy contention hit the corner cases of every matcher loop."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(8, 260)); extra = int(rng.integers(0, 60))
    bounds = None if seed % 3 else np.array(ODD_BOUNDS, np.float32)
    # SearchByProjection(F, vpMapPoints)
    s = synth.synth_local_points_scene(seed, n_points=n, n_extra=extra, cluster=float(rng.uniform(0, 0.8)), stereo=bool(seed & 1))
    if bounds is not None:
        s["bounds4"] = bounds
    th, nnr = float(rng.choice([1.0, 3.0, 5.0, 20.0])), float(rng.choice([0.6, 0.8, 0.99]))
    a, b = ob.search_local_points(**s, th=th, nnratio=nnr), ob.ref_search_local_points(**s, th=th, nnratio=nnr)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    # SearchByProjection(CurrentFrame, LastFrame)
    t = synth.synth_tracking_scene(seed, n_last=n, n_extra=extra, cluster=float(rng.uniform(0, 0.8)), stereo=bool(seed & 2))
    th = float(rng.choice([3.0, 7.0, 15.0, 40.0]))
    nr, mr, mode = ob.ref_search_by_projection_frame(**t, th=th, mono=int(seed % 4 == 0), tlw_z=float(rng.choice([-5.0, 0.0, 5.0])), check_orientation=bool(seed & 4))
    a = ob.search_by_projection_frame(**t, th=th, mode=mode, check_orientation=bool(seed & 4))
    assert a[0] == nr and np.array_equal(a[1], mr)
    # Fuse, both forms, and the two sequential KeyFrame / Frame projections
    f = synth.synth_kf_projection_scene(seed, n_points=n, n_extra=extra, cluster=float(rng.uniform(0, 0.8)))
    cam = f["cam9"].copy()
    if bounds is not None:
        cam[5:9] = bounds
    raw = _raw_dist(f, rng)
    inv_s2 = (np.float32(1.0) / (f["scale_factors"] * f["scale_factors"])).astype(np.float32)
    th = float(rng.choice([3.0, 4.0, 10.0, 30.0]))
    for mode in (0, 1):
        T12 = f["Tcw12"] if mode == 0 else (np.float32(rng.uniform(0.5, 2.0)) * f["Tcw12"]).astype(np.float32)
        nr, bir, T, Ow, d3 = ob.ref_fuse(f["kps"], f["desc"], None, T12, f["Ow3"], cam, f["scale_factors"], inv_s2, f["log_scale_factor"],
                                         f["pt_xyz"], f["pt_normal"], raw, f["pt_desc"], f["pt_flags"], th, mode)
        n_, bi, _ = ob.fuse_search(f["kps"], f["desc"], None, T, Ow, cam, f["scale_factors"], inv_s2, f["log_scale_factor"], f["pt_xyz"],
                                   f["pt_normal"], d3, f["pt_desc"], f["pt_flags"], th, mode)
        assert n_ == nr and np.array_equal(bi, bir)
        md = int(rng.choice([50, 64, 100])) if mode == 0 else 50          # the loop-closing form always uses TH_LOW (:435)
        nr, mr, T, Ow, d3 = ob.ref_search_by_projection_kf(f["kps"], f["desc"], f["occupied"], T12, cam, f["scale_factors"], f["log_scale_factor"],
                                                           f["pt_xyz"], f["pt_normal"], raw, f["pt_desc"], f["pt_flags"], f["pt_angle"], float(int(th)), md, mode, bool(seed & 1))
        n_, m = ob.search_by_projection_kf(f["kps"], f["desc"], f["occupied"], T, Ow, cam, f["scale_factors"], f["log_scale_factor"], f["pt_xyz"],
                                           f["pt_normal"], d3, f["pt_desc"], f["pt_flags"], f["pt_angle"], float(int(th)), md, mode, bool(seed & 1))
        assert n_ == nr and np.array_equal(m, mr)
    qr, vr, d3 = ob.ref_is_in_frustum(f["Tcw12"], f["Ow3"], cam, 8, f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], raw, 0.5)
    q, v = ob.is_in_frustum(f["Tcw12"], f["Ow3"], cam, 8, f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], d3, 0.5)
    assert np.array_equal(v, vr) and q[v != 0].tobytes() == qr[vr != 0].tobytes()
    # SearchForInitialization
    i = synth.synth_initialization_scene(seed, n=max(n, 20), cluster=float(rng.uniform(0, 0.8)))
    win, nnr = int(rng.choice([10, 50, 100, 400])), float(rng.choice([0.7, 0.9, 0.99]))
    a, b = ob.search_for_initialization(**i, window_size=win, nnratio=nnr, check_orientation=bool(seed & 1)), \
        ob.ref_search_for_initialization(**i, window_size=win, nnratio=nnr, check_orientation=bool(seed & 1))
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_committed_matcher_fixtures_are_what_the_reference_produces_now(tmp_path, monkeypatch):
    """tests/golden/ref_matchers.npz must be reproducible: regenerating it from the reference sources gives the same arrays."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_golden_matchers", os.path.join(root, "tools", "gen_golden_matchers.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    (tmp_path / "tests" / "golden").mkdir(parents=True)
    monkeypatch.setattr(gen, "ROOT", str(tmp_path))
    gen.main()
    new = np.load(tmp_path / "tests" / "golden" / "ref_matchers.npz"); old = np.load(os.path.join(root, "tests", "golden", "ref_matchers.npz"))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert np.asarray(old[k]).tobytes() == np.asarray(new[k]).tobytes(), k
