"""GPU parity tests (-m gpu) of the remaining ORBmatcher rows — SearchByProjection(F, vpMapPoints) for
Tracking::SearchLocalPoints, the search half of Fuse, SearchForTriangulation — through the C ABI against the CPU oracle on
the same seeded scenes. Bar: assignments, distances and match counts bit-exact."""
import numpy as np
import pytest

from oracle import binding as ob
from orb_slam2_commit_b200 import ORBVocabulary, fuse_search, search_local_points, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False), (3, True)])
def test_search_local_points_matches_oracle(seed, stereo):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:46-142): RadiusByViewingCos, the level range
    [level-1, level], best / second-best with levels, TH_HIGH, the NN-ratio rule, the stereo gate and the sequential
    "keypoint already holds an observed map point" rule under contention; last writer wins."""
    s = synth.synth_local_points_scene(seed, stereo=stereo)
    for th, nnratio in ((1.0, 0.8), (3.0, 0.8), (5.0, 0.6)):
        n, m = search_local_points(**s, th=th, nnratio=nnratio)
        no, mo = ob.search_local_points(**s, th=th, nnratio=nnratio)
        assert n == no, f"seed {seed} th {th}: nmatches {n} vs {no}"
        assert np.array_equal(m, mo), f"seed {seed} th {th}: {np.count_nonzero(m != mo)} assignments differ"
        assert n > 200
    # every map point observed (each assignment blocks later ones) / none observed (free overwriting)
    for bit in (2, 0):
        s2 = dict(s); s2["query_flags"] = (s["query_flags"] & 1) | bit
        n, m = search_local_points(**s2, th=5.0, nnratio=0.8)
        no, mo = ob.search_local_points(**s2, th=5.0, nnratio=0.8)
        assert n == no and np.array_equal(m, mo)
    # heavy contention: every point projects onto one of 40 places with one of 40 descriptors
    rng = np.random.default_rng(seed)
    s3 = dict(s); q = s["queries"].copy(); pick = rng.integers(0, 40, len(q))
    q["proj_x"] = q["proj_x"][pick]; q["proj_y"] = q["proj_y"][pick]; q["proj_xr"] = q["proj_xr"][pick]; q["level"] = q["level"][pick]
    s3["queries"] = q; s3["query_desc"] = s["query_desc"][pick]
    n, m = search_local_points(**s3, th=5.0, nnratio=0.9)
    no, mo = ob.search_local_points(**s3, th=5.0, nnratio=0.9)
    assert n == no and np.array_equal(m, mo) and n > 40


def test_search_local_points_degenerate_inputs():
    s = synth.synth_local_points_scene(5, n_points=60, n_extra=20)
    e = dict(s); e["query_flags"] = np.zeros_like(s["query_flags"])
    n, m = search_local_points(**e, th=1.0)
    assert n == 0 and (m == -1).all()
    e = dict(s); e["kps"] = s["kps"][:0]; e["desc"] = s["desc"][:0]; e["u_right"] = None; e["occupied"] = None
    n, m = search_local_points(**e, th=1.0)
    assert n == 0 and len(m) == 0
    e = dict(s); e["queries"] = s["queries"][:0]; e["query_desc"] = s["query_desc"][:0]; e["query_flags"] = s["query_flags"][:0]
    n, m = search_local_points(**e, th=1.0)
    assert n == 0 and (m == -1).all()
    # windows that miss the grid entirely
    e = dict(s); q = s["queries"].copy(); q["proj_x"] += 5000; e["queries"] = q
    n, m = search_local_points(**e, th=1.0)
    no, mo = ob.search_local_points(**e, th=1.0)
    assert n == no == 0 and np.array_equal(m, mo)


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False), (3, True)])
def test_fuse_search_matches_oracle(seed, stereo):
    """Search half of ORBmatcher::Fuse (ORBmatcher.cc:918-1092 and the Sim3 form :1094-1236): projection, IsInImage, the
    distance / viewing-angle gates, PredictScale, GetFeaturesInArea, level + chi-square gates, nearest descriptor."""
    s = synth.synth_fuse_scene(seed, stereo=stereo)
    for mode in (0, 1):
        for th in (3.0, 4.0, 10.0):
            n, bi, bd = fuse_search(**s, th=th, mode=mode)
            no, bio, bdo = ob.fuse_search(**s, th=th, mode=mode)
            assert n == no, f"seed {seed} mode {mode} th {th}: nFused {n} vs {no}"
            assert np.array_equal(bi, bio), f"seed {seed} mode {mode} th {th}: {np.count_nonzero(bi != bio)} picks differ"
            assert np.array_equal(bd, bdo)
            assert n > 100 and n == np.count_nonzero(bi >= 0)


def test_predict_scale_thresholds_match_libm_logf():
    """MapPoint::PredictScale (MapPoint.cc:407-422) on the device goes through a threshold table built from the host's logf;
    a fuse job whose points sit at distances swept finely across every level boundary must predict the oracle's level.
    The level decides the window radius and the admissible octaves, so any off-by-one shows up in the picks."""
    s = synth.synth_fuse_scene(7, n_points=4000, n_extra=100)
    Ow = s["Ow3"].astype(np.float64)
    PO = s["pt_xyz"].astype(np.float64) - Ow
    dist = np.linalg.norm(PO, axis=1)
    k = np.arange(len(dist))
    # mfMaxDistance = dist * 1.2^(k / 500): ratios sweep 1.2^0 .. 1.2^8 in 4000 steps, plus exact powers
    md = (dist * np.float64(np.float32(1.2)) ** (k / 500.0)).astype(np.float32)
    s["pt_dist"] = np.stack([np.zeros_like(md), md * 100, md], 1).astype(np.float32)
    for mode in (0, 1):
        n, bi, bd = fuse_search(**s, th=3.0, mode=mode)
        no, bio, bdo = ob.fuse_search(**s, th=3.0, mode=mode)
        assert n == no and np.array_equal(bi, bio) and np.array_equal(bd, bdo)
    lv = [ob.predict_scale(float(m), float(np.float32(d)), s["log_scale_factor"], 8) for m, d in zip(md[::97], dist[::97])]
    assert min(lv) == 0 and max(lv) == 7


def test_fuse_search_degenerate_inputs():
    s = synth.synth_fuse_scene(9, n_points=50, n_extra=10)
    e = dict(s); e["pt_flags"] = np.zeros_like(s["pt_flags"])
    n, bi, bd = fuse_search(**e, th=3.0)
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    e = dict(s); e["kps"] = s["kps"][:0]; e["desc"] = s["desc"][:0]; e["u_right"] = None
    n, bi, bd = fuse_search(**e, th=3.0, mode=1)
    no, bio, bdo = ob.fuse_search(**e, th=3.0, mode=1)
    assert n == no == 0 and np.array_equal(bi, bio) and np.array_equal(bd, bdo)
    e = dict(s)
    for k in ("pt_xyz", "pt_normal", "pt_dist", "pt_desc", "pt_flags"):
        e[k] = s[k][:0]
    n, bi, bd = fuse_search(**e, th=3.0)
    assert n == 0 and len(bi) == 0


@pytest.mark.parametrize("seed,stereo", [(1, False), (2, True), (3, False)])
def test_search_for_triangulation_matches_oracle(seed, stereo):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:738-916): node-aligned candidates, map-point / stereo flags, the
    epipole distance gate, CheckDistEpipolarLine, last-minimum-wins under `dist > bestDist`, rotation histogram."""
    voc = synth.synth_vocabulary(10, 4, 5)
    V = ORBVocabulary(10, 4, *voc); Vo = ob.Vocabulary(10, 4, *voc)
    s = synth.synth_triangulation_scene(voc, seed, stereo=stereo)
    t1, t2 = Vo.transform(s["desc1"], 2), Vo.transform(s["desc2"], 2)
    for only_stereo, ori in ((False, True), (False, False), (True, True)):
        n, m = V.search_for_triangulation(**s, levelsup=2, only_stereo=only_stereo, check_orientation=ori)
        no, mo = ob.search_for_triangulation(t1, t2, **s, only_stereo=only_stereo, check_orientation=ori)
        assert n == no, f"seed {seed} only_stereo {only_stereo} ori {ori}: {n} vs {no}"
        assert np.array_equal(m, mo), f"{np.count_nonzero(m != mo)} pairs differ"
        assert n == np.count_nonzero(m >= 0)
        if not only_stereo:
            assert n > 100
    # equal distances inside a node: duplicate keyframe-2 descriptors so that the LAST of the equal candidates must win
    s2 = dict(s); d2 = s["desc2"].copy(); d2[1::2] = d2[0::2][:len(d2[1::2])]; s2["desc2"] = d2
    k2 = s["kps2"].copy(); k2["x"][1::2] = k2["x"][0::2][:len(k2["x"][1::2])]; k2["y"][1::2] = k2["y"][0::2][:len(k2["y"][1::2])]
    k2["octave"][1::2] = k2["octave"][0::2][:len(k2["octave"][1::2])]; s2["kps2"] = k2
    n, m = V.search_for_triangulation(**s2, levelsup=2)
    no, mo = ob.search_for_triangulation(Vo.transform(s2["desc1"], 2), Vo.transform(s2["desc2"], 2), **s2)
    assert n == no and np.array_equal(m, mo)
    # no map-point flags at all / empty sides
    s3 = dict(s); s3["has_mp1"] = None; s3["has_mp2"] = None
    n, m = V.search_for_triangulation(**s3, levelsup=2)
    no, mo = ob.search_for_triangulation(t1, t2, **s3)
    assert n == no and np.array_equal(m, mo)
    s4 = dict(s); s4["kps2"] = s["kps2"][:0]; s4["desc2"] = s["desc2"][:0]; s4["has_mp2"] = s["has_mp2"][:0]
    s4["u_right2"] = None if s["u_right2"] is None else s["u_right2"][:0]
    n, m = V.search_for_triangulation(**s4, levelsup=2)
    assert n == 0 and (m == -1).all()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_search_by_projection_kf_matches_oracle(seed):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1648-1795, mode 0) and
    SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (:327-440, mode 1): gates, PredictScale, level ranges, the
    sequential "feature already holds a map point" rule under contention, the rotation check of mode 0."""
    from orb_slam2_commit_b200 import search_by_projection_kf
    s = synth.synth_kf_projection_scene(seed)
    for mode, th, md, ori in ((0, 10.0, 100, True), (0, 3.0, 64, False), (0, 10.0, 100, False), (1, 10.0, 50, True), (1, 4.0, 50, True)):
        n, m = search_by_projection_kf(**s, th=th, max_dist=md, mode=mode, check_orientation=ori)
        no, mo = ob.search_by_projection_kf(**s, th=th, max_dist=md, mode=mode, check_orientation=ori)
        assert n == no, f"seed {seed} mode {mode} th {th}: {n} vs {no}"
        assert np.array_equal(m, mo), f"seed {seed} mode {mode} th {th}: {np.count_nonzero(m != mo)} assignments differ"
        assert n > 100 and n == np.count_nonzero(m >= 0)
    # heavy contention: all points are copies of 30 points
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, 30, len(s["pt_flags"]))
    s2 = dict(s)
    for k in ("pt_xyz", "pt_normal", "pt_dist", "pt_desc", "pt_flags", "pt_angle"):
        s2[k] = s[k][pick]
    s2["pt_flags"] = np.ones_like(s2["pt_flags"])
    for mode, md in ((0, 100), (1, 50)):
        n, m = search_by_projection_kf(**s2, th=15.0, max_dist=md, mode=mode)
        no, mo = ob.search_by_projection_kf(**s2, th=15.0, max_dist=md, mode=mode)
        assert n == no and np.array_equal(m, mo)
    # degenerate
    e = dict(s); e["pt_flags"] = np.zeros_like(s["pt_flags"])
    n, m = search_by_projection_kf(**e, th=10.0, max_dist=100, mode=0)
    assert n == 0 and (m == -1).all()
    e = dict(s); e["kps"] = s["kps"][:0]; e["desc"] = s["desc"][:0]; e["occupied"] = None
    n, m = search_by_projection_kf(**e, th=10.0, max_dist=50, mode=1)
    assert n == 0 and len(m) == 0


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_search_by_sim3_matches_oracle(seed):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1238-1487): both projection directions through the similarity and the
    mutual-consistency check."""
    from orb_slam2_commit_b200 import search_by_sim3
    k1, k2, S12, S21, cam, sf, lsf = synth.synth_sim3_scene(seed)
    for th in (7.5, 3.0, 15.0):
        n, m = search_by_sim3(k1, k2, S12, S21, cam, sf, lsf, th)
        no, mo = ob.search_by_sim3(k1, k2, S12, S21, cam, sf, lsf, th)
        assert n == no and np.array_equal(m, mo), f"seed {seed} th {th}: {n} vs {no}, {np.count_nonzero(m != mo)} differ"
        assert n == np.count_nonzero(m >= 0)
    assert n > 50
    e1 = {k: (v[:0] if k != "Tcw12" else v) for k, v in k1.items()}
    n, m = search_by_sim3(e1, k2, S12, S21, cam, sf, lsf, 7.5)
    assert n == 0 and len(m) == 0
    e2 = {k: (v[:0] if k != "Tcw12" else v) for k, v in k2.items()}
    n, m = search_by_sim3(k1, e2, S12, S21, cam, sf, lsf, 7.5)
    assert n == 0 and (m == -1).all()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_search_for_initialization_matches_oracle(seed):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:442-587): level-0 keypoints only, the sequential vMatchedDistance /
    vnMatches21 rule (matches change hands), NN ratio, rotation histogram over every push, vbPrevMatched update."""
    from orb_slam2_commit_b200 import search_for_initialization
    s = synth.synth_initialization_scene(seed)
    for win, nnr, ori in ((100, 0.9, True), (100, 0.9, False), (30, 0.7, True), (1000, 0.99, True)):
        n, m, p = search_for_initialization(**s, window_size=win, nnratio=nnr, check_orientation=ori)
        no, mo, po = ob.search_for_initialization(**s, window_size=win, nnratio=nnr, check_orientation=ori)
        assert n == no, f"seed {seed} window {win}: {n} vs {no}"
        assert np.array_equal(m, mo), f"seed {seed} window {win}: {np.count_nonzero(m != mo)} differ"
        assert np.array_equal(p.view(np.uint32), po.view(np.uint32))
        assert n == np.count_nonzero(m >= 0) and n > 100
    # second call with the updated vbPrevMatched, as Tracking::MonocularInitialization does on the next frame
    s2 = dict(s); s2["prev_matched"] = p
    n, m, p2 = search_for_initialization(**s2, window_size=100)
    no, mo, po2 = ob.search_for_initialization(**s2, window_size=100)
    assert n == no and np.array_equal(m, mo) and np.array_equal(p2, po2)
    # degenerate: no level-0 keypoints in F1 / empty F2
    e = dict(s); k = s["kps1"].copy(); k["octave"] = 1; e["kps1"] = k
    n, m, _ = search_for_initialization(**e)
    assert n == 0 and (m == -1).all()
    e = dict(s); e["kps2"] = s["kps2"][:0]; e["desc2"] = s["desc2"][:0]
    n, m, _ = search_for_initialization(**e)
    assert n == 0 and (m == -1).all()


def test_keyframe_side_matchers_with_non_integer_image_bounds():
    """A distorted camera has non-integer undistorted bounds. KeyFrame keeps them as `const int` (KeyFrame.h:236-239) while
    its grid is the Frame's (float bounds): Fuse, the loop-closing SearchByProjection and SearchBySim3 must build the grid
    with the floats and test IsInImage / the window cells with the truncated values; the Frame-side matchers use the floats
    throughout. (tests/test_matcher_ref.py pins the oracle's handling to the reference itself.)"""
    from orb_slam2_commit_b200 import search_by_projection_kf, search_by_sim3
    bounds = np.array([-3.6, 643.2, -2.7, 482.9], np.float32)
    s = synth.synth_fuse_scene(11)
    s["cam9"] = s["cam9"].copy(); s["cam9"][5:9] = bounds
    for mode in (0, 1):
        n, bi, bd = fuse_search(**s, th=4.0, mode=mode)
        no, bio, bdo = ob.fuse_search(**s, th=4.0, mode=mode)
        assert n == no and np.array_equal(bi, bio) and np.array_equal(bd, bdo) and n > 100
    p = synth.synth_kf_projection_scene(12)
    p["cam9"] = p["cam9"].copy(); p["cam9"][5:9] = bounds
    for mode, md in ((0, 100), (1, 50)):
        n, m = search_by_projection_kf(**p, th=10.0, max_dist=md, mode=mode)
        no, mo = ob.search_by_projection_kf(**p, th=10.0, max_dist=md, mode=mode)
        assert n == no and np.array_equal(m, mo) and n > 100
    k1, k2, S12, S21, cam, sf, lsf = synth.synth_sim3_scene(13)
    cam = cam.copy(); cam[5:9] = bounds
    n, m = search_by_sim3(k1, k2, S12, S21, cam, sf, lsf, 7.5)
    no, mo = ob.search_by_sim3(k1, k2, S12, S21, cam, sf, lsf, 7.5)
    assert n == no and np.array_equal(m, mo) and n > 30
    lp = synth.synth_local_points_scene(14)
    lp["bounds4"] = bounds
    n, m = search_local_points(**lp, th=3.0)
    no, mo = ob.search_local_points(**lp, th=3.0)
    assert n == no and np.array_equal(m, mo) and n > 200


@pytest.mark.parametrize("seed", [1, 2])
def test_is_in_frustum_matches_oracle_and_feeds_search_local_points(seed):
    """Frame::isInFrustum (Frame.cc:315-378) on the GPU: return value and the five mTrack* fields bit-identical to the
    restatement; its output drives SearchByProjection(F, vpMapPoints) exactly as Tracking::SearchLocalPoints chains them."""
    from orb_slam2_commit_b200 import is_in_frustum
    s = synth.synth_fuse_scene(seed, n_points=3000)
    for limit in (0.5, 0.9):
        q, v = is_in_frustum(s["Tcw12"], s["Ow3"], s["cam9"], 8, s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], s["pt_dist"], limit)
        qo, vo = ob.is_in_frustum(s["Tcw12"], s["Ow3"], s["cam9"], 8, s["log_scale_factor"], s["pt_xyz"], s["pt_normal"], s["pt_dist"], limit)
        assert np.array_equal(v, vo) and q[v != 0].tobytes() == qo[vo != 0].tobytes()
        assert 300 < np.count_nonzero(v) < len(v)
    rng = np.random.default_rng(seed)
    flags = v | ((rng.random(len(v)) < 0.8).astype(np.uint8) << 1)
    args = dict(kps=s["kps"], desc=s["desc"], u_right=s["u_right"], occupied=None, bounds4=s["cam9"][5:9], scale_factors=s["scale_factors"],
                queries=q, query_desc=s["pt_desc"], query_flags=flags)
    n, m = search_local_points(**args, th=3.0)
    no, mo = ob.search_local_points(**args, th=3.0)
    assert n == no and np.array_equal(m, mo) and n > 100
    qe, ve = is_in_frustum(s["Tcw12"], s["Ow3"], s["cam9"], 8, s["log_scale_factor"], s["pt_xyz"][:0], s["pt_normal"][:0], s["pt_dist"][:0])
    assert len(qe) == 0 and len(ve) == 0


@pytest.mark.parametrize("seed", list(range(100, 124)))
def test_small_random_scenes_match_oracle(seed):
    """The GPU twin of tests/test_matcher_ref.py::test_small_random_scenes_restatements_equal_reference: many small scenes
    with random thresholds (sparse grids, empty windows, single candidates, heavy contention) through every matcher."""
    from orb_slam2_commit_b200 import (is_in_frustum, search_by_projection_frame, search_by_projection_kf, search_by_sim3,
                                       search_for_initialization)
    rng = np.random.default_rng(seed)
    n = int(rng.integers(8, 260)); extra = int(rng.integers(0, 60))
    bounds = None if seed % 3 else np.array([-3.6, 643.2, -2.7, 482.9], np.float32)
    s = synth.synth_local_points_scene(seed, n_points=n, n_extra=extra, cluster=float(rng.uniform(0, 0.8)), stereo=bool(seed & 1))
    if bounds is not None:
        s["bounds4"] = bounds
    th, nnr = float(rng.choice([1.0, 3.0, 5.0, 20.0])), float(rng.choice([0.6, 0.8, 0.99]))
    a, b = search_local_points(**s, th=th, nnratio=nnr), ob.search_local_points(**s, th=th, nnratio=nnr)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    t = synth.synth_tracking_scene(seed, n_last=n, n_extra=extra, cluster=float(rng.uniform(0, 0.8)), stereo=bool(seed & 2))
    th = float(rng.choice([3.0, 7.0, 15.0, 40.0]))
    for mode in (0, 1, 2):
        a = search_by_projection_frame(**t, th=th, mode=mode, check_orientation=bool(seed & 4))
        b = ob.search_by_projection_frame(**t, th=th, mode=mode, check_orientation=bool(seed & 4))
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
    f = synth.synth_kf_projection_scene(seed, n_points=n, n_extra=extra, cluster=float(rng.uniform(0, 0.8)))
    if bounds is not None:
        f["cam9"] = f["cam9"].copy(); f["cam9"][5:9] = bounds
    inv_s2 = (np.float32(1.0) / (f["scale_factors"] * f["scale_factors"])).astype(np.float32)
    th = float(rng.choice([3.0, 4.0, 10.0, 30.0]))
    fa = {k: v for k, v in f.items() if k not in ("occupied", "pt_angle")}
    for mode in (0, 1):
        a = fuse_search(**fa, u_right=None, inv_level_sigma2=inv_s2, th=th, mode=mode)
        b = ob.fuse_search(**fa, u_right=None, inv_level_sigma2=inv_s2, th=th, mode=mode)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        md = int(rng.choice([50, 64, 100]))
        a = search_by_projection_kf(**f, th=float(int(th)), max_dist=md, mode=mode, check_orientation=bool(seed & 1))
        b = ob.search_by_projection_kf(**f, th=float(int(th)), max_dist=md, mode=mode, check_orientation=bool(seed & 1))
        assert a[0] == b[0] and np.array_equal(a[1], b[1])
    q, v = is_in_frustum(f["Tcw12"], f["Ow3"], f["cam9"], 8, f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], f["pt_dist"], 0.5)
    qo, vo = ob.is_in_frustum(f["Tcw12"], f["Ow3"], f["cam9"], 8, f["log_scale_factor"], f["pt_xyz"], f["pt_normal"], f["pt_dist"], 0.5)
    assert np.array_equal(v, vo) and q[v != 0].tobytes() == qo[vo != 0].tobytes()
    i = synth.synth_initialization_scene(seed, n=max(n, 20), cluster=float(rng.uniform(0, 0.8)))
    win, nnr = int(rng.choice([10, 50, 100, 400])), float(rng.choice([0.7, 0.9, 0.99]))
    a = search_for_initialization(**i, window_size=win, nnratio=nnr, check_orientation=bool(seed & 1))
    b = ob.search_for_initialization(**i, window_size=win, nnratio=nnr, check_orientation=bool(seed & 1))
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    k1, k2, S12, S21, cam, sf, lsf = synth.synth_sim3_scene(seed, n_points=max(n, 30), n_extra=extra)
    a, b = search_by_sim3(k1, k2, S12, S21, cam, sf, lsf, th), ob.search_by_sim3(k1, k2, S12, S21, cam, sf, lsf, th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])


def test_matchers_reproduce_the_reference_fixtures():
    """The CUDA matchers through the C ABI against tests/golden/ref_matchers.npz: outputs of the UNMODIFIED reference code
    (ORBmatcher.cc, Frame::isInFrustum, DBoW2's TemplatedVocabulary.h) on seeded scenes, generated by tools/gen_golden_matchers.py."""
    import orb_slam2_commit_b200 as impl
    import golden_matchers
    S = golden_matchers.scenes()
    V = ORBVocabulary(10, 4, *S["voc"])

    def tri(t):
        return V.search_for_triangulation(**t, levelsup=2)
    golden_matchers.check(impl, V.transform, tri)
