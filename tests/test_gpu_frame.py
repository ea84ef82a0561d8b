"""Device-resident Frame handle (-m gpu; include/orbx.h orbx_frame_*, Frame.cc:62-123): keypoints / descriptors taken from the
extractor's device results, undistorted on the device, then SearchLocalPoints' matcher against the resident frame — compared
with the oracle (extractor, cv::undistortPoints restatement, ORBmatcher.cc:46-142 restatement) bit for bit."""
import threading

import numpy as np
import pytest

from oracle import binding as ob
from orb_slam2_commit_b200 import Frame, ORBextractor, search_local_points, synth

pytestmark = pytest.mark.gpu


def test_frame_from_extract_undistorts_what_the_extractor_left_in_hbm():
    c = synth.CONFIGS["tum1"]
    img = synth.synth_image(c["width"], c["height"], 31)
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kps, desc = ex(img)
    fr = Frame(ex.reserve(c["width"], c["height"], 1), 2000)
    fr.from_extract(ex, 0, len(kps), synth.TUM1_K4, synth.TUM1_DIST)
    assert len(fr) == len(kps)
    kun, d = fr.keypoints()
    un_o = ob.undistort_points(np.stack([kps["x"], kps["y"]], 1), synth.TUM1_K4, synth.TUM1_DIST)
    assert np.array_equal(kun["x"], un_o[:, 0]) and np.array_equal(kun["y"], un_o[:, 1])       # f64 iteration, bit-exact
    for f in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(kun[f], kps[f])
    assert np.array_equal(d, desc)
    # no distortion: mvKeysUn = mvKeys (Frame.cc:474-478)
    fr.from_extract(ex, 0, len(kps))
    k2, _ = fr.keypoints()
    assert np.array_equal(k2.view(np.uint8), kps.view(np.uint8))


def test_frame_from_a_batch_call_takes_the_right_frame():
    c = synth.CONFIGS["tum1"]
    imgs = np.stack([synth.synth_image(c["width"], c["height"], 40 + i) for i in range(5)])
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kl, dl = ex.extract_batch(list(imgs))
    fr = Frame(ex.reserve(c["width"], c["height"], 5), 16)
    for i in (4, 0, 2):
        kps, desc = kl[i], dl[i]
        fr.from_extract(ex, i, len(kps))
        k, d = fr.keypoints()
        assert np.array_equal(k.view(np.uint8), np.ascontiguousarray(kps).view(np.uint8)) and np.array_equal(d, desc)


@pytest.mark.parametrize("seed,stereo", [(3, True), (4, False), (5, True)])
def test_frame_search_local_points_matches_the_oracle(seed, stereo):
    sc = synth.synth_local_points_scene(seed, n_points=1500, n_extra=400, stereo=stereo)
    n_o, m_o = ob.search_local_points(**sc, th=3.0)
    n_h, m_h = search_local_points(**sc, th=3.0)
    assert n_h == n_o and np.array_equal(m_h, m_o)
    kps, desc = np.ascontiguousarray(sc["kps"]), np.ascontiguousarray(sc["desc"], np.uint8)
    import torch
    d_k = torch.from_numpy(kps.view(np.uint8).reshape(-1, 28).copy()).cuda(); d_d = torch.from_numpy(desc).cuda()
    fr = Frame(len(kps), len(sc["queries"]))
    fr.from_device(d_k.data_ptr(), d_d.data_ptr(), len(kps), stream=torch.cuda.current_stream().cuda_stream)
    fr.set_stereo(sc["u_right"])
    for rep in range(3):                                      # the handle's staging buffers are reused call after call
        n, m = fr.search_local_points(sc["queries"], sc["query_desc"], sc["query_flags"], sc["occupied"], sc["bounds4"],
                                      sc["scale_factors"], 3.0)
        assert n == n_o and np.array_equal(m, m_o)
    # fewer queries through the same handle
    k = len(sc["queries"]) // 3
    sc2 = dict(sc, queries=sc["queries"][:k], query_desc=sc["query_desc"][:k], query_flags=sc["query_flags"][:k])
    n2_o, m2_o = ob.search_local_points(**sc2, th=3.0)
    n2, m2 = fr.search_local_points(sc2["queries"], sc2["query_desc"], sc2["query_flags"], sc["occupied"], sc["bounds4"], sc["scale_factors"], 3.0)
    assert n2 == n2_o and np.array_equal(m2, m2_o)


def test_frames_on_several_host_threads_run_concurrently_and_stay_exact():
    scenes = [synth.synth_local_points_scene(20 + t, n_points=800, n_extra=200) for t in range(4)]
    want = [ob.search_local_points(**sc, th=3.0) for sc in scenes]
    import torch
    bad = []

    def work(t):
        sc = scenes[t]
        kps, desc = np.ascontiguousarray(sc["kps"]), np.ascontiguousarray(sc["desc"], np.uint8)
        d_k = torch.from_numpy(kps.view(np.uint8).reshape(-1, 28).copy()).cuda(); d_d = torch.from_numpy(desc).cuda()
        torch.cuda.synchronize()
        fr = Frame(len(kps), len(sc["queries"]))
        fr.from_device(d_k.data_ptr(), d_d.data_ptr(), len(kps))
        fr.set_stereo(sc["u_right"])
        for _ in range(25):
            n, m = fr.search_local_points(sc["queries"], sc["query_desc"], sc["query_flags"], sc["occupied"], sc["bounds4"], sc["scale_factors"], 3.0)
            if n != want[t][0] or not np.array_equal(m, want[t][1]):
                bad.append(t)
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [x.start() for x in th]; [x.join() for x in th]
    assert not bad
