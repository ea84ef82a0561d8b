"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs and against the committed golden vectors of the verbatim reference (tests/golden/ref_*.npz).
Bars (BASELINE.json north_star): pyramid / keypoint set / octave / response bit-exact, angle within 1e-3 deg,
descriptors bit-exact, match indices and distances bit-exact."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import binding as ob
from orb_slam2_commit_b200 import ORBextractor, ORBmatcher, hamming_top2, stereo_hamming, stereo_match, synth

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = 1e-3


def _cfg(name):
    if name == "small":
        return dict(width=200, height=150, nfeatures=150, scale=1.2, nlevels=4, ini_th=20, min_th=7)
    return synth.CONFIGS[name]


def _check_against(kps, desc, kps_o, desc_o, what):
    assert len(kps) == len(kps_o), f"{what}: {len(kps)} vs {len(kps_o)} keypoints"
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], kps_o[f]), f"{what}: field {f} differs (order included)"
    da = np.abs(kps["angle"].astype(np.float64) - kps_o["angle"].astype(np.float64))
    da = np.minimum(da, 360.0 - da)
    assert da.max() <= ANGLE_TOL_DEG, f"{what}: angle off by {da.max()}"
    same_angle = kps["angle"].view(np.uint32) == kps_o["angle"].view(np.uint32)
    rows_equal = (desc == desc_o).all(axis=1)
    assert rows_equal[same_angle].all(), f"{what}: descriptor rows differ at identical angles"
    assert rows_equal.all(), f"{what}: {np.count_nonzero(~rows_equal)} descriptor rows differ"


@pytest.mark.parametrize("name,seed", [("small", 11), ("tum1", 1), ("kitti", 2), ("euroc", 1000)])
def test_extract_matches_golden_reference(golden_dir, name, seed):
    g = np.load(os.path.join(golden_dir, f"ref_{name}.npz"))
    c = json.loads(str(g["cfg"]))
    img = synth.synth_image(c["width"], c["height"], seed)
    assert zlib.crc32(img.tobytes()) == int(g["img_crc"])
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kps, desc = ex(img)
    _check_against(kps, desc, g["keypoints"], g["descriptors"], f"golden {name}")
    for l in range(c["nlevels"]):
        whole = ex.pyramid_level(l, with_apron=True)
        assert tuple(whole.shape) == tuple(g["level_whole_shape"][l])
        assert zlib.crc32(whole.tobytes()) == int(g["level_crc"][l]), f"pyramid level {l} (apron included)"


@pytest.mark.parametrize("name,seed", [("tum1", 5), ("kitti", 6), ("euroc", 1003)])
def test_stages_match_oracle(name, seed):
    c = _cfg(name)
    img = synth.synth_image(c["width"], c["height"], seed)
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    orc = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kps, desc = ex(img)
    kps_o, desc_o = orc.extract(img)
    for l in range(c["nlevels"]):
        assert np.array_equal(ex.pyramid_level(l, with_apron=True), orc.level(l)), f"pyramid level {l}"
        bo = orc.level(l, blurred=True)              # cv::GaussianBlur of the level (None when the level has no keypoints)
        if bo is not None:
            assert np.array_equal(ex.blurred_level(l), bo), f"GaussianBlur of level {l}"
        co = orc.candidates(l)
        cg = ex.debug_candidates(l)
        assert len(cg) == len(co), f"level {l}: {len(cg)} vs {len(co)} FAST candidates"
        assert cg.tobytes() == co.tobytes(), f"level {l}: candidate list (order, response) differs"
    assert list(ex.debug_level_counts()) == orc.level_counts()
    _check_against(kps, desc, kps_o, desc_o, name)


@pytest.mark.parametrize("w,h,nf,sc,nl,it,mt,seed", [
    (400, 150, 500, 1.2, 5, 20, 7, 21), (330, 300, 300, 1.5, 3, 30, 10, 22), (640, 480, 50, 1.2, 8, 20, 7, 23),
    (640, 480, 3000, 1.1, 8, 12, 5, 24), (100, 90, 40, 1.2, 2, 20, 7, 25), (641, 479, 1000, 1.2, 8, 20, 7, 26),
    # scale factors above ~2.3 leave the resize kernel's word-window path (taps of 4 outputs span more than 8 bytes)
    (900, 700, 800, 2.5, 3, 20, 7, 27), (1000, 640, 500, 3.0, 3, 20, 7, 28), (640, 480, 500, 2.0, 3, 20, 7, 29)])
def test_odd_settings_match_oracle(w, h, nf, sc, nl, it, mt, seed):
    img = synth.synth_image(w, h, seed)
    kps, desc = ORBextractor(nf, sc, nl, it, mt)(img)
    kps_o, desc_o = ob.Extractor(nf, sc, nl, it, mt).extract(img)
    _check_against(kps, desc, kps_o, desc_o, f"{w}x{h}")


def test_flat_and_sparse_images():
    ex = ORBextractor(500, 1.2, 8, 20, 7)
    kps, desc = ex(np.full((480, 640), 77, np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    img = np.full((480, 640), 90, np.uint8)
    img[200:260, 300:380] = 200           # a single rectangle: 4 corners, almost every cell empty
    kps, desc = ex(img)
    kps_o, desc_o = ob.Extractor(500, 1.2, 8, 20, 7).extract(img)
    assert len(kps_o) > 0
    _check_against(kps, desc, kps_o, desc_o, "single rectangle")


def test_strided_input_and_reuse():
    big = synth.synth_image(700, 500, 31)
    view = big[10:490, 30:670]            # non-contiguous rows
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    k1, d1 = ex(view)
    k2, d2 = ob.Extractor(1000, 1.2, 8, 20, 7).extract(np.ascontiguousarray(view))
    _check_against(k1, d1, k2, d2, "strided")
    k3, d3 = ex(view)                     # same instance again: no state leaks between calls
    assert k3.tobytes() == k1.tobytes() and np.array_equal(d3, d1)


def test_batch_equals_single_frames():
    c = _cfg("euroc")
    imgs = [synth.synth_image(c["width"], c["height"], 1000 + i) for i in range(6)]
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kb, db = ex.extract_batch(imgs, max_batch=4)     # 6 frames in chunks of 4 + 2
    orc = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    for i, im in enumerate(imgs):
        ko, do = orc.extract(im)
        _check_against(kb[i], db[i], ko, do, f"batch frame {i}")


def test_4k_stress_matches_oracle():
    c = _cfg("4k")
    img = synth.synth_image(c["width"], c["height"], 7)
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kps, desc = ex(img)
    kps_o, desc_o = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"]).extract(img)
    _check_against(kps, desc, kps_o, desc_o, "4k")


def test_two_instances_from_two_threads():
    """Frame.cc:80-84: left and right extractors run concurrently in two std::threads."""
    import threading
    c = _cfg("kitti")
    left, right = synth.synth_stereo_pair(c["width"], c["height"], 2)
    exs = [ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"]) for _ in range(2)]
    out = [None, None]

    def run(i, im):
        for _ in range(3):
            out[i] = exs[i](im)
    th = [threading.Thread(target=run, args=(i, im)) for i, im in enumerate((left, right))]
    [t.start() for t in th]; [t.join() for t in th]
    orc = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    for i, im in enumerate((left, right)):
        ko, do = orc.extract(im)
        _check_against(out[i][0], out[i][1], ko, do, f"thread {i}")


# ------------------------------------------------------------------------------------------------ Hamming
def test_hamming_top2_matches_oracle_with_ties():
    train, query = synth.synth_descriptors(50_000, 700)
    query[5] = ~train[123]                       # complement row: distance 256 must never win
    i1, d1, d2 = hamming_top2(query, train)
    j1, e1, e2 = ob.hamming_top2(query, train, nthreads=8)
    assert np.array_equal(i1, j1) and np.array_equal(d1, e1) and np.array_equal(d2, e2)
    assert (np.delete(d1[:8], 5) <= 40).all() and i1[5] != 123


@pytest.mark.parametrize("nq,nt", [(1, 1), (3, 2), (513, 257), (512, 256), (1025, 4097), (7, 100_003)])
def test_hamming_ragged_sizes(nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8); t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if nt > 10:
        t[nt // 2] = q[0]; t[nt // 3] = q[0]     # exact duplicate of the best: lowest index wins, second == best
    i1, d1, d2 = hamming_top2(q, t)
    j1, e1, e2 = ob.hamming_top2(q, t)
    assert np.array_equal(i1, j1) and np.array_equal(d1, e1) and np.array_equal(d2, e2)


def test_hamming_empty_train_and_descriptor_distance():
    q = np.random.default_rng(1).integers(0, 256, (5, 32), dtype=np.uint8)
    i1, d1, d2 = hamming_top2(q, q[:0])
    assert (i1 == -1).all() and (d1 == 256).all() and (d2 == 256).all()
    for a, b in ((q[0], q[1]), (q[2], q[2]), (q[3], ~q[3])):
        want = ob.descriptor_distance(a, b)
        got = ORBmatcher.DescriptorDistance(a, b)
        assert got == want or (want == 256 and got == 256)


def test_hamming_full_size_properties():
    """BASELINE config 4 at full size (2048 x 1,000,000): size-independent properties instead of the oracle."""
    train, query = synth.synth_descriptors(1_000_000, 2048)
    i1, d1, d2 = hamming_top2(query, train)
    assert (i1 >= 0).all() and (d1 <= d2).all()
    # the reported best distance is the true distance to the reported row
    true_d = np.unpackbits(query ^ train[i1], axis=1).sum(1)
    assert np.array_equal(true_d, d1)
    # sharding property: top-2 of the union == merge of per-shard top-2 (checked on 4 shards of a 200k subset)
    sub = train[:200_000]
    a = hamming_top2(query, sub)
    parts = [hamming_top2(query, sub[s:s + 50_000]) for s in range(0, 200_000, 50_000)]
    best = np.full(2048, 256); idx = np.full(2048, -1); second = np.full(2048, 256)
    for p, (pi, pd1, pd2) in enumerate(parts):
        gi = np.where(pi >= 0, pi + p * 50_000, -1)
        take = pd1 < best                                       # strict: the lower shard wins distance ties
        second = np.minimum(np.maximum(best, pd1), np.minimum(second, pd2))
        idx = np.where(take, gi, idx)
        best = np.minimum(best, pd1)
    assert np.array_equal(idx, a[0]) and np.array_equal(best, a[1]) and np.array_equal(second, a[2])
    # against the oracle over the full million: the 8 duplicated-row queries (exact ties) and 248 more spread over the set
    sel = np.concatenate([np.arange(8), np.linspace(8, 2047, 248).astype(np.int64)])
    j1, e1, e2 = ob.hamming_top2(query[sel], train, nthreads=os.cpu_count() or 8)
    assert np.array_equal(i1[sel], j1) and np.array_equal(d1[sel], e1) and np.array_equal(d2[sel], e2)


def test_stereo_hamming_matches_oracle():
    c = _cfg("kitti")
    left, right = synth.synth_stereo_pair(c["width"], c["height"], 2)
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kl, dl = ex(left)
    kr, dr = ex(right)
    sf = ex.GetScaleFactors()
    maxD = np.float32(c["bf"]) / (np.float32(c["bf"]) / np.float32(c["fx"]))   # mbf / mb (Frame.cc:593-595)
    bi, bd = stereo_hamming(kl, dl, kr, dr, c["height"], sf, 0.0, float(maxD))
    oi, od = ob.stereo_hamming(kl, dl, kr, dr, c["height"], sf, 0.0, float(maxD))
    assert np.array_equal(bi, oi) and np.array_equal(bd, od)
    assert np.count_nonzero(bd < 75) > 200      # the synthetic pair does produce stereo matches


def test_hamming_shards_merge_on_device():
    """top-2 of a union == merge of per-shard top-2s. Emulates 3 ranks on ONE GPU: (a) every shard merged
    into the same packed buffer by the kernel's CAS epilogue, (b) separate packed parts + the merge kernel that follows
    the NCCL all-gather. Both must equal the single-shard result and the oracle."""
    import torch
    from orb_slam2_commit_b200 import api, dist as od
    L = api.lib()
    train, query = synth.synth_descriptors(90_001, 300, seed=9)
    dq = torch.from_numpy(query).cuda(); dt = torch.from_numpy(train).cuda()
    nq = len(query); st = torch.cuda.current_stream().cuda_stream
    bounds = [od.shard_range(len(train), 3, r) for r in range(3)]
    acc = torch.empty(nq, dtype=torch.int64, device="cuda")
    api._ck(L.orbx_hamming_init_device(acc.data_ptr(), nq, st))
    parts = torch.empty((3, nq), dtype=torch.int64, device="cuda")
    for r, (a, b) in enumerate(bounds):
        shard = dt[a:b].contiguous()
        api._ck(L.orbx_hamming_top2_device(dq.data_ptr(), nq, shard.data_ptr(), b - a, a, acc.data_ptr(), st))
        api._ck(L.orbx_hamming_init_device(parts[r].data_ptr(), nq, st))
        api._ck(L.orbx_hamming_top2_device(dq.data_ptr(), nq, shard.data_ptr(), b - a, a, parts[r].data_ptr(), st))
    out = torch.empty((2, 3, nq), dtype=torch.int32, device="cuda")
    api._ck(L.orbx_hamming_merge_device(acc.data_ptr(), 1, nq, out[0, 0].data_ptr(), out[0, 1].data_ptr(), out[0, 2].data_ptr(), st))
    api._ck(L.orbx_hamming_merge_device(parts.data_ptr(), 3, nq, out[1, 0].data_ptr(), out[1, 1].data_ptr(), out[1, 2].data_ptr(), st))
    torch.cuda.synchronize()
    want = ob.hamming_top2(query, train, nthreads=8)
    for v in range(2):
        got = out[v].cpu().numpy()
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), v
    # the packed words themselves decode with the host helper used by the gloo tests
    i, d1, d2 = od.unpack_top2(acc.cpu().numpy().view(np.uint64))
    assert np.array_equal(i, want[0]) and np.array_equal(d1, want[1]) and np.array_equal(d2, want[2])


def test_device_resident_api_matches_host_api():
    """orbx_extract_device (frames and results stay in HBM, caller's stream) == orbx_extract_batch."""
    import torch
    from orb_slam2_commit_b200 import api
    c = _cfg("tum1")
    imgs = np.stack([synth.synth_image(c["width"], c["height"], 40 + i) for i in range(5)])
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    cap = ex.reserve(c["width"], c["height"], 5)
    d_img = torch.from_numpy(imgs).cuda()
    d_kps = torch.zeros((5, cap, 28), dtype=torch.uint8, device="cuda")
    d_desc = torch.zeros((5, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(5, dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ex.extract_device(d_img.data_ptr(), 5, c["width"], c["height"], c["width"], c["width"] * c["height"],
                          d_kps.data_ptr(), cap, d_n.data_ptr(), d_desc.data_ptr(), s.cuda_stream)
    s.synchronize()
    kb, db = ex.extract_batch(list(imgs))
    n = d_n.cpu().numpy()
    for i in range(5):
        k = d_kps[i, :n[i]].cpu().numpy().view(api.KP_DTYPE).reshape(-1)
        assert k.tobytes() == kb[i].tobytes()
        assert np.array_equal(d_desc[i, :n[i]].cpu().numpy(), db[i])


@pytest.mark.parametrize("name,seed", [("kitti", 2), ("euroc", 1001)])
def test_full_stereo_match_matches_oracle(name, seed):
    """Frame::ComputeStereoMatches end to end (Frame.cc:547-788): Hamming + SAD + parabola + median cut on the
    HBM-resident pyramids of a left and a right extractor. mvuRight and mvDepth must be bit-identical floats."""
    c = _cfg(name)
    left, right = synth.synth_stereo_pair(c["width"], c["height"], seed)
    args = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    exL, exR = ORBextractor(*args), ORBextractor(*args)
    kl, dl = exL(left); kr, dr = exR(right)
    ur, dp = stereo_match(exL, exR, kl, dl, kr, dr, c["bf"], c["fx"])
    oL, oR = ob.Extractor(*args), ob.Extractor(*args)
    kl_o, dl_o = oL.extract(left); kr_o, dr_o = oR.extract(right)
    assert kl.tobytes() == kl_o.tobytes() and kr.tobytes() == kr_o.tobytes()
    ur_o, dp_o = ob.stereo_match(oL, oR, kl_o, dl_o, kr_o, dr_o, c["bf"], c["fx"])
    assert np.array_equal(ur.view(np.uint32), ur_o.view(np.uint32))
    assert np.array_equal(dp.view(np.uint32), dp_o.view(np.uint32))
    assert np.count_nonzero(ur >= 0) > 100


def test_capacity_error_and_geometry_change_on_one_instance():
    import ctypes as C
    from orb_slam2_commit_b200 import api
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    a = synth.synth_image(640, 480, 71); b = synth.synth_image(752, 480, 72)
    ka, da = ex(a); kb, db = ex(b); ka2, da2 = ex(a)          # the working set is re-reserved when the size changes
    orc = ob.Extractor(1000, 1.2, 8, 20, 7)
    _check_against(ka, da, *orc.extract(a), "640x480")
    _check_against(kb, db, *orc.extract(b), "752x480")
    assert ka2.tobytes() == ka.tobytes() and np.array_equal(da2, da)
    # a caller buffer that is too small: status 3 (ORBX_ERR_CAPACITY), *nkp still reports the required size
    L = api.lib(); n = C.c_int32(0)
    kps = np.zeros(100, api.KP_DTYPE); desc = np.zeros((100, 32), np.uint8)
    rc = L.orbx_extract(ex._h, a.ctypes.data_as(api.u8p), 640, 480, 640, kps.ctypes.data, 100, C.byref(n), desc.ctypes.data_as(api.u8p))
    assert rc == 3 and n.value == len(ka)
    assert kps.tobytes() == ka[:100].tobytes() and np.array_equal(desc, da[:100])


def test_batched_device_stereo_matches_oracle():
    """orbx_stereo_match_device: 3 stereo pairs in one call, everything resident in HBM (frames -> two
    orbx_extract_device calls -> matcher). Every pair must equal the oracle's Frame::ComputeStereoMatches."""
    import torch
    from orb_slam2_commit_b200 import api, stereo_match_device
    c = _cfg("euroc")
    args = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    pairs = [synth.synth_stereo_pair(c["width"], c["height"], 1010 + i) for i in range(3)]
    W, H, P = c["width"], c["height"], len(pairs)
    exL, exR = ORBextractor(*args), ORBextractor(*args)
    cap = exL.reserve(W, H, P); exR.reserve(W, H, P)
    dev = {}
    for side, ex, imgs in (("l", exL, [p[0] for p in pairs]), ("r", exR, [p[1] for p in pairs])):
        d_img = torch.from_numpy(np.stack(imgs)).cuda()
        kps = torch.zeros((P, cap, 28), dtype=torch.uint8, device="cuda"); desc = torch.zeros((P, cap, 32), dtype=torch.uint8, device="cuda")
        n = torch.zeros(P, dtype=torch.int32, device="cuda")
        ex.extract_device(d_img.data_ptr(), P, W, H, W, W * H, kps.data_ptr(), cap, n.data_ptr(), desc.data_ptr(), 0)
        ex.synchronize()
        dev[side] = (kps, desc, n)
    ur = torch.zeros((P, cap), dtype=torch.float32, device="cuda"); dp = torch.zeros((P, cap), dtype=torch.float32, device="cuda")
    stereo_match_device(exL, exR, P, dev["l"][0].data_ptr(), dev["l"][1].data_ptr(), dev["l"][2].data_ptr(),
                        dev["r"][0].data_ptr(), dev["r"][1].data_ptr(), dev["r"][2].data_ptr(), cap, c["bf"], c["fx"],
                        ur.data_ptr(), dp.data_ptr(), 0)
    exL.synchronize()
    nl = dev["l"][2].cpu().numpy()
    for i, (left, right) in enumerate(pairs):
        oL, oR = ob.Extractor(*args), ob.Extractor(*args)
        kl, dl = oL.extract(left); kr, dr = oR.extract(right)
        assert nl[i] == len(kl)
        ur_o, dp_o = ob.stereo_match(oL, oR, kl, dl, kr, dr, c["bf"], c["fx"])
        assert np.array_equal(ur[i, :nl[i]].cpu().numpy().view(np.uint32), ur_o.view(np.uint32)), i
        assert np.array_equal(dp[i, :nl[i]].cpu().numpy().view(np.uint32), dp_o.view(np.uint32)), i
        assert np.count_nonzero(ur_o >= 0) > 50


def test_window_top2_matches_oracle():
    """SearchByProjection's candidate loop over GetFeaturesInArea on the 64x48 Frame grid (Frame.cc:388-444,
    ORBmatcher.cc:46-142): queries = keypoints of a second frame 'projected' with noise, incl. occupied keypoints,
    stereo gating, windows leaving the image and duplicate descriptors (grid-order tie-break)."""
    from orb_slam2_commit_b200 import window_top2
    c = _cfg("euroc")
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    k1, d1 = ex(synth.synth_image(c["width"], c["height"], 1200))
    k2, d2 = ex(synth.synth_image(c["width"], c["height"], 1200)[::1, ::1])     # same frame: true matches exist
    rng = np.random.default_rng(3)
    sf = ex.GetScaleFactors()
    n = len(k2)
    q = np.zeros(n + 40, ob.WQ_DTYPE); qd = np.zeros((n + 40, 32), np.uint8)
    q["x"][:n] = k2["x"] + rng.normal(0, 3, n).astype(np.float32); q["y"][:n] = k2["y"] + rng.normal(0, 3, n).astype(np.float32)
    lvl = k2["octave"]
    q["r"][:n] = (np.where(rng.random(n) < 0.5, 2.5, 4.0) * 3).astype(np.float32) * sf[lvl]
    q["min_level"][:n] = lvl - 1; q["max_level"][:n] = lvl
    q["xr"][:n] = q["x"][:n] - rng.uniform(0, 40, n).astype(np.float32)
    qd[:n] = d2
    flip = rng.integers(0, 32, n); qd[np.arange(n), flip] ^= rng.integers(0, 256, n).astype(np.uint8)
    # 40 extra queries: windows partly / fully outside the image, no level check, huge radius
    q["x"][n:] = rng.uniform(-200, c["width"] + 200, 40); q["y"][n:] = rng.uniform(-200, c["height"] + 200, 40)
    q["r"][n:] = rng.uniform(5, 300, 40); q["min_level"][n:] = -1; q["max_level"][n:] = -1; q["xr"][n:] = -1
    qd[n:] = d1[rng.integers(0, len(d1), 40)]
    d1 = d1.copy(); d1[11] = d1[10]; d1[500] = d1[10]                    # exact ties: the grid-order first one must win
    occ = (rng.random(len(k1)) < 0.1).astype(np.uint8)
    ur = np.where(rng.random(len(k1)) < 0.5, k1["x"] - rng.uniform(0, 40, len(k1)), -1).astype(np.float32)
    minX, minY = 0.0, 0.0
    invW = np.float32(64) / np.float32(c["width"] - minX); invH = np.float32(48) / np.float32(c["height"] - minY)
    for occupied, uright in ((None, None), (occ, ur)):
        got = window_top2(k1, d1, occupied, uright, minX, minY, float(invW), float(invH), q, qd)
        want = ob.window_top2(k1, d1, occupied, uright, minX, minY, float(invW), float(invH), q, qd)
        for g, w_, name in zip(got, want, ("bestIdx", "bestDist", "bestLevel", "bestDist2", "bestLevel2")):
            assert np.array_equal(g, w_), name
        assert np.count_nonzero(want[0] >= 0) > n // 2


@pytest.mark.parametrize("channels,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_colour_input_fused_gray_conversion(channels, rgb):
    """Tracking::GrabImage* converts colour frames with cv::cvtColor before extraction (Tracking.cc:174-199); the fused
    level-0 kernel must give the result of extracting the oracle's gray image (pinned to cv2 4.13 by the golden tests)."""
    c = _cfg("tum1")
    rng = np.random.default_rng(channels * 2 + rgb)
    base = synth.synth_image(c["width"], c["height"], 90 + channels)
    col = np.stack([np.clip(base.astype(np.int32) + rng.integers(-40, 41, base.shape), 0, 255).astype(np.uint8) for _ in range(channels)], axis=2)
    gray = ob.cvt_gray(col, rgb)
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    kps, desc = ex.extract_color(col, rgb)
    assert np.array_equal(ex.pyramid_level(0), gray)
    ko, do = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"]).extract(gray)
    _check_against(kps, desc, ko, do, f"colour {channels} rgb={rgb}")
    k2, d2 = ex(gray)                                  # back to gray on the same instance
    _check_against(k2, d2, ko, do, "gray after colour")


def test_random_geometries_match_oracle():
    """Seeded sweep over image sizes / pyramid settings / thresholds (odd widths, single level, tiny levels, many levels,
    very wide frames with several quadtree roots); settings the reference itself cannot run are skipped."""
    from orb_slam2_commit_b200 import OrbxError
    rng = np.random.default_rng(2026)
    ran = 0
    for trial in range(40):
        w = int(rng.integers(90, 1000)); h = int(rng.integers(70, 700))
        nl = int(rng.integers(1, 9)); sc = float(rng.choice([1.1, 1.2, 1.25, 1.5, 2.0]))
        nf = int(rng.choice([30, 200, 1000, 3000])); ini = int(rng.integers(8, 40)); mn = int(rng.integers(1, ini + 1))
        img = synth.synth_image(w, h, 500 + trial)
        ex = ORBextractor(nf, sc, nl, ini, mn)
        try:
            kps, desc = ex(img)
        except OrbxError as e:
            assert e.code == 2, e          # ORBX_ERR_UNSUPPORTED: level < 62 px or portrait level
            continue
        ko, do = ob.Extractor(nf, sc, nl, ini, mn).extract(img)
        _check_against(kps, desc, ko, do, f"trial {trial}: {w}x{h} nf={nf} sc={sc} nl={nl} th={ini}/{mn}")
        ran += 1
    assert ran >= 15


def test_clustered_keypoints_deep_quadtree():
    """All texture packed into small patches: the quadtree has to go far deeper than a uniform spread of the quota
    would, which overflows the count pyramid of the fast path and forces the sweep-path fallback (and, with several
    patches, mixes shallow and deep leaves)."""
    rng = np.random.default_rng(5)
    tex = synth._smooth(rng.integers(0, 256, (480, 640)).astype(np.float64), 1.6)
    tex = ((tex - tex.min()) * 255 / (tex.max() - tex.min())).astype(np.uint8)
    # one patch: a node with a single non-empty child stops the reference's loop (size == prevSize,
    # ORBextractor.cc:650-653), so very few keypoints survive; three patches: deep, unbalanced trees.
    for patches, floor in (([(200, 150, 110)], 8), ([(40, 60, 70), (500, 330, 90), (320, 40, 50)], 500)):
        img = np.full((480, 640), 120, np.uint8)
        for (x, y, s) in patches:
            img[y:y + s, x:x + s] = tex[y:y + s, x:x + s]
        for nf in (1000, 3000):
            kps, desc = ORBextractor(nf, 1.2, 8, 20, 7)(img)
            ko, do = ob.Extractor(nf, 1.2, 8, 20, 7).extract(img)
            assert len(ko) > floor
            _check_against(kps, desc, ko, do, f"clustered {len(patches)} patches nf={nf}")


def test_rectified_extract_matches_oracle():
    """stereo_euroc.cc:136-137: cv::remap(im, imRect, M1, M2, INTER_LINEAR) fused into level 0. Level 0 must equal the
    oracle's remap (pinned to cv2 4.13) bit for bit and everything downstream must equal the oracle extractor run on it."""
    c = _cfg("euroc")
    m1, m2 = synth.rectify_maps(c["width"], c["height"])
    ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    ex.set_rectify_maps(m1, m2)
    imgs = [synth.synth_image(c["width"], c["height"], 1200 + i) for i in range(3)]
    kps, descs = ex.extract_rectified(imgs)
    orc = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    for i, im in enumerate(imgs):
        rect = ob.remap(im, m1, m2)
        assert np.array_equal(ex.pyramid_level(0, frame=i), rect), f"rectified level 0 of frame {i}"
        ko, do = orc.extract(rect)
        _check_against(kps[i], descs[i], ko, do, f"rectified frame {i}")
    # the plain entry point still takes rectified frames as they are
    k2, d2 = ex(ob.remap(imgs[0], m1, m2))
    _check_against(k2, d2, kps[0], descs[0], "plain call after a rectified call")


def test_rectified_wild_maps_and_source_size():
    """Maps that leave the source (BORDER_CONSTANT 0), hit exact 1/64 ties and address a source of another size."""
    rng = np.random.default_rng(9)
    sw, sh, mw, mh = 300, 200, 320, 240
    src = synth.synth_image(sw, sh, 31)
    yy, xx = np.mgrid[0:mh, 0:mw].astype(np.float32)
    m1 = (xx * (sw + 30) / mw - 15 + rng.random((mh, mw), dtype=np.float32) * 2).astype(np.float32)
    m2 = (yy * (sh + 30) / mh - 15 + rng.random((mh, mw), dtype=np.float32) * 2).astype(np.float32)
    m1[::7] = np.round(m1[::7] * 64) / 64; m2[::5] = np.round(m2[::5] * 64) / 64
    m1[3, :6] = [-1, -0.5, sw - 1, sw - 0.5, sw, 1e6]; m2[3, :6] = [-1, sh - 1, sh - 0.5, sh, -0.5, -1e6]
    ex = ORBextractor(500, 1.2, 4, 20, 7)
    ex.set_rectify_maps(m1, m2, src_size=(sw, sh))
    kps, descs = ex.extract_rectified([src])
    rect = ob.remap(src, m1, m2)
    assert np.array_equal(ex.pyramid_level(0), rect)
    ko, do = ob.Extractor(500, 1.2, 4, 20, 7).extract(rect)
    assert len(ko) > 100
    _check_against(kps[0], descs[0], ko, do, "wild maps")


def test_undistort_keypoints_bit_exact(golden_dir):
    """Frame::UndistortKeyPoints / ComputeImageBounds (Frame.cc:471-538) against cv2 4.13's undistortPoints (golden) and
    against the oracle on real keypoints of every octave."""
    from orb_slam2_commit_b200 import KP_DTYPE, image_bounds, undistort_keypoints
    g = np.load(os.path.join(golden_dir, "prims2_cv2.npz"))
    pts = g["undist_src"]
    kin = np.zeros(len(pts), KP_DTYPE); kin["x"] = pts[:, 0]; kin["y"] = pts[:, 1]; kin["octave"] = 3; kin["angle"] = 12.5
    for nd, key in ((5, "undist_dst5"), (4, "undist_dst4")):
        out = undistort_keypoints(kin, g["undist_K4"], g["undist_D"][:nd])
        assert np.array_equal(out["x"].view(np.uint32), g[key][:, 0].view(np.uint32))
        assert np.array_equal(out["y"].view(np.uint32), g[key][:, 1].view(np.uint32))
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(out[f], kin[f])
    kps, _ = ORBextractor(1000, 1.2, 8, 20, 7)(synth.synth_image(640, 480, 3))
    out = undistort_keypoints(kps, synth.TUM1_K4, synth.TUM1_DIST)
    ref = ob.undistort_points(np.stack([kps["x"], kps["y"]], 1), synth.TUM1_K4, synth.TUM1_DIST)
    assert np.array_equal(out["x"].view(np.uint32), ref[:, 0].view(np.uint32))
    assert np.array_equal(out["y"].view(np.uint32), ref[:, 1].view(np.uint32))
    assert np.abs(out["x"] - kps["x"]).max() > 0.5                      # the model really moves points
    # k1 == 0: mvKeysUn = mvKeys (Frame.cc:474-478)
    same = undistort_keypoints(kps, synth.TUM1_K4, np.zeros(5, np.float32))
    assert same.tobytes() == kps.tobytes()
    # ComputeImageBounds = min / max over the four undistorted corners (first four golden points)
    b = image_bounds(640, 480, g["undist_K4"], g["undist_D"])
    c = g["undist_dst5"][:4]
    assert np.array_equal(b, np.array([min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])], np.float32))
    assert np.array_equal(image_bounds(640, 480, synth.TUM1_K4, np.zeros(4, np.float32)), np.array([0, 640, 0, 480], np.float32))


def _bow_equal(got, ref, what):
    for k in ("word", "node", "bow_id", "fv_node", "fv_off", "fv_feat"):
        assert np.array_equal(got[k], ref[k]), f"{what}: {k} differs"
    assert np.array_equal(got["bow_val"].view(np.uint64), ref["bow_val"].view(np.uint64)), f"{what}: BowVector values differ (bitwise)"


@pytest.mark.parametrize("k,L,levelsup,scoring,weighting", [(10, 4, 2, 0, 0), (6, 5, 4, 0, 0), (7, 3, 4, 1, 1), (5, 4, 1, 5, 0),
                                                             (9, 3, 2, 0, 2), (20, 2, 1, 0, 3)])
def test_bow_transform_matches_oracle(k, L, levelsup, scoring, weighting):
    """Frame::ComputeBoW = ORBVocabulary::transform(desc, mBowVec, mFeatVec, levelsup): words, nodes, BowVector
    (bit-identical doubles) and FeatureVector against the DBoW2 restatement, batched with ragged counts, incl. stopped
    words, early leaves, L - levelsup <= 0 (nodes collapse to the root) and every weighting / normalisation branch."""
    from orb_slam2_commit_b200 import ORBVocabulary
    voc = synth.synth_vocabulary(k, L, 100 + k)
    V = ORBVocabulary(k, L, *voc, scoring=scoring, weighting=weighting)
    Vo = ob.Vocabulary(k, L, *voc, scoring=scoring, weighting=weighting)
    assert V.nwords == Vo.nwords
    rng = np.random.default_rng(k)
    descs = [synth.synth_features_near_words(voc, n, 7 + n) for n in (1000, 37, 1, 2048)]
    descs.append(rng.integers(0, 256, (300, 32), dtype=np.uint8))          # far from every word
    descs.append(np.repeat(descs[0][:5], 40, axis=0))                      # many features per word (accumulation order)
    got = V.transform_batch(descs, levelsup)
    for i, d in enumerate(descs):
        _bow_equal(got[i], Vo.transform(d, levelsup), f"k={k} L={L} frame {i}")
    assert any(len(g["bow_id"]) < (g["word"] >= 0).sum() for g in got)     # some words repeat
    _bow_equal(V.transform(descs[1], levelsup), Vo.transform(descs[1], levelsup), "single-frame form")


def test_bow_on_extractor_output_and_l1_score():
    """BoW of real extractor descriptors (batched), then L1Scoring::score between all pairs of frames, bit-identical."""
    from orb_slam2_commit_b200 import ORBVocabulary
    voc = synth.synth_vocabulary(10, 4, 5)
    V = ORBVocabulary(10, 4, *voc); Vo = ob.Vocabulary(10, 4, *voc)
    ex = ORBextractor(1000, 1.2, 8, 20, 7)
    imgs = [synth.synth_image(640, 480, 60 + i) for i in range(3)]
    imgs.append(np.roll(imgs[0], 3, axis=1))                               # near-duplicate of frame 0 -> many common words
    _, descs = ex.extract_batch(imgs)
    got = V.transform_batch(descs, 2)
    refs = [Vo.transform(d, 2) for d in descs]
    for i in range(len(descs)):
        _bow_equal(got[i], refs[i], f"frame {i}")
    a, b = np.meshgrid(np.arange(4), np.arange(4))
    s = V.score(a.ravel(), b.ravel())
    so = np.array([ob.bow_score_l1(refs[i]["bow_id"], refs[i]["bow_val"], refs[j]["bow_id"], refs[j]["bow_val"])
                   for i, j in zip(a.ravel(), b.ravel())])
    assert np.array_equal(s.view(np.uint64), so.view(np.uint64))
    assert s.reshape(4, 4)[0, 0] > 0.99 and s.max() <= 1.0 + 1e-12


@pytest.mark.parametrize("check_ori,nnratio", [(True, 0.7), (False, 0.9), (True, 0.95)])
def test_search_by_bow_matches_oracle(check_ori, nnratio):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...): per-node sequential best / second-best with the already-matched
    skip, TH_LOW, NN ratio, rotation histogram + ComputeThreeMaxima; matches and count identical to the restatement."""
    from orb_slam2_commit_b200 import ORBVocabulary
    voc = synth.synth_vocabulary(10, 4, 11)
    V = ORBVocabulary(10, 4, *voc); Vo = ob.Vocabulary(10, 4, *voc)
    ex = ORBextractor(1500, 1.2, 8, 20, 7)
    base = synth.synth_image(640, 480, 90)
    rng = np.random.default_rng(4)
    moved = np.clip(np.roll(base, (2, 5), axis=(0, 1)).astype(np.int16) + rng.integers(-6, 7, base.shape), 0, 255).astype(np.uint8)
    (kk, kf), (dk, df) = ex.extract_batch([base, moved])
    valid = (rng.random(len(kk)) < 0.8).astype(np.uint8)
    for kf_valid in (valid, None):
        n, m = V.search_by_bow(kk, dk, kf_valid, kf, df, levelsup=2, nnratio=nnratio, check_orientation=check_ori)
        to, tf = Vo.transform(dk, 2), Vo.transform(df, 2)
        no, mo = ob.search_by_bow(to, tf, dk, kk["angle"], np.ones(len(kk), np.uint8) if kf_valid is None else kf_valid, df, kf["angle"],
                                  nnratio, check_ori)
        assert n == no and np.array_equal(m, mo)
        assert n > 50 and n == np.count_nonzero(m >= 0)
    # duplicated frame descriptors inside one node: first-in-list wins, duplicates give bestDist2 == bestDist1 (ratio fails)
    df2 = df.copy(); df2[1::2] = df2[0::2][:len(df2[1::2])]
    n, m = V.search_by_bow(kk, dk, None, kf, df2, levelsup=2, nnratio=nnratio, check_orientation=check_ori)
    to, tf = Vo.transform(dk, 2), Vo.transform(df2, 2)
    no, mo = ob.search_by_bow(to, tf, dk, kk["angle"], np.ones(len(kk), np.uint8), df2, kf["angle"], nnratio, check_ori)
    assert n == no and np.array_equal(m, mo)


@pytest.mark.parametrize("seed,stereo", [(1, True), (2, False), (3, True)])
def test_search_by_projection_frame_matches_oracle(seed, stereo):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1489-1646): the assignment of
    last-frame map points to current keypoints (incl. the sequential already-taken rule under heavy contention, the
    overwrite by unobserved points, forward / backward / neutral level ranges, the stereo gate and the rotation check)
    and the match count are identical to the restatement."""
    from orb_slam2_commit_b200 import search_by_projection_frame
    s = synth.synth_tracking_scene(seed, stereo=stereo)
    for mode in (0, 1, 2):
        for th, ori in ((7.0, True), (15.0, False)):
            n, m = search_by_projection_frame(**s, th=th, mode=mode, check_orientation=ori)
            no, mo = ob.search_by_projection_frame(**s, th=th, mode=mode, check_orientation=ori)
            assert n == no, f"seed {seed} mode {mode} th {th}: nmatches {n} vs {no}"
            assert np.array_equal(m, mo), f"seed {seed} mode {mode} th {th}: {np.count_nonzero(m != mo)} assignments differ"
            assert n > 100
    # all map points observed (every assignment blocks later ones) and none observed (free overwriting)
    for bit in (2, 0):
        s2 = dict(s); s2["last_flags"] = (s["last_flags"] & 1) | bit
        n, m = search_by_projection_frame(**s2, th=15.0, mode=0, check_orientation=True)
        no, mo = ob.search_by_projection_frame(**s2, th=15.0, mode=0, check_orientation=True)
        assert n == no and np.array_equal(m, mo)
    # degenerate inputs
    e = dict(s); e["last_flags"] = np.zeros_like(s["last_flags"])
    n, m = search_by_projection_frame(**e, th=7.0, mode=0)
    assert n == 0 and (m == -1).all()


def test_pyramid_accessor_after_chunked_host_batch():
    """The host batch path runs several chunks in flight in slices of the working set; the pyramid accessor maps a frame
    index of the call to its slice and refuses frames whose slice a later chunk has reused."""
    from orb_slam2_commit_b200 import OrbxError
    imgs = [synth.synth_image(320, 240, 400 + i) for i in range(20)]
    ex = ORBextractor(300, 1.2, 4, 20, 7)
    kps, descs = ex.extract_batch(imgs, max_batch=20)          # working set 20 -> chunks of 8 in 2 slots: 8 + 8 + 4 frames
    orc = ob.Extractor(300, 1.2, 4, 20, 7)
    for i in (8, 15, 16, 19):
        ko, do = orc.extract(imgs[i])
        _check_against(kps[i], descs[i], ko, do, f"frame {i}")
        assert np.array_equal(ex.pyramid_level(0, frame=i), imgs[i])
        assert np.array_equal(ex.pyramid_level(2, frame=i, with_apron=True), orc.level(2))
    with pytest.raises(OrbxError):
        ex.pyramid_level(0, frame=3)                           # slot reused by frames 16..19


@pytest.mark.parametrize("check_ori,nnratio", [(True, 0.75), (False, 0.9)])
def test_search_by_bow_keyframe_pair_matches_oracle(check_ori, nnratio):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (ORBmatcher.cc:589-736): map-point flags on both sides,
    vbMatched2, strict `< TH_LOW`, output indexed by the first keyframe, rotation check on idx1."""
    from orb_slam2_commit_b200 import ORBVocabulary
    voc = synth.synth_vocabulary(10, 4, 21)
    V = ORBVocabulary(10, 4, *voc); Vo = ob.Vocabulary(10, 4, *voc)
    ex = ORBextractor(1500, 1.2, 8, 20, 7)
    base = synth.synth_image(640, 480, 95)
    rng = np.random.default_rng(6)
    moved = np.clip(np.roll(base, (1, 4), axis=(0, 1)).astype(np.int16) + rng.integers(-5, 6, base.shape), 0, 255).astype(np.uint8)
    (k1, k2), (d1, d2) = ex.extract_batch([base, moved])
    d2 = d2.copy(); d2[1::3] = d2[0::3][:len(d2[1::3])]                     # duplicates: ties and ratio failures
    v1 = (rng.random(len(k1)) < 0.8).astype(np.uint8); v2 = (rng.random(len(k2)) < 0.85).astype(np.uint8)
    t1, t2 = Vo.transform(d1, 2), Vo.transform(d2, 2)
    for a, b in ((v1, v2), (None, None)):
        n, m = V.search_by_bow_kf(k1, d1, a, k2, d2, b, levelsup=2, nnratio=nnratio, check_orientation=check_ori)
        ones1, ones2 = np.ones(len(k1), np.uint8), np.ones(len(k2), np.uint8)
        no, mo = ob.search_by_bow_kf(t1, t2, d1, k1["angle"], ones1 if a is None else a, d2, k2["angle"], ones2 if b is None else b,
                                     nnratio, check_ori)
        assert n == no and np.array_equal(m, mo)
        assert n > 50 and n == np.count_nonzero(m >= 0)


def test_stereo_extract_batch_host_matches_oracle():
    """orbx_stereo_extract_batch: left + right extraction and ComputeStereoMatches for a batch of pairs from host buffers,
    pipelined in chunks over several streams; every pair must equal the oracle (keypoints, descriptors, mvuRight, mvDepth)."""
    from orb_slam2_commit_b200 import KP_DTYPE, stereo_extract_host
    W, H, n = 480, 200, 11                                   # working set 11 -> chunks of 4 pairs in 2 slots
    args = (600, 1.2, 5, 20, 7)
    pairs = [synth.synth_stereo_pair(W, H, 300 + i, max_disp=40) for i in range(n)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    exL, exR = ORBextractor(*args), ORBextractor(*args)
    cap = exL.reserve(W, H, n); assert exR.reserve(W, H, n) == cap
    out = dict(kl=np.zeros((n, cap), KP_DTYPE), kr=np.zeros((n, cap), KP_DTYPE), dl=np.zeros((n, cap, 32), np.uint8),
               dr=np.zeros((n, cap, 32), np.uint8), nl=np.zeros(n, np.int32), nr=np.zeros(n, np.int32),
               u_right=np.zeros((n, cap), np.float32), depth=np.zeros((n, cap), np.float32))
    bf, fx = 386.1448 * 0.4, 718.856 * 0.4
    stereo_extract_host(exL, exR, L, R, bf, fx, out)
    oL, oR = ob.Extractor(*args), ob.Extractor(*args)
    matched = 0
    for i in range(n):
        kl_o, dl_o = oL.extract(L[i]); kr_o, dr_o = oR.extract(R[i])
        nl, nr = int(out["nl"][i]), int(out["nr"][i])
        assert out["kl"][i, :nl].tobytes() == kl_o.tobytes() and out["kr"][i, :nr].tobytes() == kr_o.tobytes(), f"pair {i}"
        assert np.array_equal(out["dl"][i, :nl], dl_o) and np.array_equal(out["dr"][i, :nr], dr_o)
        ur_o, dp_o = ob.stereo_match(oL, oR, kl_o, dl_o, kr_o, dr_o, bf, fx)
        assert np.array_equal(out["u_right"][i, :nl].view(np.uint32), ur_o.view(np.uint32)), f"pair {i}: mvuRight"
        assert np.array_equal(out["depth"][i, :nl].view(np.uint32), dp_o.view(np.uint32)), f"pair {i}: mvDepth"
        matched += int(np.count_nonzero(ur_o >= 0))
    assert matched > 20 * n


def test_rows_around_the_extractor_handle_empty_and_degenerate_inputs():
    """Empty frames inside a BoW batch, frames with one feature, empty keypoint sets for the matchers and the undistortion."""
    from orb_slam2_commit_b200 import KP_DTYPE, ORBVocabulary, search_by_projection_frame, undistort_keypoints
    voc = synth.synth_vocabulary(5, 3, 2)
    V = ORBVocabulary(5, 3, *voc); Vo = ob.Vocabulary(5, 3, *voc)
    f = synth.synth_features_near_words(voc, 40, 1)
    got = V.transform_batch([f[:0], f, f[:1], f[:0]], 1)
    for g, d in zip(got, (f[:0], f, f[:1], f[:0])):
        ref = Vo.transform(d, 1)
        for k in ("word", "node", "bow_id", "fv_node", "fv_off", "fv_feat"):
            assert np.array_equal(g[k], ref[k]), k
        assert np.array_equal(g["bow_val"].view(np.uint64), ref["bow_val"].view(np.uint64))
    assert len(got[0]["bow_id"]) == 0 and list(got[0]["fv_off"]) == [0]
    assert np.array_equal(V.score([0, 1, 1], [1, 1, 2]), np.array([ob.bow_score_l1(got[a]["bow_id"], got[a]["bow_val"], got[b]["bow_id"], got[b]["bow_val"])
                                                                     for a, b in ((0, 1), (1, 1), (1, 2))]))
    kp = np.zeros(40, KP_DTYPE); kp["angle"] = np.linspace(0, 359, 40, dtype=np.float32)
    n, m = V.search_by_bow(kp[:0], f[:0], None, kp, f, levelsup=1)
    assert n == 0 and (m == -1).all()
    n, m = V.search_by_bow(kp, f, None, kp[:0], f[:0], levelsup=1)
    assert n == 0 and len(m) == 0
    n, m = V.search_by_bow(kp, f, None, kp, f, levelsup=1, nnratio=0.99, check_orientation=True)     # a frame against itself
    no, mo = ob.search_by_bow(Vo.transform(f, 1), Vo.transform(f, 1), f, kp["angle"], np.ones(40, np.uint8), f, kp["angle"], 0.99, True)
    assert n == no and np.array_equal(m, mo)
    assert len(undistort_keypoints(kp[:0], synth.TUM1_K4, synth.TUM1_DIST)) == 0
    s = synth.synth_tracking_scene(9, n_last=50, n_extra=10)
    e = dict(s); e["cur_kps"] = s["cur_kps"][:0]; e["cur_desc"] = s["cur_desc"][:0]; e["cur_u_right"] = None; e["cur_occupied"] = None
    n, m = search_by_projection_frame(**e, th=7.0, mode=0)
    assert n == 0 and len(m) == 0
    e = dict(s); e["last_kps"] = s["last_kps"][:0]; e["last_xyz"] = s["last_xyz"][:0]; e["last_desc"] = s["last_desc"][:0]; e["last_flags"] = s["last_flags"][:0]
    n, m = search_by_projection_frame(**e, th=7.0, mode=0)
    assert n == 0 and (m == -1).all()


def test_full_stereo_match_large_scale_factor_plain_path():
    """Scale factors above 10 at the top octave leave the stereo matcher's 8-row band buckets (a right keypoint's row
    interval could touch more bands than the buckets reserve): the un-bucketed scan must give the same result."""
    W, H = 1600, 1000
    args = (1500, 1.5, 7, 20, 7)                               # 1.5^6 = 11.4
    left, right = synth.synth_stereo_pair(W, H, 77, max_disp=50)
    exL, exR = ORBextractor(*args), ORBextractor(*args)
    kl, dl = exL(left); kr, dr = exR(right)
    ur, dp = stereo_match(exL, exR, kl, dl, kr, dr, 400.0, 700.0)
    oL, oR = ob.Extractor(*args), ob.Extractor(*args)
    kl_o, dl_o = oL.extract(left); kr_o, dr_o = oR.extract(right)
    assert kl.tobytes() == kl_o.tobytes() and kr.tobytes() == kr_o.tobytes()
    ur_o, dp_o = ob.stereo_match(oL, oR, kl_o, dl_o, kr_o, dr_o, 400.0, 700.0)
    assert np.array_equal(ur.view(np.uint32), ur_o.view(np.uint32)) and np.array_equal(dp.view(np.uint32), dp_o.view(np.uint32))
    assert np.count_nonzero(ur >= 0) > 50 and kl["octave"].max() == 6


def test_batches_in_flight_give_the_results_of_the_synchronous_call():
    """orbx_extract_batch_begin / _end: two batches in flight (uploads of one overlapping the kernels of the other, shared
    working-set slices and staging buffers) return exactly what orbx_extract_batch returns for each batch; a synchronous
    call or a batch of another shape completes what is in flight first."""
    from orb_slam2_commit_b200 import KP_DTYPE, OrbxError
    W, H, B = 320, 240, 96
    imgs = [np.stack([synth.synth_image(W, H, 700 + 10 * s + (i % 7)) for i in range(B)]) for s in range(3)]
    ex = ORBextractor(300, 1.2, 4, 20, 7)
    cap = ex.reserve(W, H, B)
    ref = []
    for s in range(3):
        k = np.zeros((B, cap), KP_DTYPE); d = np.zeros((B, cap, 32), np.uint8); n = np.zeros(B, np.int32)
        ex.extract_host(imgs[s], k, d, n); ref.append((k, d, n))
    out = [(np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32)) for _ in range(3)]
    ex.extract_host_begin(imgs[0], *out[0])
    ex.extract_host_begin(imgs[1], *out[1])
    ex.extract_host_end()                                        # batch 0
    ex.extract_host_begin(imgs[2], *out[2])                      # batch 1 still in flight
    ex.extract_host_end(); ex.extract_host_end()
    for s in range(3):
        assert np.array_equal(out[s][2], ref[s][2]) and int(ref[s][2].min()) > 50
        for i in range(B):
            c = ref[s][2][i]
            assert out[s][0][i, :c].tobytes() == ref[s][0][i, :c].tobytes() and np.array_equal(out[s][1][i, :c], ref[s][1][i, :c])
    with pytest.raises(OrbxError):
        ex.extract_host_end()                                    # nothing in flight
    # a third begin completes the oldest; a synchronous call completes everything
    for s in range(3):
        for a in out[s]: a[...] = 0
        ex.extract_host_begin(imgs[s], *out[s])
    k = np.zeros((B, cap), KP_DTYPE); d = np.zeros((B, cap, 32), np.uint8); n = np.zeros(B, np.int32)
    ex.extract_host(imgs[0], k, d, n)
    for s in range(3):
        assert np.array_equal(out[s][2], ref[s][2])
    assert np.array_equal(n, ref[0][2])


def test_stereo_batches_in_flight_give_the_results_of_the_synchronous_call():
    """orbx_stereo_extract_batch_begin / _end: two batches of pairs in flight return what the blocking call returns."""
    from orb_slam2_commit_b200 import KP_DTYPE, stereo_extract_host, stereo_extract_host_begin, stereo_extract_host_end
    W, H, n = 480, 200, 24
    args = (600, 1.2, 5, 20, 7)
    exL, exR = ORBextractor(*args), ORBextractor(*args)
    cap = exL.reserve(W, H, n); assert exR.reserve(W, H, n) == cap
    bf, fx = 386.1448 * 0.4, 718.856 * 0.4

    def bufs():
        return dict(kl=np.zeros((n, cap), KP_DTYPE), kr=np.zeros((n, cap), KP_DTYPE), dl=np.zeros((n, cap, 32), np.uint8),
                    dr=np.zeros((n, cap, 32), np.uint8), nl=np.zeros(n, np.int32), nr=np.zeros(n, np.int32),
                    u_right=np.zeros((n, cap), np.float32), depth=np.zeros((n, cap), np.float32))
    batches = []
    for s in range(3):
        pairs = [synth.synth_stereo_pair(W, H, 900 + 30 * s + (i % 5), max_disp=40) for i in range(n)]
        batches.append((np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])))
    ref = []
    for L, R in batches:
        o = bufs(); stereo_extract_host(exL, exR, L, R, bf, fx, o); ref.append(o)
    out = [bufs() for _ in range(3)]
    stereo_extract_host_begin(exL, exR, *batches[0], bf, fx, out[0])
    stereo_extract_host_begin(exL, exR, *batches[1], bf, fx, out[1])
    stereo_extract_host_end(exL)
    stereo_extract_host_begin(exL, exR, *batches[2], bf, fx, out[2])
    stereo_extract_host_end(exL); stereo_extract_host_end(exL)
    for s in range(3):
        assert np.array_equal(out[s]["nl"], ref[s]["nl"]) and np.array_equal(out[s]["nr"], ref[s]["nr"])
        for i in range(n):
            nl = int(ref[s]["nl"][i])
            assert out[s]["kl"][i, :nl].tobytes() == ref[s]["kl"][i, :nl].tobytes()
            assert np.array_equal(out[s]["u_right"][i, :nl].view(np.uint32), ref[s]["u_right"][i, :nl].view(np.uint32))
            assert np.array_equal(out[s]["depth"][i, :nl].view(np.uint32), ref[s]["depth"][i, :nl].view(np.uint32))
        assert int(np.count_nonzero(ref[s]["u_right"] > 0)) > 20 * n


def test_pyramid_mirror_is_the_pyramid_with_its_apron():
    """orbx_pyramid_mirror: ONE copy of the frame's raw block; every level's header (payload pointer, device pitch as step)
    sees the payload and, by pointer arithmetic as in Frame.cc:681-700, the 19-px BORDER_REFLECT_101 apron around it."""
    import ctypes
    c = _cfg("tum1")
    img = synth.synth_image(c["width"], c["height"], 21)
    orc = ob.Extractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    orc.extract(img)
    for mirror_in_call in (True, False):
        ex = ORBextractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
        ex.set_pyramid_mirror(mirror_in_call)
        for _ in range(3):                                    # the third call replays the recorded CUDA graph
            ex(img)
        for l, v in enumerate(ex.pyramid_mirror(0)):
            want = orc.level(l)                               # (h+38) x (w+38), apron included
            h, w = v.shape
            assert want.shape == (h + 38, w + 38)
            assert np.array_equal(np.asarray(v), want[19:-19, 19:-19]), f"level {l}: payload"
            pitch = v.strides[0]
            addr = v.__array_interface__["data"][0] - 19 * pitch - 19
            flat = np.frombuffer((ctypes.c_uint8 * ((h + 38) * pitch)).from_address(addr), np.uint8)
            around = np.lib.stride_tricks.as_strided(flat, shape=(h + 38, w + 38), strides=(pitch, 1))
            assert np.array_equal(around, want), f"level {l}: apron around the mirrored payload"


def test_init_undistort_rectify_map_on_the_device_equals_cv2(golden_dir):
    """orbx_init_undistort_rectify_map (un-contracted f64 on the device) against the committed cv2 4.13 maps: bit for bit."""
    from orb_slam2_commit_b200 import init_undistort_rectify_map
    g = np.load(os.path.join(golden_dir, "prims3_cv2.npz"))
    names = sorted({k[:-5] for k in g.files if k.endswith("_size")})
    for n in names:
        size = tuple(int(x) for x in g[n + "_size"])
        m1, m2 = init_undistort_rectify_map(g[n + "_K"], g[n + "_D"], g[n + "_R"], g[n + "_P"], size)
        assert zlib.crc32(m1.tobytes()) == int(g[n + "_crc"][0]) and zlib.crc32(m2.tobytes()) == int(g[n + "_crc"][1]), n
        o1, o2 = ob.init_undistort_rectify_map(g[n + "_K"], g[n + "_D"], g[n + "_R"], g[n + "_P"], size)
        assert np.array_equal(m1, o1) and np.array_equal(m2, o2)


def test_raw_euroc_pairs_to_depth_in_one_call():
    """orbx_set_rectify_camera (initUndistortRectifyMap on the device, both EuRoC cameras) + orbx_stereo_extract_batch_rectified:
    unrectified pairs in, keypoints / descriptors / mvuRight / mvDepth out — against the oracle chain
    initUndistortRectifyMap -> remap -> ORBextractor x 2 -> ComputeStereoMatches (stereo_euroc.cc:96-97,136-137; Frame.cc:80-117)."""
    from orb_slam2_commit_b200 import api, stereo_extract_host_rectified
    c = _cfg("euroc")
    W, H = c["width"], c["height"]
    cfgargs = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    exL, exR = ORBextractor(*cfgargs), ORBextractor(*cfgargs)
    cams = (synth.EUROC_LEFT, synth.EUROC_RIGHT)
    for ex, cam in zip((exL, exR), cams):
        ex.set_rectify_camera(cam["K"], cam["D"], cam["R"], cam["P"], (W, H))
    n = 3
    pairs = [synth.synth_stereo_pair(W, H, 70 + i) for i in range(n)]
    L = np.stack([p[0] for p in pairs]); R = np.stack([p[1] for p in pairs])
    cap = exL.reserve(W, H, n); exR.reserve(W, H, n)
    out = dict(kl=np.zeros((n, cap), api.KP_DTYPE), kr=np.zeros((n, cap), api.KP_DTYPE), dl=np.zeros((n, cap, 32), np.uint8),
               dr=np.zeros((n, cap, 32), np.uint8), nl=np.zeros(n, np.int32), nr=np.zeros(n, np.int32),
               u_right=np.zeros((n, cap), np.float32), depth=np.zeros((n, cap), np.float32))
    stereo_extract_host_rectified(exL, exR, L, R, c["bf"], c["fx"], out)
    maps = [ob.init_undistort_rectify_map(cam["K"], cam["D"], cam["R"], cam["P"], (W, H)) for cam in cams]
    for i in range(n):
        oL, oR = ob.Extractor(*cfgargs), ob.Extractor(*cfgargs)
        k1, d1 = oL.extract(ob.remap(L[i], *maps[0])); k2, d2 = oR.extract(ob.remap(R[i], *maps[1]))
        ur, dp = ob.stereo_match(oL, oR, k1, d1, k2, d2, c["bf"], c["fx"])
        nl, nr = int(out["nl"][i]), int(out["nr"][i])
        assert nl == len(k1) and nr == len(k2)
        _check_against(out["kl"][i, :nl], out["dl"][i, :nl], k1, d1, f"left {i}")
        _check_against(out["kr"][i, :nr], out["dr"][i, :nr], k2, d2, f"right {i}")
        assert np.array_equal(out["u_right"][i, :nl].view(np.uint32), ur.view(np.uint32)), f"mvuRight of pair {i}"
        assert np.array_equal(out["depth"][i, :nl].view(np.uint32), dp.view(np.uint32)), f"mvDepth of pair {i}"
        assert (ur >= 0).sum() > 0        # (the synthetic pair is not a physical EuRoC pair: few rows still line up after rectification)


def test_euroc_stream_of_256_pairs_matches_the_oracle():
    """BASELINE config 3 at test size: a seeded stream of 256 EuRoC stereo pairs through the batched, chunk-pipelined host call
    (orbx_stereo_extract_batch; pairs shard by frame, so every pair is independent) — keypoints, descriptors, mvuRight and
    mvDepth of EVERY pair against the oracle (Frame.cc:80-84, 547-788). 32 distinct synthetic pairs, each reused with eight
    different gains so that all 256 pairs differ; the oracle runs pair-parallel on the host threads."""
    from concurrent.futures import ThreadPoolExecutor
    from orb_slam2_commit_b200 import api, stereo_extract_host
    c = _cfg("euroc")
    W, H = c["width"], c["height"]
    cfgargs = (c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"])
    base = [synth.synth_stereo_pair(W, H, 300 + i) for i in range(32)]
    n = 256
    L = np.empty((n, H, W), np.uint8); R = np.empty((n, H, W), np.uint8)
    for i in range(n):
        g = 1.0 - 0.04 * (i // 32)
        L[i] = np.clip(base[i % 32][0].astype(np.float32) * g + (i // 32), 0, 255).astype(np.uint8)
        R[i] = np.clip(base[i % 32][1].astype(np.float32) * g + (i // 32), 0, 255).astype(np.uint8)
    exL, exR = ORBextractor(*cfgargs), ORBextractor(*cfgargs)
    cap = exL.reserve(W, H, 64); exR.reserve(W, H, 64)
    out = dict(kl=np.zeros((n, cap), api.KP_DTYPE), kr=np.zeros((n, cap), api.KP_DTYPE), dl=np.zeros((n, cap, 32), np.uint8),
               dr=np.zeros((n, cap, 32), np.uint8), nl=np.zeros(n, np.int32), nr=np.zeros(n, np.int32),
               u_right=np.zeros((n, cap), np.float32), depth=np.zeros((n, cap), np.float32))
    stereo_extract_host(exL, exR, L, R, c["bf"], c["fx"], out)

    def oracle_pair(i):
        oL, oR = ob.Extractor(*cfgargs), ob.Extractor(*cfgargs)
        k1, d1 = oL.extract(L[i]); k2, d2 = oR.extract(R[i])
        ur, dp = ob.stereo_match(oL, oR, k1, d1, k2, d2, c["bf"], c["fx"])
        return k1, d1, k2, d2, ur, dp
    with ThreadPoolExecutor(os.cpu_count() or 8) as pool:
        want = list(pool.map(oracle_pair, range(n)))
    matched = 0
    for i, (k1, d1, k2, d2, ur, dp) in enumerate(want):
        nl, nr = int(out["nl"][i]), int(out["nr"][i])
        assert nl == len(k1) and nr == len(k2), f"pair {i}"
        _check_against(out["kl"][i, :nl], out["dl"][i, :nl], k1, d1, f"left {i}")
        _check_against(out["kr"][i, :nr], out["dr"][i, :nr], k2, d2, f"right {i}")
        assert np.array_equal(out["u_right"][i, :nl].view(np.uint32), ur.view(np.uint32)), f"mvuRight of pair {i}"
        assert np.array_equal(out["depth"][i, :nl].view(np.uint32), dp.view(np.uint32)), f"mvDepth of pair {i}"
        matched += int((ur >= 0).sum())
    assert matched > 100 * n
