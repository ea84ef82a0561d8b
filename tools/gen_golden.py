#!/usr/bin/env python3
"""Generate tests/golden/*.npz — runs ONLY in the authoring container (needs cv2 4.13 and /root/reference).

  prims_cv2.npz      inputs + outputs of the cv2 4.13 primitives the reference calls
                     (resize INTER_LINEAR, copyMakeBorder REFLECT_101, GaussianBlur 7x7 s2, FAST 9-16 NMS at
                     T=20 and T=7 on whole images and cell-sized crops, fastAtan2)
  prims2_cv2.npz     cv2.remap INTER_LINEAR through initUndistortRectifyMap maps, cv2.undistortPoints
  prims3_cv2.npz     cv2.initUndistortRectifyMap(K, D, R, P, size, CV_32FC1): EuRoC left / right cameras (map CRCs + every 8th
                     row), twelve random cameras with 4 / 5 / 8 / 12 distortion coefficients (CRCs), one small camera in full
  ref_<cfg>.npz      keypoints (28-byte cv::KeyPoint records) + descriptors + per-level pyramid CRCs produced by
                     oracle/_ref = the UNMODIFIED /root/reference/src/ORBextractor.cc (bump-allocator build)
  sincos.json        result of the exhaustive oc_cosf/oc_sinf == glibc cosf/sinf check
"""
import json, os, subprocess, sys, zlib
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import binding as ob
from orb_slam2_commit_b200 import synth

G = os.path.join(ROOT, "tests", "golden")
os.makedirs(G, exist_ok=True)
assert cv2.__version__.startswith("4.13"), cv2.__version__


def kps_to_arr(kps):
    return np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps],
                    dtype=ob.KP_DTYPE)


def prims():
    out = {"cv2_version": np.array(cv2.__version__)}
    rng = np.random.default_rng(123)
    # resize: the four BASELINE level-0 -> level-1 geometries at reduced height + odd sizes
    for i, (w, h, dw, dh) in enumerate([(640, 96, 533, 80), (1241, 60, 1034, 50), (752, 64, 627, 53),
                                        (97, 71, 81, 59), (214, 161, 179, 134), (63, 62, 52, 52)]):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"resize{i}_src"] = src
        out[f"resize{i}_dst"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    for i, (w, h) in enumerate([(160, 120), (67, 45), (9, 8)]):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        out[f"blur{i}_src"] = src
        out[f"blur{i}_dst"] = cv2.GaussianBlur(src, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        out[f"border{i}_dst"] = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101) if min(w, h) > 19 else \
            cv2.copyMakeBorder(src, 3, 3, 3, 3, cv2.BORDER_REFLECT_101)
    img = synth.synth_image(320, 240, 5)
    out["fast_img"] = img
    for T in (20, 7):
        det = cv2.FastFeatureDetector_create(T, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        out[f"fast_full_T{T}"] = kps_to_arr(det.detect(img))
        for j, (y0, x0, hh, ww) in enumerate([(16, 16, 36, 36), (100, 200, 37, 38), (3, 280, 24, 40), (200, 0, 40, 7)]):
            out[f"fast_crop{j}_T{T}"] = kps_to_arr(det.detect(np.ascontiguousarray(img[y0:y0 + hh, x0:x0 + ww])))
            out[f"fast_crop{j}_rect"] = np.array([y0, x0, hh, ww])
    ys = rng.integers(-400000, 400000, 4096).astype(np.float32)
    xs = rng.integers(-400000, 400000, 4096).astype(np.float32)
    ys[:64] = 0; xs[32:96] = 0; ys[96:128] = xs[96:128]
    out["atan2_y"] = ys; out["atan2_x"] = xs
    out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    col = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    out["gray_src"] = col
    out["gray_rgb"] = cv2.cvtColor(np.ascontiguousarray(col[..., :3]), cv2.COLOR_RGB2GRAY)
    out["gray_bgr"] = cv2.cvtColor(np.ascontiguousarray(col[..., :3]), cv2.COLOR_BGR2GRAY)
    out["gray_rgba"] = cv2.cvtColor(col, cv2.COLOR_RGBA2GRAY)
    out["gray_bgra"] = cv2.cvtColor(col, cv2.COLOR_BGRA2GRAY)
    np.savez_compressed(os.path.join(G, "prims_cv2.npz"), **out)
    print("prims_cv2.npz", {k: v.shape for k, v in out.items() if k.startswith("fast_full")})


# EuRoC MH_01 cam0 calibration as shipped with upstream ORB-SLAM2's EuRoC.yaml (LEFT.K / D / R / P), scaled by 1/4 so
# the golden maps stay small; TUM1.yaml intrinsics + distortion for undistortPoints.
EUROC_K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]])
EUROC_D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])
EUROC_R = np.array([[0.999966347530033, -0.001422739138722922, 0.008079580483432283],
                    [0.001365741834644127, 0.9999741760894847, 0.007055629199258132],
                    [-0.008089410156878961, -0.007044357138835809, 0.9999424675829176]])
EUROC_P = np.array([[435.2046959714599, 0, 367.4517211914062], [0, 435.2046959714599, 252.2008514404297], [0, 0, 1]])
TUM1_K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)
TUM1_D = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)


def prims2():
    """cv::remap (rectification, stereo_euroc.cc:136-137) and cv::undistortPoints (Frame.cc:471-538)."""
    out = {"cv2_version": np.array(cv2.__version__)}
    rng = np.random.default_rng(321)
    S = np.diag([0.25, 0.25, 1.0])
    W, H = 188, 120
    m1, m2 = cv2.initUndistortRectifyMap(S @ EUROC_K, EUROC_D, EUROC_R, S @ EUROC_P, (W, H), cv2.CV_32F)
    src = rng.integers(0, 256, (H, W), dtype=np.uint8)
    out["remap0_src"] = src; out["remap0_map1"] = m1; out["remap0_map2"] = m2
    out["remap0_dst"] = cv2.remap(src, m1, m2, cv2.INTER_LINEAR)
    # wild maps: out-of-range coordinates, negative values, exact .5/32 ties, source size != map size
    src = rng.integers(0, 256, (60, 80), dtype=np.uint8)
    m1 = (rng.random((48, 64)) * 100 - 10).astype(np.float32); m2 = (rng.random((48, 64)) * 80 - 10).astype(np.float32)
    m1[:4] = np.round(m1[:4] * 64) / 64; m2[:4] = np.round(m2[:4] * 64) / 64
    m1[4, :8] = [-1, -0.5, 79, 79.5, 80, -1.015625, 78.984375, 1e6]; m2[4, :8] = [-1, 59, 59.5, 60, -0.5, 3, 3, -1e6]
    out["remap1_src"] = src; out["remap1_map1"] = m1; out["remap1_map2"] = m2
    out["remap1_dst"] = cv2.remap(src, m1, m2, cv2.INTER_LINEAR)
    K = np.array([[TUM1_K4[0], 0, TUM1_K4[2]], [0, TUM1_K4[1], TUM1_K4[3]], [0, 0, 1]], np.float32)
    pts = np.stack([rng.random(2048) * 680 - 20, rng.random(2048) * 520 - 20], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]          # ComputeImageBounds corners
    out["undist_K4"] = TUM1_K4; out["undist_D"] = TUM1_D; out["undist_src"] = pts
    out["undist_dst5"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, TUM1_D, None, K).reshape(-1, 2)
    out["undist_dst4"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, TUM1_D[:4], None, K).reshape(-1, 2)
    np.savez_compressed(os.path.join(G, "prims2_cv2.npz"), **out)
    print("prims2_cv2.npz", os.path.getsize(os.path.join(G, "prims2_cv2.npz")))


def rectify_cameras():
    """The cameras of the initUndistortRectifyMap fixtures: (name, K, D, R, P, (w, h)); seeded, also used by the tests."""
    cams = [("euroc_left", synth.EUROC_LEFT["K"], synth.EUROC_LEFT["D"], synth.EUROC_LEFT["R"], synth.EUROC_LEFT["P"], (752, 480)),
            ("euroc_right", synth.EUROC_RIGHT["K"], synth.EUROC_RIGHT["D"], synth.EUROC_RIGHT["R"], synth.EUROC_RIGHT["P"], (752, 480))]
    rng = np.random.default_rng(7)
    for t in range(12):
        W, H = [(752, 480), (640, 480), (1241, 376), (320, 240)][t % 4]
        f = rng.uniform(300, 900)
        K = np.array([[f, 0, W / 2 + rng.uniform(-20, 20)], [0, f * rng.uniform(0.95, 1.05), H / 2 + rng.uniform(-20, 20)], [0, 0, 1]])
        nd = [4, 5, 8, 12][t % 4]
        D = np.zeros(nd); D[:2] = rng.uniform(-0.3, 0.3, 2); D[2:4] = rng.uniform(-0.005, 0.005, 2)
        if nd >= 5: D[4] = rng.uniform(-0.1, 0.1)
        if nd >= 8: D[5:8] = rng.uniform(-0.05, 0.05, 3)
        if nd >= 12: D[8:12] = rng.uniform(-0.002, 0.002, 4)
        R, _ = cv2.Rodrigues(rng.uniform(-0.02, 0.02, 3))
        fp = f * rng.uniform(0.9, 1.0)
        P = np.array([[fp, 0, W / 2 + rng.uniform(-5, 5), 0], [0, fp, H / 2 + rng.uniform(-5, 5), 0], [0, 0, 1, 0]])
        cams.append((f"rand{t}", K, D, R, P, (W, H)))
    cams.append(("small", cams[2][1] * np.array([[0.25, 1, 0.25], [1, 0.25, 0.25], [1, 1, 1]]), cams[2][2], cams[2][3],
                 cams[2][4] * np.array([[0.25, 1, 0.25, 1], [1, 0.25, 0.25, 1], [1, 1, 1, 1]]), (188, 120)))
    return cams


def prims3():
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, K, D, R, P, size in rectify_cameras():
        m1, m2 = cv2.initUndistortRectifyMap(np.asarray(K, np.float64), np.asarray(D, np.float64), np.asarray(R, np.float64),
                                             np.asarray(P, np.float64), size, cv2.CV_32FC1)
        for k, v in (("K", K), ("D", D), ("R", R), ("P", P)):
            out[f"{name}_{k}"] = np.asarray(v, np.float64)
        out[f"{name}_size"] = np.array(size)
        out[f"{name}_crc"] = np.array([zlib.crc32(m1.tobytes()), zlib.crc32(m2.tobytes())], np.int64)
        if name.startswith("euroc"):
            out[f"{name}_rows8_1"] = m1[::8].copy(); out[f"{name}_rows8_2"] = m2[::8].copy()
        if name == "small":
            out[f"{name}_map1"] = m1; out[f"{name}_map2"] = m2
    np.savez_compressed(os.path.join(G, "prims3_cv2.npz"), **out)
    print("prims3_cv2.npz", os.path.getsize(os.path.join(G, "prims3_cv2.npz")))


def ref_cases():
    assert ob.ref_available() or True
    ob.build(force=True)
    cases = [("tum1", 1), ("kitti", 2), ("euroc", 1000), ("small", 11)]
    for name, seed in cases:
        if name == "small":
            c = dict(width=200, height=150, nfeatures=150, scale=1.2, nlevels=4, ini_th=20, min_th=7)
        else:
            c = synth.CONFIGS[name]
        img = synth.synth_image(c["width"], c["height"], seed)
        r = ob.RefExtractor(c["nfeatures"], c["scale"], c["nlevels"], c["ini_th"], c["min_th"], deterministic=True)
        kps, desc = r.extract(img)
        crcs = np.array([zlib.crc32(r.level(l).tobytes()) for l in range(c["nlevels"])], np.uint32)
        sizes = np.array([r.level(l).shape for l in range(c["nlevels"])], np.int32)
        t = r.tables()
        np.savez_compressed(os.path.join(G, f"ref_{name}.npz"), seed=seed, img_crc=zlib.crc32(img.tobytes()),
                            cfg=json.dumps(c), keypoints=kps, descriptors=desc, level_crc=crcs, level_whole_shape=sizes,
                            **t)
        print(f"ref_{name}.npz", len(kps), sizes[0], hex(crcs[0]))


def sincos():
    exe = "/tmp/check_sincos"
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-mfma", "-o", exe,
                           os.path.join(ROOT, "oracle", "check_sincos.c"), os.path.join(ROOT, "oracle", "orb_oracle.c"),
                           "-I", os.path.join(ROOT, "oracle"), "-lm", "-lpthread"])
    res = json.loads(subprocess.check_output([exe, "1"]).decode())
    res["libm"] = subprocess.check_output(["ldd", "--version"]).decode().splitlines()[0]
    json.dump(res, open(os.path.join(G, "sincos.json"), "w"), indent=1)
    print(res)


if __name__ == "__main__":
    if "--prims3-only" in sys.argv:
        prims3()
        sys.exit(0)
    if "--prims2-only" in sys.argv:
        prims2()
        sys.exit(0)
    prims()
    prims2()
    prims3()
    ref_cases()
    if "--sincos" in sys.argv:
        sincos()
