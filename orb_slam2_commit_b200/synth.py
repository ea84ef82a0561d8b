"""Seeded synthetic inputs (numpy only, bit-reproducible on any host).

The reference ships no images and no datasets, so every test and bench line runs on
frames from this generator: smoothed noise stretched to 0..255 (corner-rich), random flat rectangles
(strong FAST corners + cells that need the minThFAST retry) and one large flat region (empty cells).
"""
from __future__ import annotations

import numpy as np

# extractor settings named by BASELINE.json `configs` (the YAML files are absent from the reference tree)
CONFIGS = {
    "tum1": dict(width=640, height=480, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7),
    "kitti": dict(width=1241, height=376, nfeatures=2000, scale=1.2, nlevels=8, ini_th=20, min_th=7,
                  fx=718.856, bf=386.1448),
    "euroc": dict(width=752, height=480, nfeatures=1200, scale=1.2, nlevels=8, ini_th=20, min_th=7,
                  fx=435.2047, bf=47.90639384423901),
    "4k": dict(width=3840, height=2160, nfeatures=8000, scale=1.2, nlevels=12, ini_th=20, min_th=7),
}


def _smooth(a: np.ndarray, sigma: float) -> np.ndarray:
    """Separable Gaussian with reflect-101 borders, summed in a fixed order (deterministic float64)."""
    r = int(3 * sigma + 0.5)
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (r, r)
        p = np.pad(a, pad, mode="reflect")
        out = np.zeros_like(a)
        n = a.shape[axis]
        for i in range(2 * r + 1):
            sl = [slice(None), slice(None)]
            sl[axis] = slice(i, i + n)
            out += k[i] * p[tuple(sl)]
        a = out
    return a


def synth_image(width: int, height: int, seed: int, sigma: float = 1.6) -> np.ndarray:
    """One 8-bit grayscale frame, C-contiguous (height, width)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (height, width)).astype(np.float64)
    s = _smooth(base, sigma)
    lo, hi = s.min(), s.max()
    img = np.clip((s - lo) * (255.0 / (hi - lo)), 0, 255)
    img = np.floor(img + 0.5).astype(np.uint8)
    nrect = max(4, (width * height) // 1500)
    xs = rng.integers(0, width, nrect)
    ys = rng.integers(0, height, nrect)
    ws = rng.integers(5, 61, nrect)
    hs = rng.integers(5, 61, nrect)
    gs = rng.integers(0, 256, nrect)
    keep = rng.random(nrect) < 0.35  # most of the textured background stays visible
    for x, y, w, h, g, k in zip(xs, ys, ws, hs, gs, keep):
        if k:
            img[y:y + h, x:x + w] = g
    # one large flat region -> cells with no corner at either threshold
    fw, fh = width // 5, height // 4
    fx = int(rng.integers(0, width - fw))
    fy = int(rng.integers(0, height - fh))
    img[fy:fy + fh, fx:fx + fw] = int(rng.integers(0, 256))
    # a low-contrast patch: corners only at the min threshold
    lw, lh = width // 6, height // 5
    lx = int(rng.integers(0, width - lw))
    ly = int(rng.integers(0, height - lh))
    patch = img[ly:ly + lh, lx:lx + lw].astype(np.int32)
    img[ly:ly + lh, lx:lx + lw] = (128 + (patch - 128) // 6).astype(np.uint8)
    return np.ascontiguousarray(img)


def synth_stereo_pair(width: int, height: int, seed: int, max_disp: int = 60):
    """Left frame + a right frame = left shifted by a per-row-band disparity, plus noise (sigma 2)."""
    left = synth_image(width, height, seed)
    rng = np.random.default_rng(seed + 7_000_003)
    right = np.empty_like(left)
    band = 24
    for y0 in range(0, height, band):
        d = int(rng.integers(0, max_disp + 1))
        rows = left[y0:y0 + band]
        shifted = np.empty_like(rows)
        if d:
            shifted[:, :width - d] = rows[:, d:]
            shifted[:, width - d:] = rows[:, -1:]
        else:
            shifted[:] = rows
        right[y0:y0 + band] = shifted
    noise = np.floor(rng.normal(0.0, 2.0, right.shape) + 0.5).astype(np.int32)
    right = np.clip(right.astype(np.int32) + noise, 0, 255).astype(np.uint8)
    return left, np.ascontiguousarray(right)


def synth_descriptors(n_train: int, n_query: int, seed: int = 42, max_flips: int = 40, n_dup: int = 8):
    """Config 4 inputs: random 256-bit train rows; queries = train rows with 0..max_flips bit flips.
    A few train rows are duplicated so exact distance ties (lowest index must win) occur."""
    rng = np.random.default_rng(seed)
    train = rng.integers(0, 256, (n_train, 32), dtype=np.uint8)
    rng2 = np.random.default_rng(seed + 1)
    perm = rng2.integers(0, n_train, n_query)
    query = train[perm].copy()
    flips = rng2.integers(0, max_flips + 1, n_query)
    for i in range(n_query):
        if flips[i]:
            bits = rng2.choice(256, int(flips[i]), replace=False)
            np.bitwise_xor.at(query[i], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    for j in range(min(n_dup, n_query, n_train // 2)):
        src = int(perm[j])
        dst = (src + 1 + j * 7919) % n_train
        train[dst] = train[src]
    return np.ascontiguousarray(train), np.ascontiguousarray(query)


# Camera models as shipped with upstream ORB-SLAM2's example settings (the YAMLs are not in the reference snapshot):
# EuRoC.yaml LEFT.K / LEFT.D / LEFT.R / LEFT.P (stereo rectification, Examples/Stereo/stereo_euroc.cc:70-98) and
# TUM1.yaml Camera.* (Frame::UndistortKeyPoints, Frame.cc:471-506).
EUROC_LEFT = dict(
    K=np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]]),
    D=np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0]),
    R=np.array([[0.999966347530033, -0.001422739138722922, 0.008079580483432283],
                [0.001365741834644127, 0.9999741760894847, 0.007055629199258132],
                [-0.008089410156878961, -0.007044357138835809, 0.9999424675829176]]),
    P=np.array([[435.2046959714599, 0, 367.4517211914062], [0, 435.2046959714599, 252.2008514404297], [0, 0, 1]]))
EUROC_RIGHT = dict(
    K=np.array([[457.587, 0, 379.999], [0, 456.134, 255.238], [0, 0, 1]]),
    D=np.array([-0.28368365, 0.07451284, -0.00010473, -3.555907e-05, 0.0]),
    R=np.array([[0.9999633526194376, -0.003625811871560086, 0.007755443660172947],
                [0.003680398547259526, 0.9999684752771629, -0.007035845251224894],
                [-0.007729688520722713, 0.007064130529506649, 0.999945173484644]]),
    P=np.array([[435.2046959714599, 0, 367.4517211914062, -47.90639384423901], [0, 435.2046959714599, 252.2008514404297, 0], [0, 0, 1, 0]]))
TUM1_K4 = np.array([517.306408, 516.469215, 318.643040, 255.313989], np.float32)
TUM1_DIST = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)


def rectify_maps(width: int, height: int, cam: dict = EUROC_LEFT, scale: float = 1.0):
    """The CV_32FC1 map pair cv::initUndistortRectifyMap(K, D, R, P, size, CV_32F) would hand to cv::remap: for every
    rectified pixel the source coordinate (inverse mapping through R^-1, the distortion model and K). float64 numpy,
    rounded to float32 once — inputs for the remap parity tests, not a bit-exact copy of OpenCV's map builder."""
    S = np.diag([scale, scale, 1.0])
    K = S @ cam["K"]; P = S @ cam["P"]
    k1, k2, p1, p2, k3 = cam["D"]
    iR = np.linalg.inv(P @ cam["R"])
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    X = iR[0, 0] * u + iR[0, 1] * v + iR[0, 2]
    Y = iR[1, 0] * u + iR[1, 1] * v + iR[1, 2]
    W = iR[2, 0] * u + iR[2, 1] * v + iR[2, 2]
    x = X / W; y = Y / W
    r2 = x * x + y * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * kr + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return (K[0, 0] * xd + K[0, 2]).astype(np.float32), (K[1, 1] * yd + K[1, 2]).astype(np.float32)


def synth_vocabulary(k: int, L: int, seed: int, early_leaf: float = 0.03, stop_frac: float = 0.02):
    """A DBoW2-style vocabulary tree in text-file order (ORBvoc.txt is not in the reference snapshot): returns
    (parent[int32], is_leaf[uint8], desc[n,32] uint8, weight[float64]) for the non-root nodes, numbered level by level
    so that siblings are consecutive (as DBoW2's HKmeansStep creates them). Child descriptor = parent descriptor with
    fewer and fewer random bit flips per level; a few nodes above depth L are leaves, a few have fewer than k children,
    and `stop_frac` of the words carry weight 0 (stopped words)."""
    rng = np.random.default_rng(seed)
    parents, leaves, descs, weights = [], [], [], []
    cur_ids = np.array([0]); cur_desc = np.zeros((1, 32), np.uint8)
    next_id = 1
    for depth in range(1, L + 1):
        nch = np.full(len(cur_ids), k)
        if depth > 1:
            nch -= (rng.random(len(cur_ids)) < 0.1) * rng.integers(1, max(2, k // 2), len(cur_ids))
        par = np.repeat(cur_ids, nch)
        pdesc = np.repeat(cur_desc, nch, axis=0)
        p = 0.5 if depth == 1 else 0.5 / (1.7 ** (depth - 1))
        flips = np.packbits(rng.random((len(par), 256)) < p, axis=1, bitorder="little")
        d = pdesc ^ flips
        leaf = np.ones(len(par), np.uint8) if depth == L else (rng.random(len(par)) < early_leaf).astype(np.uint8)
        parents.append(par); leaves.append(leaf); descs.append(d)
        w = rng.random(len(par)) * 9.5 + 0.1
        w[rng.random(len(par)) < stop_frac] = 0.0
        weights.append(np.where(leaf > 0, w, 0.0))
        ids = np.arange(next_id, next_id + len(par)); next_id += len(par)
        keep = leaf == 0
        cur_ids = ids[keep]; cur_desc = d[keep]
    return (np.concatenate(parents).astype(np.int32), np.concatenate(leaves), np.concatenate(descs),
            np.concatenate(weights))


def synth_features_near_words(voc, n: int, seed: int, max_flips: int = 30):
    """n descriptors = random leaf descriptors of `voc` with up to max_flips random bit flips."""
    parent, is_leaf, desc, _ = voc
    rng = np.random.default_rng(seed)
    leaf_idx = np.flatnonzero(is_leaf)
    pick = leaf_idx[rng.integers(0, len(leaf_idx), n)]
    out = desc[pick].copy()
    for i in range(n):
        for b in rng.integers(0, 256, rng.integers(0, max_flips + 1)):
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def synth_tracking_scene(seed: int, n_last: int = 1000, n_extra: int = 400, cluster: float = 0.3, stereo: bool = True):
    """Inputs of ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1489-1646) for a TUM1-like
    camera: map points in front of the camera, a last frame that observes them and a current frame a small motion away.
    `cluster` of the map points are near-duplicates (same place, near-identical descriptors) so that several of them
    compete for one current keypoint — the case where the reference's sequential "already taken" rule matters.
    Returns a dict with the arguments of api.search_by_projection_frame / oracle.search_by_projection_frame."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, mbf = 517.306408, 516.469215, 318.643040, 255.313989, 40.0
    W, H = 640.0, 480.0
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                         ("octave", "<i4"), ("class_id", "<i4")])
    n_base = max(1, int(n_last * (1 - cluster)))
    z = rng.uniform(1.5, 12.0, n_base)
    xyz = np.stack([(rng.uniform(20, W - 20, n_base) - cx) / fx * z, (rng.uniform(20, H - 20, n_base) - cy) / fy * z, z], 1)
    desc = rng.integers(0, 256, (n_base, 32), dtype=np.uint8)
    dup = rng.integers(0, n_base, n_last - n_base)
    xyz = np.concatenate([xyz, xyz[dup] + rng.normal(0, 0.002, (len(dup), 3))]).astype(np.float32)
    ddup = desc[dup].copy()
    for i in range(len(dup)):
        for b in rng.integers(0, 256, rng.integers(0, 4)):
            ddup[i, b >> 3] ^= np.uint8(1 << (b & 7))
    mp_desc = np.concatenate([desc, ddup])
    order = rng.permutation(n_last)
    xyz, mp_desc = xyz[order], mp_desc[order]
    last = np.zeros(n_last, kp_dtype)
    last["x"] = fx * xyz[:, 0] / xyz[:, 2] + cx; last["y"] = fy * xyz[:, 1] / xyz[:, 2] + cy
    last["octave"] = rng.integers(0, 8, n_last); last["angle"] = rng.uniform(0, 360, n_last).astype(np.float32)
    flags = (rng.random(n_last) < 0.92).astype(np.uint8) | ((rng.random(n_last) < 0.7).astype(np.uint8) << 1)
    # current pose: small rotation about y and a forward / sideways translation
    a = 0.01
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([0.03, -0.01, -0.08])
    Tcw = np.concatenate([R.ravel(), t]).astype(np.float32)
    pc = xyz.astype(np.float64) @ R.T + t
    u = fx * pc[:, 0] / pc[:, 2] + cx; v = fy * pc[:, 1] / pc[:, 2] + cy
    # current keypoints: one per BASE map point (duplicates share it), jittered, plus unrelated ones
    base_of = np.empty(n_last, np.int64); base_of[order] = np.concatenate([np.arange(n_base), dup])
    first = {}
    for i in range(n_last):
        first.setdefault(base_of[i], i)
    src = np.array(sorted(first.values()))
    n_cur = len(src) + n_extra
    cur = np.zeros(n_cur, kp_dtype)
    cur["x"][:len(src)] = u[src] + rng.normal(0, 1.5, len(src)); cur["y"][:len(src)] = v[src] + rng.normal(0, 1.5, len(src))
    cur["octave"][:len(src)] = np.clip(last["octave"][src] + rng.integers(-1, 2, len(src)), 0, 7)
    cur["angle"][:len(src)] = (last["angle"][src] + rng.normal(4, 3, len(src)) + (rng.random(len(src)) < 0.1) * rng.uniform(0, 360, len(src))) % 360
    cur["x"][len(src):] = rng.uniform(0, W, n_extra); cur["y"][len(src):] = rng.uniform(0, H, n_extra)
    cur["octave"][len(src):] = rng.integers(0, 8, n_extra); cur["angle"][len(src):] = rng.uniform(0, 360, n_extra)
    cdesc = np.concatenate([mp_desc[src], rng.integers(0, 256, (n_extra, 32), dtype=np.uint8)])
    for i in range(len(src)):
        for b in rng.integers(0, 256, rng.integers(0, 60)):
            cdesc[i, b >> 3] ^= np.uint8(1 << (b & 7))
    perm = rng.permutation(n_cur)
    cur, cdesc = cur[perm], cdesc[perm]
    ur = None
    if stereo:
        zc = np.concatenate([pc[src, 2], rng.uniform(1.5, 12, n_extra)])[perm]
        ur = (cur["x"] - mbf / zc + rng.normal(0, 0.7, n_cur)).astype(np.float32)
        ur[rng.random(n_cur) < 0.25] = -1.0
    occ = (rng.random(n_cur) < 0.03).astype(np.uint8)
    cam9 = np.array([fx, fy, cx, cy, mbf, 0.0, W, 0.0, H], np.float32)
    return dict(cur_kps=cur, cur_desc=cdesc, cur_u_right=ur, cur_occupied=occ, Tcw12=Tcw, cam9=cam9, scale_factors=sf,
                last_kps=last, last_xyz=xyz, last_desc=mp_desc, last_flags=flags)


TRACKQ_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4")])


def synth_local_points_scene(seed: int, n_points: int = 1500, n_extra: int = 400, cluster: float = 0.3, stereo: bool = True):
    """Inputs of ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:46-142) as Tracking::SearchLocalPoints
    calls it: local map points with the fields Frame::isInFrustum leaves in them, and a frame whose keypoints observe most
    of them. `cluster` of the points are near-duplicates competing for one keypoint (the sequential rule), some keypoints
    carry a second near-identical neighbour on the same level (the NN-ratio rule).
    Returns the keyword arguments of api.search_local_points / oracle.search_local_points (without th / nnratio)."""
    sc = synth_tracking_scene(seed, n_last=n_points, n_extra=n_extra, cluster=cluster, stereo=stereo)
    rng = np.random.default_rng(seed + 7919)
    fx, fy, cx, cy, mbf = [float(v) for v in sc["cam9"][:5]]
    T = sc["Tcw12"].astype(np.float64); R = T[:9].reshape(3, 3); t = T[9:]
    pc = sc["last_xyz"].astype(np.float64) @ R.T + t
    q = np.zeros(n_points, TRACKQ_DTYPE)
    q["proj_x"] = fx * pc[:, 0] / pc[:, 2] + cx; q["proj_y"] = fy * pc[:, 1] / pc[:, 2] + cy
    q["proj_xr"] = q["proj_x"] - mbf / pc[:, 2]
    q["view_cos"] = np.where(rng.random(n_points) < 0.5, rng.uniform(0.9985, 1.0, n_points), rng.uniform(0.5, 0.9975, n_points))
    q["level"] = sc["last_kps"]["octave"]
    flags = (rng.random(n_points) < 0.9).astype(np.uint8) | ((rng.random(n_points) < 0.8).astype(np.uint8) << 1)
    kps, desc = sc["cur_kps"].copy(), sc["cur_desc"].copy()
    # second-best competitors: clone some keypoints (same level, 1-2 px away, a few bits flipped)
    n_twin = len(kps) // 8
    src = rng.choice(len(kps), n_twin, replace=False)
    twin = kps[src].copy(); twin["x"] += rng.normal(0, 1.5, n_twin).astype(np.float32); twin["y"] += rng.normal(0, 1.5, n_twin).astype(np.float32)
    twin["octave"] = np.where(rng.random(n_twin) < 0.7, twin["octave"], np.clip(twin["octave"] - 1, 0, 7))
    tdesc = desc[src].copy()
    for i in range(n_twin):
        for b in rng.integers(0, 256, rng.integers(0, 12)):
            tdesc[i, b >> 3] ^= np.uint8(1 << (b & 7))
    kps = np.concatenate([kps, twin]); desc = np.concatenate([desc, tdesc])
    ur = None
    if stereo:
        ur = np.concatenate([sc["cur_u_right"], sc["cur_u_right"][src]]).astype(np.float32)
    occ = np.concatenate([sc["cur_occupied"], np.zeros(n_twin, np.uint8)])
    perm = rng.permutation(len(kps))
    kps, desc, occ = kps[perm], desc[perm], occ[perm]
    if ur is not None:
        ur = ur[perm]
    bounds4 = np.array([sc["cam9"][5], sc["cam9"][6], sc["cam9"][7], sc["cam9"][8]], np.float32)
    return dict(kps=kps, desc=desc, u_right=ur, occupied=occ, bounds4=bounds4, scale_factors=sc["scale_factors"], queries=q,
                query_desc=sc["last_desc"], query_flags=flags)


def synth_fuse_scene(seed: int, n_points: int = 1200, n_extra: int = 500, stereo: bool = True):
    """Inputs of the search half of ORBmatcher::Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:918-1092): a keyframe (pose, camera,
    undistorted keypoints, mvuRight) and map points of its neighbours — most seen by the keyframe at a plausible scale,
    some behind the camera, outside the image, outside their distance-invariance range or seen from a wrong angle.
    Returns the keyword arguments of api.fuse_search / oracle.fuse_search (without th / mode)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, mbf = 517.306408, 516.469215, 318.643040, 255.313989, 40.0
    W, H = 640.0, 480.0
    nlevels = 8
    sf = (np.float32(1.2) ** np.arange(nlevels)).astype(np.float32)
    inv_s2 = (np.float32(1.0) / (sf * sf)).astype(np.float32)
    log_sf = np.float32(np.log(np.float32(1.2)))
    a, b = 0.05, -0.03
    Ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    R = (Ry @ Rx).astype(np.float32); t = np.array([0.2, -0.1, 0.3], np.float32)
    Ow = (-R.T.astype(np.float64) @ t.astype(np.float64)).astype(np.float32)
    Tcw = np.concatenate([R.ravel(), t]).astype(np.float32)
    z = rng.uniform(1.0, 15.0, n_points)
    pc = np.stack([(rng.uniform(-40, W + 40, n_points) - cx) / fx * z, (rng.uniform(-40, H + 40, n_points) - cy) / fy * z, z], 1)
    pc[rng.random(n_points) < 0.04, 2] *= -1                                       # behind the camera
    xyz = ((pc - t.astype(np.float64)) @ R.astype(np.float64)).astype(np.float32)   # Rcw^T (pc - tcw)
    PO = xyz.astype(np.float64) - Ow
    dist = np.linalg.norm(PO, axis=1)
    normal = PO / dist[:, None] + rng.normal(0, 0.25, (n_points, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    flip = rng.random(n_points) < 0.08
    normal[flip] *= -1                                                              # seen from behind
    lvl = rng.integers(0, nlevels, n_points)
    max_d = (dist * sf[lvl] * rng.uniform(0.85, 1.15, n_points)).astype(np.float32)  # mfMaxDistance = dist * scale of the reference octave
    min_d = (max_d / sf[nlevels - 1]).astype(np.float32)
    pt_dist = np.stack([np.float32(0.8) * min_d, np.float32(1.2) * max_d, max_d], 1).astype(np.float32)
    far = rng.random(n_points) < 0.06
    pt_dist[far, 1] = (dist[far] * 0.7).astype(np.float32)                          # outside the invariance range
    pt_desc = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)
    flags = (rng.random(n_points) < 0.9).astype(np.uint8)
    # keyframe keypoints: one near the projection of most points, at the predicted level or one below
    u = fx * pc[:, 0] / pc[:, 2] + cx; v = fy * pc[:, 1] / pc[:, 2] + cy
    ok = (pc[:, 2] > 0) & (u > 0) & (u < W) & (v > 0) & (v < H)
    src = np.flatnonzero(ok & (rng.random(n_points) < 0.85))
    n = len(src) + n_extra
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                         ("octave", "<i4"), ("class_id", "<i4")])
    kps = np.zeros(n, kp_dtype)
    ratio = max_d[src] / dist[src].astype(np.float32)
    pred = np.clip(np.ceil(np.log(ratio) / log_sf), 0, nlevels - 1).astype(np.int64)
    kps["octave"][:len(src)] = np.clip(pred - rng.integers(0, 3, len(src)) + (rng.random(len(src)) < 0.1), 0, nlevels - 1)
    jit = sf[kps["octave"][:len(src)]] * 1.2
    kps["x"][:len(src)] = u[src] + rng.normal(0, 1, len(src)) * jit; kps["y"][:len(src)] = v[src] + rng.normal(0, 1, len(src)) * jit
    kps["x"][len(src):] = rng.uniform(0, W, n_extra); kps["y"][len(src):] = rng.uniform(0, H, n_extra)
    kps["octave"][len(src):] = rng.integers(0, nlevels, n_extra)
    kps["angle"] = rng.uniform(0, 360, n)
    desc = np.concatenate([pt_desc[src], rng.integers(0, 256, (n_extra, 32), dtype=np.uint8)])
    for i in range(len(src)):
        for bb in rng.integers(0, 256, rng.integers(0, 70)):
            desc[i, bb >> 3] ^= np.uint8(1 << (bb & 7))
    ur = None
    if stereo:
        zc = np.concatenate([pc[src, 2], rng.uniform(1, 15, n_extra)])
        ur = (kps["x"] - mbf / zc + rng.normal(0, 0.8, n)).astype(np.float32)
        ur[rng.random(n) < 0.3] = -1.0
    perm = rng.permutation(n)
    kps, desc = kps[perm], desc[perm]
    if ur is not None:
        ur = ur[perm]
    cam9 = np.array([fx, fy, cx, cy, mbf, 0.0, W, 0.0, H], np.float32)
    return dict(kps=kps, desc=desc, u_right=ur, Tcw12=Tcw, Ow3=Ow, cam9=cam9, scale_factors=sf, inv_level_sigma2=inv_s2,
                log_scale_factor=float(log_sf), pt_xyz=xyz, pt_normal=normal.astype(np.float32), pt_dist=pt_dist, pt_desc=pt_desc,
                pt_flags=flags)


def synth_triangulation_scene(voc, seed: int, n_points: int = 1200, n_extra: int = 300, stereo: bool = False):
    """Inputs of ORBmatcher::SearchForTriangulation (ORBmatcher.cc:738-916): two keyframes seeing the same 3-D points from
    two poses (descriptors near vocabulary words so that matching features share FeatureVector nodes), F12 from the poses,
    some features already holding map points, some epipolar outliers. `voc` = (parent, is_leaf, desc, weight) of
    synth_vocabulary. Returns the keyword arguments of ORBVocabulary.search_for_triangulation (without flags)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, mbf = 517.306408, 516.469215, 318.643040, 255.313989, 40.0
    W, H = 640.0, 480.0
    nlevels = 8
    sf = (np.float32(1.2) ** np.arange(nlevels)).astype(np.float32)
    s2 = (sf * sf).astype(np.float32)
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    a = 0.06
    R1 = np.eye(3); t1 = np.zeros(3)
    R2 = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); t2 = np.array([-0.35, 0.02, 0.05])
    # F12 as LocalMapping::ComputeF12 builds it: K1^-T * [t12]x * R12 * K2^-1 with R12 = R1w R2w^T, t12 = -R12 t2w + t1w
    R12 = R1 @ R2.T; t12 = -R12 @ t2 + t1
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = (np.linalg.inv(K).T @ tx @ R12 @ np.linalg.inv(K)).astype(np.float32)
    Cw1 = (-R1.T @ t1).astype(np.float32)
    z = rng.uniform(1.5, 10.0, n_points)
    Xw = np.stack([(rng.uniform(10, W - 10, n_points) - cx) / fx * z, (rng.uniform(10, H - 10, n_points) - cy) / fy * z, z], 1)
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                         ("octave", "<i4"), ("class_id", "<i4")])
    base_desc = synth_features_near_words(voc, n_points, seed + 1, max_flips=6)
    out = {}
    for tag, R, t in (("1", R1, t1), ("2", R2, t2)):
        pc = Xw @ R.T + t
        u = fx * pc[:, 0] / pc[:, 2] + cx; v = fy * pc[:, 1] / pc[:, 2] + cy
        vis = np.flatnonzero((pc[:, 2] > 0.2) & (u > 2) & (u < W - 2) & (v > 2) & (v < H - 2))
        n = len(vis) + n_extra
        kps = np.zeros(n, kp_dtype)
        oct_ = rng.integers(0, nlevels, n)
        noise = rng.normal(0, 0.6, (len(vis), 2)) * sf[oct_[:len(vis)], None]
        outl = rng.random(len(vis)) < 0.1
        noise[outl] += rng.normal(0, 25, (int(outl.sum()), 2))                     # epipolar outliers
        kps["x"][:len(vis)] = u[vis] + noise[:, 0]; kps["y"][:len(vis)] = v[vis] + noise[:, 1]
        kps["x"][len(vis):] = rng.uniform(0, W, n_extra); kps["y"][len(vis):] = rng.uniform(0, H, n_extra)
        kps["octave"] = oct_
        kps["angle"] = np.concatenate([(37.0 * vis + rng.normal(3 if tag == "2" else 0, 3, len(vis))) % 360, rng.uniform(0, 360, n_extra)])
        d = np.concatenate([base_desc[vis], synth_features_near_words(voc, n_extra, seed + 10 + int(tag), max_flips=6)])
        for i in range(len(vis)):
            for b in rng.integers(0, 256, rng.integers(0, 14)):
                d[i, b >> 3] ^= np.uint8(1 << (b & 7))
        ur = None
        if stereo:
            zc = np.concatenate([pc[vis, 2], rng.uniform(1.5, 10, n_extra)])
            ur = (kps["x"] - mbf / zc).astype(np.float32); ur[rng.random(n) < 0.5] = -1.0
        perm = rng.permutation(n)
        out["kps" + tag] = kps[perm]; out["desc" + tag] = d[perm]
        out["has_mp" + tag] = (rng.random(n) < 0.35).astype(np.uint8)
        out["u_right" + tag] = None if ur is None else ur[perm]
    geom = np.concatenate([F12.ravel(), Cw1, R2.astype(np.float32).ravel(), t2.astype(np.float32), np.array([fx, fy, cx, cy], np.float32)])
    out.update(geom28=geom.astype(np.float32), scale_factors=sf, level_sigma2=s2)
    return out


def synth_kf_projection_scene(seed: int, n_points: int = 1200, n_extra: int = 500, cluster: float = 0.3):
    """Inputs of ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1648-1795) and
    of SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (:327-440): the scene of synth_fuse_scene plus near-duplicate
    points that compete for one feature (the sequential occupancy rule), features that are occupied beforehand and the
    keyframe-side angles for the rotation check. Returns the keyword arguments of api.search_by_projection_kf (without th,
    max_dist, mode, check_orientation)."""
    s = synth_fuse_scene(seed, n_points=n_points, n_extra=n_extra, stereo=False)
    rng = np.random.default_rng(seed + 104729)
    n_dup = int(n_points * cluster)
    src = rng.integers(0, n_points, n_dup)
    xyz = np.concatenate([s["pt_xyz"], s["pt_xyz"][src] + rng.normal(0, 0.002, (n_dup, 3)).astype(np.float32)])
    d = s["pt_desc"][src].copy()
    for i in range(n_dup):
        for b in rng.integers(0, 256, rng.integers(0, 5)):
            d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    order = rng.permutation(n_points + n_dup)
    cat = lambda a, b: np.concatenate([a, b])[order]
    out = dict(kps=s["kps"], desc=s["desc"], occupied=(rng.random(len(s["kps"])) < 0.08).astype(np.uint8), Tcw12=s["Tcw12"], Ow3=s["Ow3"],
               cam9=s["cam9"], scale_factors=s["scale_factors"], log_scale_factor=s["log_scale_factor"],
               pt_xyz=xyz[order], pt_normal=cat(s["pt_normal"], s["pt_normal"][src]), pt_dist=cat(s["pt_dist"], s["pt_dist"][src]),
               pt_desc=cat(s["pt_desc"], d), pt_flags=cat(s["pt_flags"], s["pt_flags"][src]),
               pt_angle=None)
    # angles: features around 10 deg, keyframe side around 45 deg (10 % anywhere), so that the rotation histogram has dominant bins
    kps = out["kps"].copy(); kps["angle"] = rng.normal(10, 8, len(kps)) % 360; out["kps"] = kps
    npt = n_points + n_dup
    out["pt_angle"] = np.where(rng.random(npt) < 0.1, rng.uniform(0, 360, npt), rng.normal(45, 12, npt) % 360).astype(np.float32)
    return out


def synth_sim3_scene(seed: int, n_points: int = 1500, n_extra: int = 300):
    """Inputs of ORBmatcher::SearchBySim3 (ORBmatcher.cc:1238-1487): two keyframes of one camera that see the same map points
    from two poses related by a similarity (scale drift 3 %). Every feature of a keyframe may hold a map point (mp_*).
    Returns (kf1, kf2, S12, S21, cam9, scale_factors, log_scale_factor)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy, mbf = 517.306408, 516.469215, 318.643040, 255.313989, 40.0
    W, H = 640.0, 480.0
    nlevels = 8
    sf = (np.float32(1.2) ** np.arange(nlevels)).astype(np.float32)
    log_sf = np.float32(np.log(np.float32(1.2)))
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                         ("octave", "<i4"), ("class_id", "<i4")])
    a = 0.08
    R1 = np.eye(3); t1 = np.array([0.05, 0.0, 0.1])
    R2 = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]); t2 = np.array([-0.4, 0.03, 0.2])
    z = rng.uniform(2.0, 12.0, n_points)
    Xw = np.stack([(rng.uniform(-60, W + 60, n_points) - cx) / fx * z, (rng.uniform(-60, H + 60, n_points) - cy) / fy * z, z], 1)
    pdesc = rng.integers(0, 256, (n_points, 32), dtype=np.uint8)
    lvl = rng.integers(0, nlevels, n_points)
    kfs = []
    for R, t in ((R1, t1), (R2, t2)):
        pc = Xw @ R.T + t
        dist = np.linalg.norm(pc, axis=1)
        u = fx * pc[:, 0] / pc[:, 2] + cx; v = fy * pc[:, 1] / pc[:, 2] + cy
        vis = np.flatnonzero((pc[:, 2] > 0.3) & (u > 1) & (u < W - 1) & (v > 1) & (v < H - 1) & (rng.random(n_points) < 0.85))
        n = len(vis) + n_extra
        kps = np.zeros(n, kp_dtype)
        max_d = (dist * sf[lvl] * rng.uniform(0.9, 1.1, n_points)).astype(np.float32)
        ratio = max_d[vis] / dist[vis].astype(np.float32)
        pred = np.clip(np.ceil(np.log(ratio) / log_sf), 0, nlevels - 1).astype(np.int64)
        kps["octave"][:len(vis)] = np.clip(pred - rng.integers(0, 3, len(vis)), 0, nlevels - 1)
        jit = sf[kps["octave"][:len(vis)]]
        kps["x"][:len(vis)] = u[vis] + rng.normal(0, 1.0, len(vis)) * jit; kps["y"][:len(vis)] = v[vis] + rng.normal(0, 1.0, len(vis)) * jit
        kps["x"][len(vis):] = rng.uniform(0, W, n_extra); kps["y"][len(vis):] = rng.uniform(0, H, n_extra)
        kps["octave"][len(vis):] = rng.integers(0, nlevels, n_extra)
        desc = np.concatenate([pdesc[vis], rng.integers(0, 256, (n_extra, 32), dtype=np.uint8)])
        for i in range(len(vis)):
            for b in rng.integers(0, 256, rng.integers(0, 80)):
                desc[i, b >> 3] ^= np.uint8(1 << (b & 7))
        # the map point each feature holds: its own point for most, a random other point or none for the rest
        own = np.concatenate([vis, rng.integers(0, n_points, n_extra)])
        wrong = rng.random(n) < 0.1
        own[wrong] = rng.integers(0, n_points, int(wrong.sum()))
        min_d = (max_d / sf[nlevels - 1]).astype(np.float32)
        mp_dist = np.stack([np.float32(0.8) * min_d[own], np.float32(1.2) * max_d[own], max_d[own]], 1).astype(np.float32)
        flags = (rng.random(n) < 0.8).astype(np.uint8)
        perm = rng.permutation(n)
        kfs.append(dict(kps=kps[perm], desc=desc[perm], mp_xyz=Xw[own].astype(np.float32)[perm], mp_dist=mp_dist[perm],
                        mp_desc=pdesc[own][perm], mp_flags=flags[perm], Tcw12=np.concatenate([R.ravel(), t]).astype(np.float32)))
    s12 = np.float32(1.03)
    R12 = (R1 @ R2.T).astype(np.float32); t12 = (t1 - (R1 @ R2.T) @ t2).astype(np.float32)
    sR12 = (s12 * R12).astype(np.float32)
    sR21 = ((np.float64(1.0) / np.float64(s12)) * R12.T.astype(np.float64)).astype(np.float32)
    t21 = (-(sR21.astype(np.float64) @ t12.astype(np.float64))).astype(np.float32)
    S12 = np.concatenate([sR12.ravel(), t12]).astype(np.float32); S21 = np.concatenate([sR21.ravel(), t21]).astype(np.float32)
    cam9 = np.array([fx, fy, cx, cy, mbf, 0.0, W, 0.0, H], np.float32)
    return kfs[0], kfs[1], S12, S21, cam9, sf, float(log_sf)


def synth_initialization_scene(seed: int, n: int = 1500, cluster: float = 0.25):
    """Inputs of ORBmatcher::SearchForInitialization (ORBmatcher.cc:442-587): an initial frame and a current frame a small
    image motion away; half of the keypoints on level 0; `cluster` of the F1 keypoints are near-duplicates of another one
    (close by, near-identical descriptor), so that F2 keypoints change hands during the loop.
    Returns the keyword arguments of api.search_for_initialization (without window_size / nnratio / check_orientation)."""
    rng = np.random.default_rng(seed)
    W, H = 640.0, 480.0
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                         ("octave", "<i4"), ("class_id", "<i4")])
    n_base = int(n * (1 - cluster))
    x = rng.uniform(20, W - 20, n_base); y = rng.uniform(20, H - 20, n_base)
    d = rng.integers(0, 256, (n_base, 32), dtype=np.uint8)
    octv = np.where(rng.random(n_base) < 0.55, 0, rng.integers(1, 8, n_base))
    ang = rng.uniform(0, 360, n_base)
    dup = rng.integers(0, n_base, n - n_base)
    x1 = np.concatenate([x, x[dup] + rng.normal(0, 3, len(dup))]); y1 = np.concatenate([y, y[dup] + rng.normal(0, 3, len(dup))])
    d1 = np.concatenate([d, d[dup]])
    for i in range(n_base, n):
        for b in rng.integers(0, 256, rng.integers(0, 6)):
            d1[i, b >> 3] ^= np.uint8(1 << (b & 7))
    o1 = np.concatenate([octv, octv[dup]]); a1 = np.concatenate([ang, ang[dup]])
    p1 = rng.permutation(n)
    k1 = np.zeros(n, kp_dtype); k1["x"] = x1[p1]; k1["y"] = y1[p1]; k1["octave"] = o1[p1]; k1["angle"] = a1[p1]; d1 = d1[p1]
    # F2: the base keypoints moved by a smooth flow, descriptors with noise, some dropped, some new
    keep = rng.random(n_base) < 0.85
    n_new = 300
    k2 = np.zeros(int(keep.sum()) + n_new, kp_dtype)
    m = int(keep.sum())
    k2["x"][:m] = x[keep] + 12 + 0.02 * (y[keep] - H / 2) + rng.normal(0, 1, m); k2["y"][:m] = y[keep] - 6 + rng.normal(0, 1, m)
    k2["octave"][:m] = octv[keep]
    k2["angle"][:m] = (ang[keep] + rng.normal(5, 3, m) + (rng.random(m) < 0.1) * rng.uniform(0, 360, m)) % 360
    k2["x"][m:] = rng.uniform(0, W, n_new); k2["y"][m:] = rng.uniform(0, H, n_new)
    k2["octave"][m:] = np.where(rng.random(n_new) < 0.55, 0, rng.integers(1, 8, n_new)); k2["angle"][m:] = rng.uniform(0, 360, n_new)
    d2 = np.concatenate([d[keep], rng.integers(0, 256, (n_new, 32), dtype=np.uint8)])
    for i in range(m):
        for b in rng.integers(0, 256, rng.integers(0, 50)):
            d2[i, b >> 3] ^= np.uint8(1 << (b & 7))
    p2 = rng.permutation(len(k2))
    k2, d2 = k2[p2], d2[p2]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)                      # Tracking.cc: mvbPrevMatched[i] = mInitialFrame.mvKeysUn[i].pt
    return dict(kps1=k1, desc1=d1, kps2=k2, desc2=d2, bounds4=np.array([0.0, W, 0.0, H], np.float32), prev_matched=prev)


def golden_matcher_scenes():
    """The scene set behind tests/golden/ref_matchers.npz, shared by tools/gen_golden_matchers.py and the tests so that generator
    and checks cannot drift apart."""
    s = {}
    s["local"] = synth_local_points_scene(41, n_points=600, n_extra=150)
    s["track"] = synth_tracking_scene(42, n_last=500, n_extra=150)
    s["kf"] = synth_kf_projection_scene(43, n_points=500, n_extra=200)
    mx = s["kf"]["pt_dist"][:, 2].copy(); mn = (mx / s["kf"]["scale_factors"][-1]).astype(np.float32)
    s["kf_raw"] = np.stack([mn, mx], 1).astype(np.float32)
    s["voc"] = synth_vocabulary(10, 4, 5)
    s["tri"] = synth_triangulation_scene(s["voc"], 44, n_points=600, n_extra=150)
    s["init"] = synth_initialization_scene(45, n=700)
    s["sim3"] = synth_sim3_scene(46, n_points=700, n_extra=150)
    return s
