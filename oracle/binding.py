"""ctypes binding of the CPU ORACLE (oracle/liborb_oracle.so) and of oracle/_ref/liborb_ref.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liborb_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "liborb_ref.so")
MATCHER_REF_LIB = os.path.join(HERE, "_ref", "libmatcher_ref.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference exists). Building is not using."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "orb_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "-B", "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/ORBextractor.cc"):
        srcs = [os.path.join(HERE, "ref_glue.cc"), os.path.join(HERE, "cvshim", "opencv2", "core", "core.hpp")]
        if all(os.path.exists(s) for s in srcs):
            stale = (not os.path.exists(REF_LIB)) or any(os.path.getmtime(REF_LIB) < os.path.getmtime(s) for s in srcs)
            if force or stale:
                subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/ORBmatcher.cc"):
        srcs = [os.path.join(HERE, "matcher_glue.cc")] + [os.path.join(HERE, "slamshim", f) for f in
                                                          ("Frame.h", "KeyFrame.h", "MapPoint.h", "opencv2/core/core.hpp")]
        stale = (not os.path.exists(MATCHER_REF_LIB)) or any(os.path.getmtime(MATCHER_REF_LIB) < os.path.getmtime(s) for s in srcs)
        if force or stale:
            subprocess.check_call(["make", "-C", HERE, "ref_matcher"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"):
        bow_lib = os.path.join(HERE, "_ref", "libbow_ref.so")
        srcs = [os.path.join(HERE, "bow_glue.cc"), os.path.join(HERE, "slamshim", "opencv2", "core", "core.hpp")]
        if force or (not os.path.exists(bow_lib)) or any(os.path.getmtime(bow_lib) < os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["make", "-C", HERE, "ref_bow"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.oc_round_f.restype = C.c_int; L.oc_round_f.argtypes = [C.c_float]
        L.oc_fast_atan2.restype = C.c_float; L.oc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.oc_cosf.restype = C.c_float; L.oc_cosf.argtypes = [C.c_float]
        L.oc_sinf.restype = C.c_float; L.oc_sinf.argtypes = [C.c_float]
        L.oc_resize_linear_8u.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.oc_border_reflect101.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oc_gaussian7x7_s2.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.oc_cvt_gray.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.oc_remap_linear_8u.argtypes = [u8p, C.c_int, C.c_int, C.c_int, f32p, f32p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.oc_undistort_points.argtypes = [f32p, C.c_int, f32p, f32p, C.c_int, f32p]
        L.oc_fast_score.restype = C.c_int; L.oc_fast_score.argtypes = [u8p, C.c_int]
        L.oc_search_by_bow_kf.restype = C.c_int
        L.oc_search_by_bow_kf.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + \
            [C.c_void_p] * 3 + [C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.oc_search_by_projection_frame.restype = C.c_int
        L.oc_search_by_projection_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                                    C.c_int, C.c_int, C.c_void_p]
        L.oc_vocab_create.restype = C.c_void_p
        L.oc_vocab_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, u8p, u8p, C.POINTER(C.c_double)]
        L.oc_vocab_destroy.argtypes = [C.c_void_p]
        L.oc_vocab_words.restype = C.c_int; L.oc_vocab_words.argtypes = [C.c_void_p]
        L.oc_vocab_transform.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int] + [C.c_void_p] * 9
        L.oc_bow_score_l1.restype = C.c_double
        L.oc_bow_score_l1.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oc_search_by_bow.restype = C.c_int
        L.oc_search_by_bow.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 5 + \
            [C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.oc_fast9_16.restype = C.c_int
        L.oc_fast9_16.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.oc_ic_angle.restype = C.c_float; L.oc_ic_angle.argtypes = [C.c_void_p, C.c_int, i32p]
        L.oc_orb_descriptor.argtypes = [C.c_void_p, C.c_int, C.c_float, u8p]
        L.oc_distribute_octtree.restype = C.c_int
        L.oc_distribute_octtree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.oc_descriptor_distance.restype = C.c_int; L.oc_descriptor_distance.argtypes = [u8p, u8p]
        L.oc_hamming_top2.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, i32p, i32p, C.c_int]
        L.oc_stereo_hamming.argtypes = [C.c_void_p, u8p, C.c_int, C.c_void_p, u8p, C.c_int, C.c_int, f32p,
                                        C.c_float, C.c_float, i32p, i32p]
        L.oc_create.restype = C.c_void_p
        L.oc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.oc_destroy.argtypes = [C.c_void_p]
        L.oc_levels.restype = C.c_int; L.oc_levels.argtypes = [C.c_void_p]
        L.oc_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, i32p, i32p]
        L.oc_extract.restype = C.c_int
        L.oc_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, u8p]
        L.oc_level_size.restype = C.c_int
        L.oc_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p]
        L.oc_level_ptr.restype = C.c_void_p; L.oc_level_ptr.argtypes = [C.c_void_p, C.c_int]
        L.oc_level_blur_ptr.restype = C.c_void_p; L.oc_level_blur_ptr.argtypes = [C.c_void_p, C.c_int]
        L.oc_level_candidates.restype = C.c_int
        L.oc_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oc_level_nkeypoints.restype = C.c_int; L.oc_level_nkeypoints.argtypes = [C.c_void_p, C.c_int]
        L.oc_window_top2.argtypes = [C.c_void_p, u8p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, u8p, C.c_int, i32p, i32p, i32p, i32p, i32p]
        L.oc_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u8p, C.c_int, C.c_void_p, u8p, C.c_int,
                                      C.c_float, C.c_float, f32p, f32p]
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(u8p)


def resize_linear(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().oc_resize_linear_8u(_u8(src), src.shape[1], src.shape[0], src.shape[1], _u8(dst), dw, dh, dw)
    return dst


def border101(img: np.ndarray, b: int = 19) -> np.ndarray:
    h, w = img.shape
    whole = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
    whole[b:b + h, b:b + w] = img
    lib().oc_border_reflect101(_u8(whole), w, h, w + 2 * b, b)
    return whole


def gaussian7(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().oc_gaussian7x7_s2(_u8(img), img.shape[1], img.shape[0], img.shape[1], _u8(out), img.shape[1])
    return out


def cvt_gray(img: np.ndarray, rgb: bool) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    out = np.empty((h, w), np.uint8)
    lib().oc_cvt_gray(_u8(img), w, h, w * ch, ch, int(rgb), _u8(out), w)
    return out


def remap(img: np.ndarray, map1: np.ndarray, map2: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    map1 = np.ascontiguousarray(map1, np.float32); map2 = np.ascontiguousarray(map2, np.float32)
    dh, dw = map1.shape
    out = np.empty((dh, dw), np.uint8)
    lib().oc_remap_linear_8u(_u8(img), img.shape[1], img.shape[0], img.shape[1], map1.ctypes.data_as(f32p),
                             map2.ctypes.data_as(f32p), dw, dw, dh, _u8(out), dw)
    return out


def init_undistort_rectify_map(K, D, R, P, size):
    """cv::initUndistortRectifyMap(K, D, R, P, size, CV_32FC1) -> (map1, map2); size = (width, height); P 3x3 or 3x4."""
    import ctypes as C
    K = np.ascontiguousarray(K, np.float64).reshape(3, 3); D = np.ascontiguousarray(D, np.float64).ravel()
    R = np.ascontiguousarray(R, np.float64).reshape(3, 3); Ar = np.ascontiguousarray(np.asarray(P, np.float64)[:, :3])
    w, h = size
    m1 = np.empty((h, w), np.float32); m2 = np.empty((h, w), np.float32)
    f64p = C.POINTER(C.c_double)
    L = lib()
    L.oc_init_undistort_rectify_map.argtypes = [f64p, f64p, C.c_int, f64p, f64p, C.c_int, C.c_int, f32p, f32p]
    L.oc_init_undistort_rectify_map.restype = None
    L.oc_init_undistort_rectify_map(K.ctypes.data_as(f64p), D.ctypes.data_as(f64p), len(D), R.ctypes.data_as(f64p), Ar.ctypes.data_as(f64p),
                                    w, h, m1.ctypes.data_as(f32p), m2.ctypes.data_as(f32p))
    return m1, m2


def undistort_points(xy: np.ndarray, K4, dist) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    K4 = np.ascontiguousarray(K4, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(xy)
    lib().oc_undistort_points(xy.ctypes.data_as(f32p), len(xy), K4.ctypes.data_as(f32p), dist.ctypes.data_as(f32p),
                              len(dist), out.ctypes.data_as(f32p))
    return out


def fast(roi: np.ndarray, threshold: int, nms: bool = True) -> np.ndarray:
    roi = np.ascontiguousarray(roi, np.uint8)
    cap = roi.size
    out = np.zeros(cap, KP_DTYPE)
    n = lib().oc_fast9_16(_u8(roi), roi.shape[1], roi.shape[0], roi.shape[1], threshold, int(nms), out.ctypes.data, cap)
    return out[:n]


def fast_atan2(y, x) -> np.ndarray:
    y = np.asarray(y, np.float32).ravel(); x = np.asarray(x, np.float32).ravel()
    L = lib()
    return np.array([L.oc_fast_atan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)


def distribute_octtree(cand: np.ndarray, minX, maxX, minY, maxY, N) -> np.ndarray:
    cand = np.ascontiguousarray(cand, KP_DTYPE)
    out = np.zeros(max(len(cand), 1), KP_DTYPE)
    n = lib().oc_distribute_octtree(cand.ctypes.data, len(cand), minX, maxX, minY, maxY, N, out.ctypes.data, len(out))
    if n < 0:
        raise RuntimeError(f"oc_distribute_octtree status {n}")
    return out[:n]


def hamming_top2(q: np.ndarray, t: np.ndarray, nthreads: int = 1):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    nq, nt = len(q), len(t)
    idx = np.empty(nq, np.int32); d1 = np.empty(nq, np.int32); d2 = np.empty(nq, np.int32)
    lib().oc_hamming_top2(_u8(q), nq, _u8(t), nt, idx.ctypes.data_as(i32p), d1.ctypes.data_as(i32p),
                          d2.ctypes.data_as(i32p), nthreads)
    return idx, d1, d2


def descriptor_distance(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().oc_descriptor_distance(_u8(a), _u8(b))


def stereo_hamming(kl, dl, kr, dr, rows, scale_factors, minD, maxD):
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    bi = np.empty(len(kl), np.int32); bd = np.empty(len(kl), np.int32)
    lib().oc_stereo_hamming(kl.ctypes.data, _u8(dl), len(kl), kr.ctypes.data, _u8(dr), len(kr), rows,
                            sf.ctypes.data_as(f32p), minD, maxD, bi.ctypes.data_as(i32p), bd.ctypes.data_as(i32p))
    return bi, bd


WQ_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"), ("xr", "<f4")])


def window_top2(kps, desc, occupied, u_right, minX, minY, invW, invH, queries, qdesc):
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    queries = np.ascontiguousarray(queries, WQ_DTYPE); qdesc = np.ascontiguousarray(qdesc, np.uint8)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    nq = len(queries)
    out = [np.empty(nq, np.int32) for _ in range(5)]
    lib().oc_window_top2(kps.ctypes.data, _u8(desc), len(kps), None if occ is None else occ.ctypes.data,
                         None if ur is None else ur.ctypes.data, minX, minY, invW, invH, queries.ctypes.data, _u8(qdesc), nq,
                         *[o.ctypes.data_as(i32p) for o in out])
    return out


def stereo_match(left: "Extractor", right: "Extractor", kl, dl, kr, dr, mbf, fx):
    """Full Frame::ComputeStereoMatches on two oracle extractors that just processed the left / right image."""
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32)
    lib().oc_stereo_match(left.h, right.h, kl.ctypes.data, _u8(dl), len(kl), kr.ctypes.data, _u8(dr), len(kr),
                          mbf, fx, ur.ctypes.data_as(f32p), dp.ctypes.data_as(f32p))
    return ur, dp


class Extractor:
    """Oracle mirror of ORB_SLAM2::ORBextractor (ORBextractor.h:51-145)."""

    def __init__(self, nfeatures, scale, nlevels, ini_th, min_th):
        self.L = lib()
        self.h = self.L.oc_create(nfeatures, scale, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("oc_create failed")
        self.nlevels = nlevels
        self.nfeatures = nfeatures

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oc_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sf = np.zeros(n, np.float32); inv = np.zeros(n, np.float32)
        s2 = np.zeros(n, np.float32); is2 = np.zeros(n, np.float32)
        fpl = np.zeros(n, np.int32); um = np.zeros(16, np.int32)
        self.L.oc_tables(self.h, sf.ctypes.data_as(f32p), inv.ctypes.data_as(f32p), s2.ctypes.data_as(f32p),
                         is2.ctypes.data_as(f32p), fpl.ctypes.data_as(i32p), um.ctypes.data_as(i32p))
        return dict(scale_factors=sf, inv_scale_factors=inv, sigma2=s2, inv_sigma2=is2,
                    features_per_level=fpl, umax=um)

    def extract(self, img: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures * 2 + 64 * self.nlevels + 4096
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.oc_extract(self.h, _u8(img), img.shape[1], img.shape[0], img.shape[1], kps.ctypes.data, cap, _u8(desc))
        if n < 0:
            raise RuntimeError(f"oc_extract status {n}")
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l: int, blurred: bool = False) -> np.ndarray:
        w = C.c_int32(); h = C.c_int32(); s = C.c_int32()
        if self.L.oc_level_size(self.h, l, C.byref(w), C.byref(h), C.byref(s)) != 0:
            raise RuntimeError("no level")
        if blurred:
            p = self.L.oc_level_blur_ptr(self.h, l)
            if not p:
                return None
            buf = (C.c_uint8 * (w.value * h.value)).from_address(p)
            return np.frombuffer(buf, np.uint8).reshape(h.value, w.value).copy()
        p = self.L.oc_level_ptr(self.h, l)
        # view including the apron
        base = p - 19 * s.value - 19
        buf = (C.c_uint8 * (s.value * (h.value + 38))).from_address(base)
        whole = np.frombuffer(buf, np.uint8).reshape(h.value + 38, s.value).copy()
        return whole

    def candidates(self, l: int) -> np.ndarray:
        n = self.L.oc_level_candidates(self.h, l, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.L.oc_level_candidates(self.h, l, out.ctypes.data, n)
        return out[:n]

    def level_counts(self):
        return [self.L.oc_level_nkeypoints(self.h, l) for l in range(self.nlevels)]


# ---------------------------------------------------------------- oracle/_ref (verbatim reference + cvshim)
_ref = None


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def ref_lib():
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_LIB)
        R.orbref_create.restype = C.c_void_p
        R.orbref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        R.orbref_destroy.argtypes = [C.c_void_p]
        R.orbref_set_deterministic.argtypes = [C.c_int]
        R.orbref_extract.restype = C.c_int
        R.orbref_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, u8p]
        R.orbref_level.restype = C.c_int
        R.orbref_level.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p, C.POINTER(C.c_void_p)]
        R.orbref_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p]
        R.orbref_bench.restype = C.c_double
        R.orbref_bench.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        _ref = R
    return _ref


class RefExtractor:
    """The verbatim reference ORBextractor (compiled against oracle/cvshim)."""

    def __init__(self, nfeatures, scale, nlevels, ini_th, min_th, deterministic=True):
        self.R = ref_lib()
        self.R.orbref_set_deterministic(int(deterministic))
        self.h = self.R.orbref_create(nfeatures, scale, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.nfeatures = nfeatures

    def __del__(self):
        if getattr(self, "h", None):
            self.R.orbref_destroy(self.h)
            self.h = None

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures * 2 + 64 * self.nlevels + 4096
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.R.orbref_extract(self.h, _u8(img), img.shape[1], img.shape[0], img.shape[1], kps.ctypes.data, cap, _u8(desc))
        if n < 0:
            raise RuntimeError(f"orbref_extract status {n}")
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        w = C.c_int32(); h = C.c_int32(); s = C.c_int32(); p = C.c_void_p()
        self.R.orbref_level(self.h, l, C.byref(w), C.byref(h), C.byref(s), C.byref(p))
        buf = (C.c_uint8 * (s.value * (h.value + 38))).from_address(p.value - 19 * s.value - 19)
        return np.frombuffer(buf, np.uint8).reshape(h.value + 38, s.value).copy()

    def tables(self):
        n = self.nlevels
        a = [np.zeros(n, np.float32) for _ in range(4)]
        self.R.orbref_tables(self.h, *[x.ctypes.data_as(f32p) for x in a])
        return dict(scale_factors=a[0], inv_scale_factors=a[1], sigma2=a[2], inv_sigma2=a[3])


class Vocabulary:
    """DBoW2 vocabulary (restated; see orb_oracle.h). parent / is_leaf / desc / weight: one entry per non-root node in
    text-file order."""

    def __init__(self, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0):
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        self._h = lib().oc_vocab_create(k, L, scoring, weighting, len(parent), parent.ctypes.data_as(i32p), _u8(is_leaf), _u8(desc),
                                        weight.ctypes.data_as(C.POINTER(C.c_double)))
        assert self._h, "invalid vocabulary"
        self.nwords = lib().oc_vocab_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oc_vocab_destroy(self._h); self._h = None

    def transform(self, desc, levelsup=4):
        """-> dict(word, node, bow_id, bow_val, fv_node, fv_off, fv_feat)"""
        desc = np.ascontiguousarray(desc, np.uint8); n = len(desc)
        word = np.zeros(n, np.int32); node = np.zeros(n, np.int32)
        bow_id = np.zeros(n, np.int32); bow_val = np.zeros(n, np.float64); fv_node = np.zeros(n, np.int32)
        fv_off = np.zeros(n + 1, np.int32); fv_feat = np.zeros(n, np.int32); nb = C.c_int32(0); nf = C.c_int32(0)
        lib().oc_vocab_transform(self._h, _u8(desc), n, levelsup, word.ctypes.data, node.ctypes.data, bow_id.ctypes.data,
                                 bow_val.ctypes.data, C.addressof(nb), fv_node.ctypes.data, fv_off.ctypes.data, fv_feat.ctypes.data,
                                 C.addressof(nf))
        return dict(word=word, node=node, bow_id=bow_id[:nb.value], bow_val=bow_val[:nb.value], fv_node=fv_node[:nf.value],
                    fv_off=fv_off[:nf.value + 1], fv_feat=fv_feat[:fv_off[nf.value]])


def bow_score_l1(id1, v1, id2, v2) -> float:
    id1 = np.ascontiguousarray(id1, np.int32); id2 = np.ascontiguousarray(id2, np.int32)
    v1 = np.ascontiguousarray(v1, np.float64); v2 = np.ascontiguousarray(v2, np.float64)
    return lib().oc_bow_score_l1(id1.ctypes.data, v1.ctypes.data, len(id1), id2.ctypes.data, v2.ctypes.data, len(id2))


def search_by_bow(kf, f, kf_desc, kf_angle, kf_valid, f_desc, f_angle, nnratio=0.7, check_orientation=True):
    """kf / f: dicts from Vocabulary.transform. -> (nmatches, match_f)"""
    kf_desc = np.ascontiguousarray(kf_desc, np.uint8); f_desc = np.ascontiguousarray(f_desc, np.uint8)
    kf_angle = np.ascontiguousarray(kf_angle, np.float32); f_angle = np.ascontiguousarray(f_angle, np.float32)
    kf_valid = np.ascontiguousarray(kf_valid, np.uint8)
    a = [np.ascontiguousarray(kf[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    b = [np.ascontiguousarray(f[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    match = np.zeros(len(f_desc), np.int32)
    n = lib().oc_search_by_bow(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[0]), b[0].ctypes.data, b[1].ctypes.data,
                               b[2].ctypes.data, len(b[0]), kf_desc.ctypes.data, kf_angle.ctypes.data, kf_valid.ctypes.data,
                               f_desc.ctypes.data, f_angle.ctypes.data, len(f_desc), nnratio, int(check_orientation), match.ctypes.data)
    return n, match


def search_by_projection_frame(cur_kps, cur_desc, cur_u_right, cur_occupied, Tcw12, cam9, scale_factors, last_kps, last_xyz,
                               last_desc, last_flags, th, mode, check_orientation=True):
    """-> (nmatches, match_cur)"""
    cur_kps = np.ascontiguousarray(cur_kps, KP_DTYPE); last_kps = np.ascontiguousarray(last_kps, KP_DTYPE)
    cur_desc = np.ascontiguousarray(cur_desc, np.uint8); last_desc = np.ascontiguousarray(last_desc, np.uint8)
    ur = None if cur_u_right is None else np.ascontiguousarray(cur_u_right, np.float32)
    occ = None if cur_occupied is None else np.ascontiguousarray(cur_occupied, np.uint8)
    T = np.ascontiguousarray(Tcw12, np.float32); cam = np.ascontiguousarray(cam9, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    xyz = np.ascontiguousarray(last_xyz, np.float32); fl = np.ascontiguousarray(last_flags, np.uint8)
    match = np.zeros(max(len(cur_kps), 1), np.int32)
    n = lib().oc_search_by_projection_frame(cur_kps.ctypes.data, cur_desc.ctypes.data, len(cur_kps),
                                            None if ur is None else ur.ctypes.data, None if occ is None else occ.ctypes.data,
                                            T.ctypes.data, cam.ctypes.data, sf.ctypes.data, last_kps.ctypes.data, xyz.ctypes.data,
                                            last_desc.ctypes.data, fl.ctypes.data, len(last_kps), th, mode, int(check_orientation),
                                            match.ctypes.data)
    return n, match[:len(cur_kps)]


def search_by_bow_kf(t1, t2, desc1, angle1, valid1, desc2, angle2, valid2, nnratio=0.75, check_orientation=True):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12); t1 / t2: dicts from Vocabulary.transform. -> (nmatches, match12)"""
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    angle1 = np.ascontiguousarray(angle1, np.float32); angle2 = np.ascontiguousarray(angle2, np.float32)
    valid1 = np.ascontiguousarray(valid1, np.uint8); valid2 = np.ascontiguousarray(valid2, np.uint8)
    a = [np.ascontiguousarray(t1[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    b = [np.ascontiguousarray(t2[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    match = np.zeros(max(len(desc1), 1), np.int32)
    n = lib().oc_search_by_bow_kf(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[0]), b[0].ctypes.data, b[1].ctypes.data,
                                  b[2].ctypes.data, len(b[0]), desc1.ctypes.data, angle1.ctypes.data, valid1.ctypes.data, len(desc1),
                                  desc2.ctypes.data, angle2.ctypes.data, valid2.ctypes.data, len(desc2), nnratio,
                                  int(check_orientation), match.ctypes.data)
    return n, match[:len(desc1)]


TRACKQ_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4")])


def _p(a):
    return None if a is None else a.ctypes.data


def search_local_points(kps, desc, u_right, occupied, bounds4, scale_factors, queries, query_desc, query_flags, th, nnratio=0.8):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:46-142) -> (nmatches, match)"""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    b4 = np.ascontiguousarray(bounds4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    q = np.ascontiguousarray(queries, TRACKQ_DTYPE); qd = np.ascontiguousarray(query_desc, np.uint8)
    qf = np.ascontiguousarray(query_flags, np.uint8)
    match = np.zeros(max(len(kps), 1), np.int32)
    f = lib().oc_search_local_points
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
    n = f(kps.ctypes.data, desc.ctypes.data, len(kps), _p(ur), _p(occ), b4.ctypes.data, sf.ctypes.data, len(sf), q.ctypes.data,
          qd.ctypes.data, qf.ctypes.data, len(q), th, nnratio, match.ctypes.data)
    return n, match[:len(kps)]


def predict_scale(max_distance, current_dist, log_scale_factor, nlevels):
    f = lib().oc_predict_scale
    f.restype = C.c_int; f.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    return f(max_distance, current_dist, log_scale_factor, nlevels)


def fuse_search(kps, desc, u_right, Tcw12, Ow3, cam9, scale_factors, inv_level_sigma2, log_scale_factor, pt_xyz, pt_normal,
                pt_dist, pt_desc, pt_flags, th, mode=0):
    """Search half of ORBmatcher::Fuse (ORBmatcher.cc:918-1092 / 1094-1236) -> (nFused, best_idx, best_dist)"""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    T = np.ascontiguousarray(Tcw12, np.float32); Ow = np.ascontiguousarray(Ow3, np.float32)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = np.ascontiguousarray(pt_normal, np.float32)
    dst = np.ascontiguousarray(pt_dist, np.float32); pd = np.ascontiguousarray(pt_desc, np.uint8)
    pf = np.ascontiguousarray(pt_flags, np.uint8)
    npts = len(pf)
    bi = np.zeros(max(npts, 1), np.int32); bd = np.zeros(max(npts, 1), np.int32)
    f = lib().oc_fuse_search
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_float] + [C.c_void_p] * 5 + \
        [C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    n = f(kps.ctypes.data, desc.ctypes.data, len(kps), _p(ur), T.ctypes.data, Ow.ctypes.data, cam.ctypes.data, sf.ctypes.data,
          s2.ctypes.data, len(sf), float(log_scale_factor), xyz.ctypes.data, nrm.ctypes.data, dst.ctypes.data, pd.ctypes.data,
          pf.ctypes.data, npts, th, mode, bi.ctypes.data, bd.ctypes.data)
    return n, bi[:npts], bd[:npts]


def search_for_triangulation(t1, t2, kps1, desc1, has_mp1, u_right1, kps2, desc2, has_mp2, u_right2, geom28, scale_factors,
                             level_sigma2, only_stereo=False, check_orientation=True):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:738-916); t1 / t2: dicts from Vocabulary.transform. geom28 = F12 (9),
    Cw1 (3), R2w (9), t2w (3), K2 (4). -> (nmatches, match12)"""
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    m1 = None if has_mp1 is None else np.ascontiguousarray(has_mp1, np.uint8)
    m2 = None if has_mp2 is None else np.ascontiguousarray(has_mp2, np.uint8)
    r1 = None if u_right1 is None else np.ascontiguousarray(u_right1, np.float32)
    r2 = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
    g = np.ascontiguousarray(geom28, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    s2 = np.ascontiguousarray(level_sigma2, np.float32)
    a = [np.ascontiguousarray(t1[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    b = [np.ascontiguousarray(t2[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    match = np.zeros(max(len(kps1), 1), np.int32)
    f = lib().oc_search_for_triangulation
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + \
        [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]
    n = f(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[0]), b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data, len(b[0]),
          kps1.ctypes.data, desc1.ctypes.data, _p(m1), _p(r1), len(kps1), kps2.ctypes.data, desc2.ctypes.data, _p(m2), _p(r2), len(kps2),
          g[0:9].ctypes.data, g[9:12].ctypes.data, g[12:24].ctypes.data, g[24:28].ctypes.data, sf.ctypes.data, s2.ctypes.data,
          int(only_stereo), int(check_orientation), match.ctypes.data)
    return n, match[:len(kps1)]


def search_by_projection_kf(kps, desc, occupied, Tcw12, Ow3, cam9, scale_factors, log_scale_factor, pt_xyz, pt_normal, pt_dist,
                            pt_desc, pt_flags, pt_angle, th, max_dist, mode, check_orientation=True):
    """ORBmatcher.cc:1648-1795 (mode 0) / :327-440 (mode 1) -> (nmatches, match)"""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    T = np.ascontiguousarray(Tcw12, np.float32); Ow = np.ascontiguousarray(Ow3, np.float32)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = None if pt_normal is None else np.ascontiguousarray(pt_normal, np.float32)
    dst = np.ascontiguousarray(pt_dist, np.float32); pd = np.ascontiguousarray(pt_desc, np.uint8)
    pf = np.ascontiguousarray(pt_flags, np.uint8); pa = None if pt_angle is None else np.ascontiguousarray(pt_angle, np.float32)
    match = np.zeros(max(len(kps), 1), np.int32)
    f = lib().oc_search_by_projection_seq
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_float] + [C.c_void_p] * 6 + \
        [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]
    n = f(kps.ctypes.data, desc.ctypes.data, len(kps), _p(occ), T.ctypes.data, Ow.ctypes.data, cam.ctypes.data, sf.ctypes.data, len(sf),
          float(log_scale_factor), xyz.ctypes.data, _p(nrm), dst.ctypes.data, pd.ctypes.data, pf.ctypes.data, _p(pa), len(pf), th,
          max_dist, mode, int(check_orientation), match.ctypes.data)
    return n, match[:len(kps)]


def search_by_sim3(kf1, kf2, S12, S21, cam9, scale_factors, log_scale_factor, th):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1238-1487) -> (nFound, match12)"""
    def arrs(k):
        return [np.ascontiguousarray(k["kps"], KP_DTYPE), np.ascontiguousarray(k["desc"], np.uint8), np.ascontiguousarray(k["mp_xyz"], np.float32),
                np.ascontiguousarray(k["mp_dist"], np.float32), np.ascontiguousarray(k["mp_desc"], np.uint8), np.ascontiguousarray(k["mp_flags"], np.uint8)]
    a, b = arrs(kf1), arrs(kf2)
    T1 = np.ascontiguousarray(kf1["Tcw12"], np.float32); T2 = np.ascontiguousarray(kf2["Tcw12"], np.float32)
    s12 = np.ascontiguousarray(S12, np.float32); s21 = np.ascontiguousarray(S21, np.float32)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    match = np.zeros(max(len(a[0]), 1), np.int32)
    f = lib().oc_search_by_sim3
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + \
        [C.c_void_p] * 6 + [C.c_int, C.c_float, C.c_float, C.c_void_p]
    n = f(a[0].ctypes.data, a[1].ctypes.data, len(a[0]), a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data, a[5].ctypes.data,
          b[0].ctypes.data, b[1].ctypes.data, len(b[0]), b[2].ctypes.data, b[3].ctypes.data, b[4].ctypes.data, b[5].ctypes.data,
          T1.ctypes.data, T2.ctypes.data, s12.ctypes.data, s21.ctypes.data, cam.ctypes.data, sf.ctypes.data, len(sf),
          float(log_scale_factor), th, match.ctypes.data)
    return n, match[:len(a[0])]


def search_for_initialization(kps1, desc1, kps2, desc2, bounds4, prev_matched, window_size=100, nnratio=0.9, check_orientation=True):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:442-587) -> (nmatches, vnMatches12, updated vbPrevMatched)"""
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    b4 = np.ascontiguousarray(bounds4, np.float32)
    prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    match = np.zeros(max(len(kps1), 1), np.int32)
    f = lib().oc_search_for_initialization
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int,
                  C.c_void_p]
    n = f(kps1.ctypes.data, desc1.ctypes.data, len(kps1), kps2.ctypes.data, desc2.ctypes.data, len(kps2), b4.ctypes.data,
          prev.ctypes.data, window_size, nnratio, int(check_orientation), match.ctypes.data)
    return n, match[:len(kps1)], prev


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref/libmatcher_ref.so: the UNMODIFIED reference ORBmatcher.cc behind oracle/matcher_glue.cc (see its header).
_mref = None


def matcher_ref():
    """ctypes handle of the verbatim-reference matcher build, or None when it has not been built (no reference tree)."""
    global _mref
    if _mref is None:
        build()
        if not os.path.exists(MATCHER_REF_LIB):
            return None
        _mref = C.CDLL(MATCHER_REF_LIB)
    return _mref


def _dist4(pt_dist_raw):
    """(mfMinDistance, mfMaxDistance) per point -> rows (0.8f*min, 1.2f*max, max, min) as MapPoint.cc:395-405 return them"""
    raw = np.ascontiguousarray(pt_dist_raw, np.float32).reshape(-1, 2)
    return np.stack([np.float32(0.8) * raw[:, 0], np.float32(1.2) * raw[:, 1], raw[:, 1], raw[:, 0]], 1).astype(np.float32)


def ref_search_local_points(kps, desc, u_right, occupied, bounds4, scale_factors, queries, query_desc, query_flags, th, nnratio=0.8):
    R = matcher_ref()
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    b4 = np.ascontiguousarray(bounds4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    q = np.ascontiguousarray(queries, TRACKQ_DTYPE); qd = np.ascontiguousarray(query_desc, np.uint8); qf = np.ascontiguousarray(query_flags, np.uint8)
    q4 = np.ascontiguousarray(np.stack([q["proj_x"], q["proj_y"], q["proj_xr"], q["view_cos"]], 1), np.float32)
    ql = np.ascontiguousarray(q["level"], np.int32)
    match = np.zeros(max(len(kps), 1), np.int32)
    f = R.mref_search_local_points
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_float, C.c_void_p]
    n = f(kps.ctypes.data, desc.ctypes.data, len(kps), _p(ur), _p(occ), b4.ctypes.data, sf.ctypes.data, len(sf), q4.ctypes.data, ql.ctypes.data,
          qd.ctypes.data, qf.ctypes.data, len(q), th, nnratio, match.ctypes.data)
    return n, match[:len(kps)]


def ref_search_by_projection_frame(cur_kps, cur_desc, cur_u_right, cur_occupied, Tcw12, cam9, scale_factors, last_kps, last_xyz,
                                   last_desc, last_flags, th, mono, tlw_z, check_orientation=True):
    """-> (nmatches, match_cur, mode the reference derived from tlc)"""
    R = matcher_ref()
    cur_kps = np.ascontiguousarray(cur_kps, KP_DTYPE); last_kps = np.ascontiguousarray(last_kps, KP_DTYPE)
    cur_desc = np.ascontiguousarray(cur_desc, np.uint8); last_desc = np.ascontiguousarray(last_desc, np.uint8)
    ur = None if cur_u_right is None else np.ascontiguousarray(cur_u_right, np.float32)
    occ = None if cur_occupied is None else np.ascontiguousarray(cur_occupied, np.uint8)
    T = np.ascontiguousarray(Tcw12, np.float32); cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    xyz = np.ascontiguousarray(last_xyz, np.float32); fl = np.ascontiguousarray(last_flags, np.uint8)
    match = np.zeros(max(len(cur_kps), 1), np.int32); mode = C.c_int32(0)
    f = R.mref_search_by_projection_frame
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_float, C.c_int, C.c_float,
                                                                                                    C.c_int, C.c_void_p, C.c_void_p]
    n = f(cur_kps.ctypes.data, cur_desc.ctypes.data, len(cur_kps), _p(ur), _p(occ), T.ctypes.data, cam.ctypes.data, sf.ctypes.data, len(sf),
          last_kps.ctypes.data, xyz.ctypes.data, last_desc.ctypes.data, fl.ctypes.data, len(last_kps), th, int(mono), tlw_z,
          int(check_orientation), match.ctypes.data, C.addressof(mode))
    return n, match[:len(cur_kps)], mode.value


def ref_fuse(kps, desc, u_right, T12, Ow3, cam9, scale_factors, inv_level_sigma2, log_scale_factor, pt_xyz, pt_normal, pt_dist_raw,
             pt_desc, pt_flags, th, mode):
    """T12: the keyframe pose (mode 0) or the 3x4 top of Scw (mode 1). pt_dist_raw: (mfMinDistance, mfMaxDistance) per point.
    -> (nFused, best_idx, Tcw12 and Ow3 as the reference derived them, pt_dist rows for the restatement)"""
    R = matcher_ref()
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    T = np.ascontiguousarray(T12, np.float32); Ow = np.zeros(3, np.float32) if Ow3 is None else np.ascontiguousarray(Ow3, np.float32)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = np.ascontiguousarray(pt_normal, np.float32)
    d4 = _dist4(pt_dist_raw); pd = np.ascontiguousarray(pt_desc, np.uint8); pf = np.ascontiguousarray(pt_flags, np.uint8)
    bi = np.zeros(max(len(pf), 1), np.int32); To = np.zeros(12, np.float32); Oo = np.zeros(3, np.float32)
    f = R.mref_fuse
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_float] + [C.c_void_p] * 5 + [C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 3
    n = f(kps.ctypes.data, desc.ctypes.data, len(kps), _p(ur), T.ctypes.data, Ow.ctypes.data, cam.ctypes.data, sf.ctypes.data, s2.ctypes.data,
          len(sf), float(log_scale_factor), xyz.ctypes.data, nrm.ctypes.data, d4.ctypes.data, pd.ctypes.data, pf.ctypes.data, len(pf), th, mode,
          bi.ctypes.data, To.ctypes.data, Oo.ctypes.data)
    return n, bi[:len(pf)], To, Oo, np.ascontiguousarray(d4[:, :3])


def ref_search_by_projection_kf(kps, desc, occupied, T12, cam9, scale_factors, log_scale_factor, pt_xyz, pt_normal, pt_dist_raw, pt_desc,
                                pt_flags, pt_angle, th, max_dist, mode, check_orientation=True):
    """-> (nmatches, match, Tcw12, Ow3 as derived by the reference, pt_dist rows for the restatement)"""
    R = matcher_ref()
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    T = np.ascontiguousarray(T12, np.float32); cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = None if pt_normal is None else np.ascontiguousarray(pt_normal, np.float32)
    d4 = _dist4(pt_dist_raw); pd = np.ascontiguousarray(pt_desc, np.uint8); pf = np.ascontiguousarray(pt_flags, np.uint8)
    pa = None if pt_angle is None else np.ascontiguousarray(pt_angle, np.float32)
    match = np.zeros(max(len(kps), 1), np.int32); To = np.zeros(12, np.float32); Oo = np.zeros(3, np.float32)
    f = R.mref_search_by_projection_seq
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_float] + [C.c_void_p] * 6 + \
        [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3
    n = f(kps.ctypes.data, desc.ctypes.data, len(kps), _p(occ), T.ctypes.data, cam.ctypes.data, sf.ctypes.data, len(sf), float(log_scale_factor),
          xyz.ctypes.data, _p(nrm), d4.ctypes.data, pd.ctypes.data, pf.ctypes.data, _p(pa), len(pf), th, max_dist, mode, int(check_orientation),
          match.ctypes.data, To.ctypes.data, Oo.ctypes.data)
    return n, match[:len(kps)], To, Oo, np.ascontiguousarray(d4[:, :3])


def ref_search_by_sim3(kf1, kf2, s12, R12, t12, cam9, scale_factors, log_scale_factor, th):
    """kf dicts as for search_by_sim3 but with mp_dist_raw = (mfMinDistance, mfMaxDistance). -> (nFound, match12, S12, S21, dist rows 1, 2)"""
    R = matcher_ref()

    def arrs(k):
        return [np.ascontiguousarray(k["kps"], KP_DTYPE), np.ascontiguousarray(k["desc"], np.uint8), np.ascontiguousarray(k["mp_xyz"], np.float32),
                _dist4(k["mp_dist_raw"]), np.ascontiguousarray(k["mp_desc"], np.uint8), np.ascontiguousarray(k["mp_flags"], np.uint8)]
    a, b = arrs(kf1), arrs(kf2)
    T1 = np.ascontiguousarray(kf1["Tcw12"], np.float32); T2 = np.ascontiguousarray(kf2["Tcw12"], np.float32)
    R12 = np.ascontiguousarray(R12, np.float32); t12 = np.ascontiguousarray(t12, np.float32)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    match = np.zeros(max(len(a[0]), 1), np.int32); S12 = np.zeros(12, np.float32); S21 = np.zeros(12, np.float32)
    f = R.mref_search_by_sim3
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + \
        [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float] + [C.c_void_p] * 3
    n = f(a[0].ctypes.data, a[1].ctypes.data, len(a[0]), a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data, a[5].ctypes.data,
          b[0].ctypes.data, b[1].ctypes.data, len(b[0]), b[2].ctypes.data, b[3].ctypes.data, b[4].ctypes.data, b[5].ctypes.data,
          T1.ctypes.data, T2.ctypes.data, float(s12), R12.ctypes.data, t12.ctypes.data, cam.ctypes.data, sf.ctypes.data, len(sf),
          float(log_scale_factor), th, match.ctypes.data, S12.ctypes.data, S21.ctypes.data)
    return n, match[:len(a[0])], S12, S21, np.ascontiguousarray(a[3][:, :3]), np.ascontiguousarray(b[3][:, :3])


def ref_search_for_triangulation(t1, t2, kps1, desc1, has_mp1, u_right1, kps2, desc2, has_mp2, u_right2, geom28, scale_factors,
                                 level_sigma2, only_stereo=False, check_orientation=True):
    R = matcher_ref()
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    m1 = None if has_mp1 is None else np.ascontiguousarray(has_mp1, np.uint8); m2 = None if has_mp2 is None else np.ascontiguousarray(has_mp2, np.uint8)
    r1 = None if u_right1 is None else np.ascontiguousarray(u_right1, np.float32); r2 = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
    g = np.ascontiguousarray(geom28, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32); s2 = np.ascontiguousarray(level_sigma2, np.float32)
    a = [np.ascontiguousarray(t1[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    b = [np.ascontiguousarray(t2[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    match = np.zeros(max(len(kps1), 1), np.int32)
    f = R.mref_search_for_triangulation
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + \
        [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_int, C.c_void_p]
    n = f(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[0]), b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data, len(b[0]),
          kps1.ctypes.data, desc1.ctypes.data, _p(m1), _p(r1), len(kps1), kps2.ctypes.data, desc2.ctypes.data, _p(m2), _p(r2), len(kps2),
          g.ctypes.data, sf.ctypes.data, s2.ctypes.data, len(sf), int(only_stereo), int(check_orientation), match.ctypes.data)
    return n, match[:len(kps1)]


def ref_search_for_initialization(kps1, desc1, kps2, desc2, bounds4, prev_matched, window_size=100, nnratio=0.9, check_orientation=True):
    R = matcher_ref()
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    b4 = np.ascontiguousarray(bounds4, np.float32)
    prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    match = np.zeros(max(len(kps1), 1), np.int32)
    f = R.mref_search_for_initialization
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = f(kps1.ctypes.data, desc1.ctypes.data, len(kps1), kps2.ctypes.data, desc2.ctypes.data, len(kps2), b4.ctypes.data, prev.ctypes.data,
          window_size, nnratio, int(check_orientation), match.ctypes.data)
    return n, match[:len(kps1)], prev


def ref_search_by_bow(t1, t2, kps1, desc1, valid1, kps2, desc2, valid2, nnratio, check_orientation, kf_mode):
    """kf_mode 0: SearchByBoW(pKF, F, ..) -> match indexed by F's features; 1: SearchByBoW(pKF1, pKF2, ..) -> indexed by pKF1's"""
    R = matcher_ref()
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    v1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8); v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
    a = [np.ascontiguousarray(t1[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    b = [np.ascontiguousarray(t2[k], np.int32) for k in ("fv_node", "fv_off", "fv_feat")]
    nout = len(kps1) if kf_mode else len(kps2)
    match = np.zeros(max(nout, 1), np.int32)
    f = R.mref_search_by_bow
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + \
        [C.c_float, C.c_int, C.c_int, C.c_void_p]
    n = f(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, len(a[0]), b[0].ctypes.data, b[1].ctypes.data, b[2].ctypes.data, len(b[0]),
          kps1.ctypes.data, desc1.ctypes.data, _p(v1), len(kps1), kps2.ctypes.data, desc2.ctypes.data, _p(v2), len(kps2), nnratio,
          int(check_orientation), kf_mode, match.ctypes.data)
    return n, match[:nout]


def ref_stereo_match(left: "Extractor", right: "Extractor", kl, dl, kr, dr, mbf, fx):
    """The verbatim Frame::ComputeStereoMatches (Frame.cc:547-788) on the pyramids of two ORACLE extractors -> (mvuRight, mvDepth)"""
    R = matcher_ref()
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    n = left.nlevels
    lv = [(left.level(l), right.level(l)) for l in range(n)]                 # copies incl. the 19-px apron
    whs = np.zeros((n, 3), np.int32)
    pl = (C.c_void_p * n)(); pr = (C.c_void_p * n)()
    for l, (a, b) in enumerate(lv):
        h, s = a.shape[0] - 38, a.shape[1]
        w = C.c_int32(); hh = C.c_int32(); ss = C.c_int32()
        left.L.oc_level_size(left.h, l, C.byref(w), C.byref(hh), C.byref(ss))
        whs[l] = (w.value, h, s)
        pl[l] = a.ctypes.data + 19 * s + 19; pr[l] = b.ctypes.data + 19 * s + 19
    t = left.tables()
    sf = np.ascontiguousarray(t["scale_factors"], np.float32); isf = np.ascontiguousarray(t["inv_scale_factors"], np.float32)
    ur = np.empty(max(len(kl), 1), np.float32); dp = np.empty(max(len(kl), 1), np.float32)
    f = R.mref_stereo_match
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                  C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    f(kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data, dr.ctypes.data, len(kr), pl, pr, whs.ctypes.data, n, sf.ctypes.data,
      isf.ctypes.data, mbf, fx, ur.ctypes.data, dp.ctypes.data)
    return ur[:len(kl)], dp[:len(kl)]


def is_in_frustum(Tcw12, Ow3, cam9, nlevels, log_scale_factor, pt_xyz, pt_normal, pt_dist, viewing_cos_limit=0.5):
    """Frame::isInFrustum (Frame.cc:315-378) -> (queries, in_view)"""
    T = np.ascontiguousarray(Tcw12, np.float32); Ow = np.ascontiguousarray(Ow3, np.float32); cam = np.ascontiguousarray(cam9, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = np.ascontiguousarray(pt_normal, np.float32); dst = np.ascontiguousarray(pt_dist, np.float32)
    n = len(xyz)
    q = np.zeros(max(n, 1), TRACKQ_DTYPE); v = np.zeros(max(n, 1), np.uint8)
    f = lib().oc_is_in_frustum
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    f(T.ctypes.data, Ow.ctypes.data, cam.ctypes.data, nlevels, float(log_scale_factor), xyz.ctypes.data, nrm.ctypes.data, dst.ctypes.data, n,
      viewing_cos_limit, q.ctypes.data, v.ctypes.data)
    return q[:n], v[:n]


def ref_is_in_frustum(Tcw12, Ow3, cam9, nlevels, log_scale_factor, pt_xyz, pt_normal, pt_dist_raw, viewing_cos_limit=0.5):
    """the verbatim Frame::isInFrustum -> (queries, in_view, pt_dist rows for the restatement)"""
    R = matcher_ref()
    T = np.ascontiguousarray(Tcw12, np.float32); Ow = np.ascontiguousarray(Ow3, np.float32); cam = np.ascontiguousarray(cam9, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = np.ascontiguousarray(pt_normal, np.float32); d4 = _dist4(pt_dist_raw)
    n = len(xyz)
    q4 = np.zeros((max(n, 1), 4), np.float32); lv = np.zeros(max(n, 1), np.int32); v = np.zeros(max(n, 1), np.uint8)
    f = R.mref_is_in_frustum
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p,
                  C.c_void_p, C.c_void_p]
    f(T.ctypes.data, Ow.ctypes.data, cam.ctypes.data, nlevels, float(log_scale_factor), xyz.ctypes.data, nrm.ctypes.data, d4.ctypes.data, n,
      viewing_cos_limit, q4.ctypes.data, lv.ctypes.data, v.ctypes.data)
    q = np.zeros(n, TRACKQ_DTYPE)
    q["proj_x"] = q4[:n, 0]; q["proj_y"] = q4[:n, 1]; q["proj_xr"] = q4[:n, 2]; q["view_cos"] = q4[:n, 3]; q["level"] = lv[:n]
    return q, v[:n], np.ascontiguousarray(d4[:, :3])


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref/libbow_ref.so: the reference's own DBoW2 TemplatedVocabulary.h (verbatim header template) behind oracle/bow_glue.cc
BOW_REF_LIB = os.path.join(HERE, "_ref", "libbow_ref.so")
_bref = None


def bow_ref():
    global _bref
    if _bref is None:
        if os.path.exists("/root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"):
            srcs = [os.path.join(HERE, "bow_glue.cc"), os.path.join(HERE, "slamshim", "opencv2", "core", "core.hpp")]
            if (not os.path.exists(BOW_REF_LIB)) or any(os.path.getmtime(BOW_REF_LIB) < os.path.getmtime(s) for s in srcs):
                subprocess.check_call(["make", "-C", HERE, "ref_bow"], stdout=subprocess.DEVNULL)
        if not os.path.exists(BOW_REF_LIB):
            return None
        R = C.CDLL(BOW_REF_LIB)
        R.bref_vocab_load.restype = C.c_void_p; R.bref_vocab_load.argtypes = [C.c_char_p]
        R.bref_vocab_destroy.argtypes = [C.c_void_p]
        R.bref_vocab_words.argtypes = [C.c_void_p]
        R.bref_vocab_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 7
        R.bref_vocab_score.restype = C.c_double
        R.bref_vocab_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _bref = R
    return _bref


def write_vocabulary_text(path, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0):
    """ORBvoc.txt format (TemplatedVocabulary.h:1338-1420): header `k L scoring weighting`, then one line per non-root node:
    parent id, leaf flag, 32 descriptor bytes, weight. No trailing newline: the loader's `while(!f.eof())` would read an
    empty last line into a phantom node."""
    lines = [f"{k} {L} {scoring} {weighting}"]
    for p, l, d, w in zip(parent, is_leaf, desc, weight):
        lines.append(f"{int(p)} {int(l)} " + " ".join(str(int(b)) for b in d) + f" {float(w)!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines))


class RefVocabulary:
    """The reference's ORBVocabulary (TemplatedVocabulary<FORB::TDescriptor, FORB>) loaded through its own text loader."""

    def __init__(self, path):
        self.R = bow_ref()
        self._h = self.R.bref_vocab_load(path.encode())
        assert self._h, "loadFromTextFile failed"
        self.nwords = self.R.bref_vocab_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self.R.bref_vocab_destroy(self._h); self._h = None

    def transform(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8); n = len(desc)
        bow_id = np.zeros(max(n, 1), np.int32); bow_val = np.zeros(max(n, 1), np.float64); fv_node = np.zeros(max(n, 1), np.int32)
        fv_off = np.zeros(n + 1, np.int32); fv_feat = np.zeros(max(n, 1), np.int32); nb = C.c_int32(0); nf = C.c_int32(0)
        self.R.bref_vocab_transform(self._h, desc.ctypes.data, n, levelsup, bow_id.ctypes.data, bow_val.ctypes.data, C.addressof(nb),
                                    fv_node.ctypes.data, fv_off.ctypes.data, fv_feat.ctypes.data, C.addressof(nf))
        return dict(bow_id=bow_id[:nb.value], bow_val=bow_val[:nb.value], fv_node=fv_node[:nf.value], fv_off=fv_off[:nf.value + 1],
                    fv_feat=fv_feat[:fv_off[nf.value]])

    def score(self, a, b):
        i1 = np.ascontiguousarray(a["bow_id"], np.int32); v1 = np.ascontiguousarray(a["bow_val"], np.float64)
        i2 = np.ascontiguousarray(b["bow_id"], np.int32); v2 = np.ascontiguousarray(b["bow_val"], np.float64)
        return self.R.bref_vocab_score(self._h, i1.ctypes.data, v1.ctypes.data, len(i1), i2.ctypes.data, v2.ctypes.data, len(i2))
