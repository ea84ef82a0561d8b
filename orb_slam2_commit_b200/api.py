"""ctypes binding of liborbx.so + the host-side mirror of the reference interface."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])  # == cv::KeyPoint, 28 bytes

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"orbx status {code}: {msg}")
        self.code = code


def library_path() -> str:
    return os.path.join(CSRC, "liborbx.so")


def build_library(force: bool = False) -> str:
    """Compile liborbx.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc"))]
    srcs.append(os.path.join(HERE, "..", "include", "orbx.h"))
    so = library_path()
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        args = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
        subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    """Load liborbx.so. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        so = library_path()
        if not os.path.exists(so):
            raise OrbxError(-1, f"{so} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
        L = C.CDLL(so)
        L.orbx_last_error.restype = C.c_char_p
        L.orbx_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.orbx_destroy.argtypes = [C.c_void_p]
        L.orbx_get_levels.argtypes = [C.c_void_p]
        L.orbx_get_scale_factor.argtypes = [C.c_void_p]; L.orbx_get_scale_factor.restype = C.c_float
        L.orbx_get_tables.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, i32p, i32p]
        L.orbx_reserve.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orbx_max_keypoints.argtypes = [C.c_void_p]
        L.orbx_set_pyramid_mirror.argtypes = [C.c_void_p, C.c_int]
        L.orbx_pyramid_mirror.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_uint8))]
        L.orbx_pyramid_level_layout.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        L.orbx_extract.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, i32p, u8p]
        L.orbx_extract_batch.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_int, i32p, C.c_void_p]
        L.orbx_extract_batch_color.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p, C.c_int, i32p, C.c_void_p]
        L.orbx_extract_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_set_rectify_maps.argtypes = [C.c_void_p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orbx_extract_batch_rectified.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_int, i32p,
                                                   C.c_void_p]
        L.orbx_extract_device_rectified.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_undistort_keypoints.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.c_void_p, C.c_int]
        L.orbx_undistort_keypoints_device.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbx_image_bounds.argtypes = [C.c_int, C.c_int, f32p, f32p, C.c_int, f32p, C.c_int]
        L.orbx_vocab_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, u8p, u8p, C.POINTER(C.c_double), C.c_int,
                                        C.POINTER(C.c_void_p)]
        L.orbx_vocab_destroy.argtypes = [C.c_void_p]; L.orbx_vocab_destroy.restype = None
        L.orbx_vocab_words.argtypes = [C.c_void_p]; L.orbx_vocab_nodes.argtypes = [C.c_void_p]
        L.orbx_vocab_lock.argtypes = [C.c_void_p]; L.orbx_vocab_unlock.argtypes = [C.c_void_p]
        L.orbx_bow_transform_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orbx_bow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orbx_bow_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        L.orbx_bow_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orbx_bow_score_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbx_search_by_bow_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_search_by_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.orbx_search_by_projection.argtypes = [C.c_void_p, f32p, f32p, C.c_int, C.c_float, C.c_int, C.c_int]
        L.orbx_search_by_projection_device.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orbx_search_by_bow_kf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.orbx_search_by_bow_kf_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_stereo_extract_batch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                                C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orbx_stereo_extract_batch_begin.argtypes = L.orbx_stereo_extract_batch.argtypes
        L.orbx_stereo_extract_batch_end.argtypes = [C.c_void_p]
        L.orbx_synchronize.argtypes = [C.c_void_p]
        L.orbx_enable_timing.argtypes = [C.c_void_p, C.c_int]
        L.orbx_get_stage_ms.argtypes = [C.c_void_p, f32p, i32p]
        L.orbx_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
        L.orbx_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.c_int, u8p, C.c_int]
        L.orbx_debug_blurred_level.argtypes = [C.c_void_p, C.c_int, C.c_int, u8p, C.c_int]
        L.orbx_pyramid_level_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), i32p]
        L.orbx_debug_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, i32p]
        L.orbx_debug_level_counts.argtypes = [C.c_void_p, C.c_int, i32p]
        L.orbx_hamming_top2.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, i32p, i32p, C.c_int]
        L.orbx_hamming_init_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orbx_hamming_top2_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.orbx_hamming_merge_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_stereo_hamming.argtypes = [C.c_void_p, u8p, C.c_int, C.c_void_p, u8p, C.c_int, C.c_int, f32p, C.c_int,
                                          C.c_float, C.c_float, i32p, i32p, C.c_int]
        L.orbx_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u8p, C.c_int, C.c_void_p, u8p, C.c_int,
                                        C.c_float, C.c_float, f32p, f32p]
        L.orbx_stereo_match_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_window_top2.argtypes = [C.c_void_p, u8p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                       C.c_void_p, u8p, C.c_int, i32p, i32p, i32p, i32p, i32p, C.c_int]
        L.orbx_search_local_points.argtypes = [C.c_void_p, f32p, f32p, C.c_int, C.c_float, C.c_float, C.c_int]
        L.orbx_search_local_points_device.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p]
        L.orbx_fuse_search.argtypes = [C.c_void_p, f32p, f32p, f32p, C.c_int, C.c_float, C.c_int]
        L.orbx_fuse_search_device.argtypes = [C.c_void_p, C.c_int, f32p, f32p, f32p, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.orbx_search_for_triangulation.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, f32p, f32p, f32p, C.c_int, C.c_int,
                                                    C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orbx_search_for_triangulation_device.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [f32p, f32p, C.c_int, C.c_int,
                                                                                                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_search_by_projection_kf.argtypes = [C.c_void_p, f32p, f32p, C.c_int, C.c_float, C.c_int, C.c_int]
        L.orbx_search_by_projection_kf_device.argtypes = [C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orbx_search_by_sim3.argtypes = [C.c_void_p, C.c_void_p, f32p, f32p, f32p, f32p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                          C.c_void_p, C.c_int]
        L.orbx_search_by_sim3_device.argtypes = [C.c_void_p, C.c_void_p, f32p, f32p, f32p, f32p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orbx_search_for_initialization.argtypes = [C.c_void_p, f32p, C.c_float, C.c_int, C.c_int]
        L.orbx_search_for_initialization_device.argtypes = [C.c_void_p, C.c_int, f32p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        L.orbx_is_in_frustum.argtypes = [f32p, f32p, f32p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_int]
        L.orbx_is_in_frustum_device.argtypes = [f32p, f32p, f32p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                                C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orbx_extract_batch_begin.argtypes = L.orbx_extract_batch.argtypes
        L.orbx_extract_batch_end.argtypes = [C.c_void_p]
        L.orbx_peer_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_char_p]
        L.orbx_peer_connect.argtypes = [C.c_void_p, C.c_char_p]
        L.orbx_peer_hamming_top2.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orbx_peer_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise OrbxError(rc, lib().orbx_last_error().decode())


def _u8(a):
    return a.ctypes.data_as(u8p)


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:51-145): same constructor arguments, same getters,
    `__call__(image, mask)` -> (keypoints, descriptors) like operator(), `mvImagePyramid` like the public member."""

    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int, device: int = 0):
        self._L = lib()
        self._h = C.c_void_p()
        _ck(self._L.orbx_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device, C.byref(self._h)))
        self.nfeatures, self.nlevels, self.device = nfeatures, nlevels, device

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.orbx_destroy(h)
            self._h = None

    # ---- getters (ORBextractor.h:81-101)
    def GetLevels(self) -> int:
        return self._L.orbx_get_levels(self._h)

    def GetScaleFactor(self) -> float:
        return self._L.orbx_get_scale_factor(self._h)

    def _tables(self):
        n = self.nlevels
        a = [np.zeros(n, np.float32) for _ in range(4)]
        fpl = np.zeros(n, np.int32); um = np.zeros(16, np.int32)
        _ck(self._L.orbx_get_tables(self._h, *[x.ctypes.data_as(f32p) for x in a], fpl.ctypes.data_as(i32p), um.ctypes.data_as(i32p)))
        return a + [fpl, um]

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        return self._tables()[4]

    def umax(self):
        return self._tables()[5]

    # ---- operator() (ORBextractor.h:77)
    def __call__(self, image: np.ndarray, mask=None):
        """image: 2-D uint8 array (any row stride). mask is ignored, as in the reference (ORBextractor.h:68).
        Returns (keypoints[KP_DTYPE], descriptors[n,32] uint8); an empty image returns empty outputs."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.ndim == 3:
            return self.extract_color(image, rgb=True)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 required (ORBextractor.cc:1146)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        _ck(self._L.orbx_reserve(self._h, w, h, 1))
        cap = self._L.orbx_max_keypoints(self._h)
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = C.c_int32(0)
        _ck(self._L.orbx_extract(self._h, _u8(image), w, h, image.strides[0], kps.ctypes.data, cap, C.byref(n), _u8(desc)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_color(self, image: np.ndarray, rgb: bool):
        """(h, w, 3|4) interleaved uint8 frame: Tracking::GrabImage*'s cvtColor (Tracking.cc:174-199) fused into level 0.
        rgb=True <=> channel 0 is R (the reference's mbRGB)."""
        image = np.ascontiguousarray(image, np.uint8)
        h, w, ch = image.shape
        _ck(self._L.orbx_reserve(self._h, w, h, 1))
        cap = self._L.orbx_max_keypoints(self._h)
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8); n = np.zeros(1, np.int32)
        ptrs = (C.c_void_p * 1)(image.ctypes.data)
        _ck(self._L.orbx_extract_batch_color(self._h, ptrs, 1, w, h, w * ch, ch, int(rgb), kps.ctypes.data, cap,
                                             n.ctypes.data_as(i32p), desc.ctypes.data))
        return kps[:n[0]].copy(), desc[:n[0]].copy()

    def extract_batch(self, images, max_batch: int = 64):
        """images: sequence of equally sized 2-D uint8 arrays, or one (n,h,w) array. Returns lists."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        if n == 0:
            return [], []
        h, w = imgs[0].shape
        _ck(self._L.orbx_reserve(self._h, w, h, min(n, max_batch)))
        cap = self._L.orbx_max_keypoints(self._h)
        kps = np.zeros((n, cap), KP_DTYPE); desc = np.zeros((n, cap, 32), np.uint8); nk = np.zeros(n, np.int32)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        _ck(self._L.orbx_extract_batch(self._h, ptrs, n, w, h, w, kps.ctypes.data, cap, nk.ctypes.data_as(i32p), desc.ctypes.data))
        return [kps[i, :nk[i]].copy() for i in range(n)], [desc[i, :nk[i]].copy() for i in range(n)]

    # ---- stereo rectification in front of the extractor (Examples/Stereo/stereo_euroc.cc:97-98, 136-137)
    def set_rectify_maps(self, map1, map2, src_size=None):
        """map1 / map2: the CV_32FC1 pair of cv::initUndistortRectifyMap (None, None clears). src_size = (w, h) of the
        unrectified frames (default: the maps' own size)."""
        if map1 is None and map2 is None:
            _ck(self._L.orbx_set_rectify_maps(self._h, None, None, 0, 0, 0, 0, 0)); return
        map1 = np.ascontiguousarray(map1, np.float32); map2 = np.ascontiguousarray(map2, np.float32)
        assert map1.shape == map2.shape and map1.ndim == 2
        mh, mw = map1.shape
        sw, sh = src_size if src_size is not None else (mw, mh)
        _ck(self._L.orbx_set_rectify_maps(self._h, map1.ctypes.data_as(f32p), map2.ctypes.data_as(f32p), mw, mh, mw, sw, sh))
        self._map_size = (mw, mh)

    def set_rectify_camera(self, K, D, R, P, map_size, src_size=None):
        """initUndistortRectifyMap(K, D, R, P, map_size) built on the device and installed (stereo_euroc.cc:96-97); sizes = (w, h)."""
        f64p = C.POINTER(C.c_double)
        K = np.ascontiguousarray(K, np.float64).reshape(3, 3); D = np.ascontiguousarray(D, np.float64).ravel()
        R = np.ascontiguousarray(R, np.float64).reshape(3, 3); P = np.ascontiguousarray(P, np.float64)
        mw, mh = map_size
        sw, sh = src_size if src_size is not None else map_size
        self._L.orbx_set_rectify_camera.argtypes = [C.c_void_p, f64p, f64p, C.c_int, f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _ck(self._L.orbx_set_rectify_camera(self._h, K.ctypes.data_as(f64p), D.ctypes.data_as(f64p) if len(D) else None, len(D),
                                            R.ctypes.data_as(f64p), P.ctypes.data_as(f64p), P.shape[1], mw, mh, sw, sh))
        self._map_size = (mw, mh)

    def extract_rectified(self, images):
        """images: (n, h, w) or a list of UNRECTIFIED uint8 frames; cv::remap(.., INTER_LINEAR) is fused into level 0."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        n = len(imgs)
        mw, mh = self._map_size
        _ck(self._L.orbx_reserve(self._h, mw, mh, min(n, 64)))
        cap = self._L.orbx_max_keypoints(self._h)
        kps = np.zeros((n, cap), KP_DTYPE); desc = np.zeros((n, cap, 32), np.uint8); nk = np.zeros(n, np.int32)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        _ck(self._L.orbx_extract_batch_rectified(self._h, ptrs, n, imgs[0].shape[1], kps.ctypes.data, cap,
                                                 nk.ctypes.data_as(i32p), desc.ctypes.data))
        return [kps[i, :nk[i]].copy() for i in range(n)], [desc[i, :nk[i]].copy() for i in range(n)]

    def extract_device_rectified(self, d_images, n, stride, frame_pitch, d_kps, cap, d_nkp, d_desc, stream=0):
        _ck(self._L.orbx_extract_device_rectified(self._h, d_images, n, stride, frame_pitch, d_kps, cap, d_nkp, d_desc, stream))

    # ---- device-resident form used by bench.py (pointers are plain integers, e.g. torch.Tensor.data_ptr())
    def reserve(self, width, height, max_batch):
        _ck(self._L.orbx_reserve(self._h, width, height, max_batch))
        return self._L.orbx_max_keypoints(self._h)

    def extract_device(self, d_images, n, width, height, stride, frame_pitch, d_kps, cap, d_nkp, d_desc, stream=0):
        _ck(self._L.orbx_extract_device(self._h, d_images, n, width, height, stride, frame_pitch, d_kps, cap, d_nkp, d_desc, stream))

    def synchronize(self):
        _ck(self._L.orbx_synchronize(self._h))

    def enable_timing(self, on=True):
        _ck(self._L.orbx_enable_timing(self._h, int(on)))

    def stage_ms(self):
        ms = np.zeros(4, np.float32); n = C.c_int32(0)
        _ck(self._L.orbx_get_stage_ms(self._h, ms.ctypes.data_as(f32p), C.byref(n)))
        return ms, n.value

    def extract_host(self, images: np.ndarray, kps: np.ndarray, desc: np.ndarray, nkp: np.ndarray):
        """orbx_extract_batch on caller-owned (ideally pinned) buffers: images (n,h,w) u8 contiguous,
        kps (n,cap) KP_DTYPE, desc (n,cap,32) u8, nkp (n,) i32. No allocation, no copies on the Python side."""
        n, h, w = images.shape
        cap = kps.shape[1]
        base = images.ctypes.data
        ptrs = (C.c_void_p * n)(*[base + i * h * w for i in range(n)])
        _ck(self._L.orbx_extract_batch(self._h, ptrs, n, w, h, w, kps.ctypes.data, cap, nkp.ctypes.data_as(i32p), desc.ctypes.data))

    def extract_host_begin(self, images: np.ndarray, kps: np.ndarray, desc: np.ndarray, nkp: np.ndarray):
        """orbx_extract_batch_begin: enqueue one batch (buffers as for extract_host, cap = reserve()) and return; up to two
        batches in flight. The buffers belong to the batch until the matching extract_host_end()."""
        n, h, w = images.shape
        cap = kps.shape[1]
        base = images.ctypes.data
        ptrs = (C.c_void_p * n)(*[base + i * h * w for i in range(n)])
        _ck(self._L.orbx_extract_batch_begin(self._h, ptrs, n, w, h, w, kps.ctypes.data, cap, nkp.ctypes.data_as(i32p), desc.ctypes.data))

    def extract_host_end(self):
        """orbx_extract_batch_end: wait for the oldest batch begun."""
        _ck(self._L.orbx_extract_batch_end(self._h))

    def extract_host_rectified(self, images: np.ndarray, kps: np.ndarray, desc: np.ndarray, nkp: np.ndarray):
        """orbx_extract_batch_rectified on caller-owned buffers (see extract_host); images are UNRECTIFIED frames."""
        n, h, w = images.shape
        cap = kps.shape[1]
        base = images.ctypes.data
        ptrs = (C.c_void_p * n)(*[base + i * h * w for i in range(n)])
        _ck(self._L.orbx_extract_batch_rectified(self._h, ptrs, n, w, kps.ctypes.data, cap, nkp.ctypes.data_as(i32p), desc.ctypes.data))

    # ---- mvImagePyramid (ORBextractor.h:104)
    def level_size(self, level):
        w = C.c_int32(); h = C.c_int32()
        _ck(self._L.orbx_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level(self, level: int, frame: int = 0, with_apron: bool = False) -> np.ndarray:
        w, h = self.level_size(level)
        whole = np.zeros((h + 38, w + 38), np.uint8)
        _ck(self._L.orbx_pyramid_level(self._h, frame, level, _u8(whole), w + 38))
        return whole if with_apron else whole[19:19 + h, 19:19 + w]

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    # ---- stage taps for the parity tests
    def blurred_level(self, level: int, frame: int = 0) -> np.ndarray:
        """The level after cv::GaussianBlur 7x7 sigma 2 (`workingMat`, ORBextractor.cc:1188-1190)."""
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        _ck(self._L.orbx_debug_blurred_level(self._h, frame, level, _u8(out), w))
        return out

    def debug_candidates(self, level: int, frame: int = 0) -> np.ndarray:
        n = C.c_int32(0)
        _ck(self._L.orbx_debug_candidates(self._h, frame, level, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), KP_DTYPE)
        _ck(self._L.orbx_debug_candidates(self._h, frame, level, out.ctypes.data, n.value, C.byref(n)))
        return out[:n.value]

    def set_pyramid_mirror(self, on: bool):
        """Single-frame host calls also bring the frame's raw pyramid block to a pinned host mirror (one asynchronous copy)."""
        _ck(self._L.orbx_set_pyramid_mirror(self._h, int(bool(on))))

    def pyramid_mirror(self, frame: int = 0):
        """-> list of per-level uint8 views (h x w, strides (pitch, 1)) into the handle's pinned mirror of `frame`'s pyramid;
        the 19-px apron lies around every view in memory. Valid until the next call on the extractor."""
        blk = C.POINTER(C.c_uint8)()
        _ck(self._L.orbx_pyramid_mirror(self._h, frame, C.byref(blk)))
        out = []
        for l in range(self.nlevels):
            w = C.c_int(0); h = C.c_int(0); pitch = C.c_int(0); off = C.c_size_t(0); tot = C.c_size_t(0)
            _ck(self._L.orbx_level_size(self._h, l, C.byref(w), C.byref(h)))
            _ck(self._L.orbx_pyramid_level_layout(self._h, l, C.byref(off), C.byref(pitch), C.byref(tot)))
            whole = np.ctypeslib.as_array(blk, shape=(tot.value,))
            out.append(np.lib.stride_tricks.as_strided(whole[off.value:], shape=(h.value, w.value), strides=(pitch.value, 1), writeable=False))
        return out

    def debug_level_counts(self, frame: int = 0):
        c = np.zeros(self.nlevels, np.int32)
        _ck(self._L.orbx_debug_level_counts(self._h, frame, c.ctypes.data_as(i32p)))
        return c


class Frame:
    """Device-resident Frame (include/orbx.h orbx_frame_*; Frame.cc:62-123): keypoints, descriptors, undistorted keypoints and
    mvuRight stay in HBM; matcher calls upload only the projected map points."""

    def __init__(self, max_keypoints: int, max_queries: int, device: int = 0):
        self._L = lib(); self._h = C.c_void_p()
        L = self._L
        L.orbx_frame_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.orbx_frame_destroy.argtypes = [C.c_void_p]; L.orbx_frame_destroy.restype = None
        L.orbx_frame_from_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, f32p, f32p, C.c_int]
        L.orbx_frame_from_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.c_void_p]
        L.orbx_frame_set_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orbx_frame_size.argtypes = [C.c_void_p]
        L.orbx_frame_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, i32p]
        L.orbx_frame_search_local_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, f32p, f32p, C.c_int,
                                                     C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        _ck(L.orbx_frame_create(device, max_keypoints, max_queries, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orbx_frame_destroy(self._h); self._h = None

    @staticmethod
    def _kd(K4, dist):
        K = None if K4 is None else np.ascontiguousarray(K4, np.float32)
        D = None if dist is None else np.ascontiguousarray(dist, np.float32)
        return K, D, (0 if D is None else len(D))

    def from_extract(self, ex: "ORBextractor", frame_index: int, n: int, K4=None, dist=None):
        K, D, nd = self._kd(K4, dist)
        _ck(self._L.orbx_frame_from_extract(self._h, ex._h, frame_index, n, None if K is None else K.ctypes.data_as(f32p),
                                            None if D is None else D.ctypes.data_as(f32p), nd))

    def from_device(self, d_kps: int, d_desc: int, n: int, K4=None, dist=None, stream: int = 0):
        K, D, nd = self._kd(K4, dist)
        _ck(self._L.orbx_frame_from_device(self._h, d_kps, d_desc, n, None if K is None else K.ctypes.data_as(f32p),
                                           None if D is None else D.ctypes.data_as(f32p), nd, stream))

    def set_stereo(self, u_right):
        if u_right is None:
            _ck(self._L.orbx_frame_set_stereo(self._h, None, 0)); return
        u = np.ascontiguousarray(u_right, np.float32)
        _ck(self._L.orbx_frame_set_stereo(self._h, u.ctypes.data, 0))

    def __len__(self):
        return self._L.orbx_frame_size(self._h)

    def keypoints(self):
        n = len(self)
        k = np.zeros(max(n, 1), KP_DTYPE); d = np.zeros((max(n, 1), 32), np.uint8); m = C.c_int32(0)
        _ck(self._L.orbx_frame_keypoints(self._h, k.ctypes.data, d.ctypes.data, max(n, 1), C.byref(m)))
        return k[:n], d[:n]

    def search_local_points(self, queries, query_descriptors, query_flags, occupied, bounds4, scale_factors, th, nnratio=0.8,
                            out=None):
        """ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th) -> (nmatches, match). `out` (int32, >= n) avoids an allocation."""
        q = np.ascontiguousarray(queries); qd = np.ascontiguousarray(query_descriptors, np.uint8); qf = np.ascontiguousarray(query_flags, np.uint8)
        oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        b4 = np.ascontiguousarray(bounds4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        n = len(self)
        match = out if out is not None else np.empty(max(n, 1), np.int32)
        nm = C.c_int32(0)
        _ck(self._L.orbx_frame_search_local_points(self._h, q.ctypes.data, qd.ctypes.data, qf.ctypes.data, len(q),
                                                   None if oc is None else oc.ctypes.data, b4.ctypes.data_as(f32p), sf.ctypes.data_as(f32p),
                                                   len(sf), th, nnratio, match.ctypes.data, C.addressof(nm)))
        return nm.value, match[:n]


def hamming_top2(query: np.ndarray, train: np.ndarray, device: int = 0):
    """Best / second-best Hamming search (ORBmatcher.cc:84-126 idiom over DescriptorDistance :1844-1860).
    Returns (idx1, dist1, dist2) int32 arrays."""
    q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    nq, nt = len(q), len(t)
    idx = np.full(nq, -1, np.int32); d1 = np.full(nq, 256, np.int32); d2 = np.full(nq, 256, np.int32)
    if nq:
        _ck(lib().orbx_hamming_top2(_u8(q), nq, _u8(t) if nt else None, nt, idx.ctypes.data_as(i32p),
                                    d1.ctypes.data_as(i32p), d2.ctypes.data_as(i32p), device))
    return idx, d1, d2


def stereo_hamming(kp_left, desc_left, kp_right, desc_right, rows, scale_factors, minD, maxD, device: int = 0):
    """Hamming stage of Frame::ComputeStereoMatches (Frame.cc:554-663). Returns (best_idx_right, best_dist)."""
    kl = np.ascontiguousarray(kp_left, KP_DTYPE); kr = np.ascontiguousarray(kp_right, KP_DTYPE)
    dl = np.ascontiguousarray(desc_left, np.uint8); dr = np.ascontiguousarray(desc_right, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    bi = np.full(len(kl), -1, np.int32); bd = np.full(len(kl), 100, np.int32)
    _ck(lib().orbx_stereo_hamming(kl.ctypes.data, _u8(dl), len(kl), kr.ctypes.data, _u8(dr), len(kr), rows,
                                  sf.ctypes.data_as(f32p), len(sf), minD, maxD, bi.ctypes.data_as(i32p),
                                  bd.ctypes.data_as(i32p), device))
    return bi, bd


def stereo_match(left: ORBextractor, right: ORBextractor, kp_left, desc_left, kp_right, desc_right, mbf: float, fx: float):
    """Frame::ComputeStereoMatches (Frame.cc:547-788) on the HBM-resident pyramids of two extractors that have just
    extracted the left / right image. Returns (mvuRight, mvDepth): float32 arrays, -1 where unmatched."""
    kl = np.ascontiguousarray(kp_left, KP_DTYPE); kr = np.ascontiguousarray(kp_right, KP_DTYPE)
    dl = np.ascontiguousarray(desc_left, np.uint8); dr = np.ascontiguousarray(desc_right, np.uint8)
    ur = np.full(len(kl), -1, np.float32); dp = np.full(len(kl), -1, np.float32)
    _ck(lib().orbx_stereo_match(left._h, right._h, kl.ctypes.data, _u8(dl), len(kl), kr.ctypes.data, _u8(dr), len(kr),
                                mbf, fx, ur.ctypes.data_as(f32p), dp.ctypes.data_as(f32p)))
    return ur, dp


def stereo_match_device(left: ORBextractor, right: ORBextractor, pairs, d_kl, d_dl, d_nl, d_kr, d_dr, d_nr, cap, mbf, fx,
                        d_u_right, d_depth, stream=0):
    """Batched device-resident Frame::ComputeStereoMatches (orbx_stereo_match_device); all d_* are device pointers."""
    _ck(lib().orbx_stereo_match_device(left._h, right._h, pairs, d_kl, d_dl, d_nl, d_kr, d_dr, d_nr, cap, mbf, fx,
                                       d_u_right, d_depth, stream))


WQ_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("r", "<f4"), ("min_level", "<i4"), ("max_level", "<i4"), ("xr", "<f4")])


def undistort_keypoints(keypoints, K4, dist, device: int = 0):
    """Frame::UndistortKeyPoints (Frame.cc:471-506): mvKeys -> mvKeysUn. K4 = (fx, fy, cx, cy), dist = (k1, k2, p1, p2[, k3])."""
    kps = np.ascontiguousarray(keypoints, KP_DTYPE)
    K4 = np.ascontiguousarray(K4, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(kps)
    _ck(lib().orbx_undistort_keypoints(kps.ctypes.data, len(kps), K4.ctypes.data_as(f32p), dist.ctypes.data_as(f32p), len(dist),
                                       out.ctypes.data, device))
    return out


def image_bounds(width, height, K4, dist, device: int = 0):
    """Frame::ComputeImageBounds (Frame.cc:508-538) -> (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
    K4 = np.ascontiguousarray(K4, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    b = np.zeros(4, np.float32)
    _ck(lib().orbx_image_bounds(width, height, K4.ctypes.data_as(f32p), dist.ctypes.data_as(f32p), len(dist),
                                b.ctypes.data_as(f32p), device))
    return b


def stereo_extract_host_begin(left: ORBextractor, right: ORBextractor, images_left: np.ndarray, images_right: np.ndarray, mbf: float,
                              fx: float, out: dict):
    """orbx_stereo_extract_batch_begin (buffers as for stereo_extract_host); up to two batches in flight, each owning its `out`."""
    n, h, w = images_left.shape
    cap = out["kl"].shape[1]
    pl = (C.c_void_p * n)(*[images_left.ctypes.data + i * h * w for i in range(n)])
    pr = (C.c_void_p * n)(*[images_right.ctypes.data + i * h * w for i in range(n)])
    _ck(lib().orbx_stereo_extract_batch_begin(left._h, right._h, pl, pr, n, w, h, w, mbf, fx, out["kl"].ctypes.data, out["dl"].ctypes.data,
                                              out["nl"].ctypes.data, out["kr"].ctypes.data, out["dr"].ctypes.data, out["nr"].ctypes.data,
                                              cap, out["u_right"].ctypes.data, out["depth"].ctypes.data))


def stereo_extract_host_end(left: ORBextractor):
    """orbx_stereo_extract_batch_end: wait for the oldest batch of pairs begun."""
    _ck(lib().orbx_stereo_extract_batch_end(left._h))


def stereo_extract_host(left: ORBextractor, right: ORBextractor, images_left: np.ndarray, images_right: np.ndarray, mbf: float,
                        fx: float, out: dict):
    """orbx_stereo_extract_batch on caller-owned (ideally pinned) buffers: images (n,h,w) u8; out = dict with kl / kr
    (n,cap) KP_DTYPE, dl / dr (n,cap,32) u8, nl / nr (n,) i32, u_right / depth (n,cap) f32; cap = left.reserve(...)."""
    n, h, w = images_left.shape
    cap = out["kl"].shape[1]
    pl = (C.c_void_p * n)(*[images_left.ctypes.data + i * h * w for i in range(n)])
    pr = (C.c_void_p * n)(*[images_right.ctypes.data + i * h * w for i in range(n)])
    _ck(lib().orbx_stereo_extract_batch(left._h, right._h, pl, pr, n, w, h, w, mbf, fx, out["kl"].ctypes.data, out["dl"].ctypes.data,
                                        out["nl"].ctypes.data, out["kr"].ctypes.data, out["dr"].ctypes.data, out["nr"].ctypes.data,
                                        cap, out["u_right"].ctypes.data, out["depth"].ctypes.data))


def stereo_extract_host_rectified(left: ORBextractor, right: ORBextractor, images_left: np.ndarray, images_right: np.ndarray, mbf: float,
                                  fx: float, out: dict):
    """orbx_stereo_extract_batch_rectified: UNRECTIFIED frames (n, src_h, src_w) in, both extractors carry their camera's
    rectification (set_rectify_camera / set_rectify_maps); out as in stereo_extract_host, cap = reserve(map_w, map_h, ..)."""
    n, h, w = images_left.shape
    cap = out["kl"].shape[1]
    pl = (C.c_void_p * n)(*[images_left.ctypes.data + i * h * w for i in range(n)])
    pr = (C.c_void_p * n)(*[images_right.ctypes.data + i * h * w for i in range(n)])
    L = lib()
    L.orbx_stereo_extract_batch_rectified.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                                      C.c_float, C.c_float] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p]
    _ck(L.orbx_stereo_extract_batch_rectified(left._h, right._h, pl, pr, n, w, mbf, fx, out["kl"].ctypes.data, out["dl"].ctypes.data,
                                              out["nl"].ctypes.data, out["kr"].ctypes.data, out["dr"].ctypes.data, out["nr"].ctypes.data,
                                              cap, out["u_right"].ctypes.data, out["depth"].ctypes.data))


def init_undistort_rectify_map(K, D, R, P, size, device: int = 0):
    """cv::initUndistortRectifyMap(K, D, R, P, size, CV_32FC1) on the device -> (map1, map2) float32 (h, w); size = (w, h)."""
    f64p = C.POINTER(C.c_double)
    K = np.ascontiguousarray(K, np.float64).reshape(3, 3); D = np.ascontiguousarray(D, np.float64).ravel()
    R = np.ascontiguousarray(R, np.float64).reshape(3, 3); P = np.ascontiguousarray(P, np.float64)
    w, h = size
    m1 = np.empty((h, w), np.float32); m2 = np.empty((h, w), np.float32)
    L = lib()
    L.orbx_init_undistort_rectify_map.argtypes = [f64p, f64p, C.c_int, f64p, f64p, C.c_int, C.c_int, C.c_int, f32p, f32p, C.c_int]
    _ck(L.orbx_init_undistort_rectify_map(K.ctypes.data_as(f64p), D.ctypes.data_as(f64p) if len(D) else None, len(D), R.ctypes.data_as(f64p),
                                          P.ctypes.data_as(f64p), P.shape[1], w, h, m1.ctypes.data_as(f32p), m2.ctypes.data_as(f32p), device))
    return m1, m2


def window_top2(keypoints, descriptors, occupied, u_right, minX, minY, invW, invH, queries, query_descriptors, device: int = 0):
    """Candidate loop of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) over Frame::GetFeaturesInArea.
    Returns (bestIdx, bestDist, bestLevel, bestDist2, bestLevel2) int32 arrays, one entry per query."""
    kps = np.ascontiguousarray(keypoints, KP_DTYPE); desc = np.ascontiguousarray(descriptors, np.uint8)
    q = np.ascontiguousarray(queries, WQ_DTYPE); qd = np.ascontiguousarray(query_descriptors, np.uint8)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    out = [np.empty(len(q), np.int32) for _ in range(5)]
    _ck(lib().orbx_window_top2(kps.ctypes.data, _u8(desc), len(kps), None if occ is None else occ.ctypes.data,
                               None if ur is None else ur.ctypes.data, minX, minY, invW, invH, q.ctypes.data, _u8(qd), len(q),
                               *[o.ctypes.data_as(i32p) for o in out], device))
    return out


class OrbxProjectionPair(C.Structure):
    """include/orbx.h OrbxProjectionPair"""
    _fields_ = [("cur_keypoints", C.c_void_p), ("cur_descriptors", C.c_void_p), ("cur_u_right", C.c_void_p),
                ("cur_occupied", C.c_void_p), ("n_cur", C.c_int32),
                ("last_keypoints", C.c_void_p), ("last_xyz", C.c_void_p), ("last_descriptors", C.c_void_p),
                ("last_flags", C.c_void_p), ("n_last", C.c_int32),
                ("Tcw", C.c_float * 12), ("mode", C.c_int32), ("match", C.c_void_p), ("nmatches", C.c_void_p)]


def search_by_projection_frame(cur_kps, cur_desc, cur_u_right, cur_occupied, Tcw12, cam9, scale_factors, last_kps, last_xyz,
                               last_desc, last_flags, th, mode, check_orientation=True, device: int = 0):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1489-1646) -> (nmatches, match_cur);
    match_cur[k] = index of the last-frame keypoint whose map point current keypoint k now holds, -1 = none."""
    cur_kps = np.ascontiguousarray(cur_kps, KP_DTYPE); last_kps = np.ascontiguousarray(last_kps, KP_DTYPE)
    cur_desc = np.ascontiguousarray(cur_desc, np.uint8); last_desc = np.ascontiguousarray(last_desc, np.uint8)
    ur = None if cur_u_right is None else np.ascontiguousarray(cur_u_right, np.float32)
    occ = None if cur_occupied is None else np.ascontiguousarray(cur_occupied, np.uint8)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    xyz = np.ascontiguousarray(last_xyz, np.float32); fl = np.ascontiguousarray(last_flags, np.uint8)
    match = np.zeros(max(len(cur_kps), 1), np.int32); nm = np.zeros(1, np.int32)
    P = OrbxProjectionPair()
    P.cur_keypoints = cur_kps.ctypes.data; P.cur_descriptors = cur_desc.ctypes.data
    P.cur_u_right = None if ur is None else ur.ctypes.data; P.cur_occupied = None if occ is None else occ.ctypes.data
    P.n_cur = len(cur_kps)
    P.last_keypoints = last_kps.ctypes.data; P.last_xyz = xyz.ctypes.data; P.last_descriptors = last_desc.ctypes.data
    P.last_flags = fl.ctypes.data; P.n_last = len(last_kps)
    P.Tcw = (C.c_float * 12)(*np.asarray(Tcw12, np.float32).ravel().tolist()); P.mode = mode
    P.match = match.ctypes.data; P.nmatches = nm.ctypes.data
    _ck(lib().orbx_search_by_projection(C.byref(P), cam.ctypes.data_as(f32p), sf.ctypes.data_as(f32p), len(sf), th,
                                        int(check_orientation), device))
    return int(nm[0]), match[:len(cur_kps)]


TRACKQ_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4")])


def is_in_frustum(Tcw12, Ow3, cam9, nlevels, log_scale_factor, pt_xyz, pt_normal, pt_dist, viewing_cos_limit=0.5, device: int = 0):
    """Frame::isInFrustum (Frame.cc:315-378) for every map point -> (queries TRACKQ_DTYPE, in_view uint8); queries[i] is
    meaningful where in_view[i] != 0 (the reference leaves the mTrack* fields untouched otherwise)."""
    T = np.ascontiguousarray(Tcw12, np.float32); Ow = np.ascontiguousarray(Ow3, np.float32); cam = np.ascontiguousarray(cam9, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = np.ascontiguousarray(pt_normal, np.float32); dst = np.ascontiguousarray(pt_dist, np.float32)
    n = len(xyz)
    q = np.zeros(max(n, 1), TRACKQ_DTYPE); v = np.zeros(max(n, 1), np.uint8)
    _ck(lib().orbx_is_in_frustum(T.ctypes.data_as(f32p), Ow.ctypes.data_as(f32p), cam.ctypes.data_as(f32p), nlevels, float(log_scale_factor),
                                 xyz.ctypes.data, nrm.ctypes.data, dst.ctypes.data, n, viewing_cos_limit, q.ctypes.data, v.ctypes.data, device))
    return q[:n], v[:n]


class OrbxLocalPointsFrame(C.Structure):
    """include/orbx.h OrbxLocalPointsFrame"""
    _fields_ = [("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("u_right", C.c_void_p), ("occupied", C.c_void_p),
                ("n", C.c_int32), ("queries", C.c_void_p), ("query_descriptors", C.c_void_p), ("query_flags", C.c_void_p),
                ("nq", C.c_int32), ("match", C.c_void_p), ("nmatches", C.c_void_p)]


def search_local_points(kps, desc, u_right, occupied, bounds4, scale_factors, queries, query_desc, query_flags, th, nnratio=0.8,
                        device: int = 0):
    """ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:46-142) -> (nmatches, match);
    match[k] = index into vpMapPoints that keypoint k holds after the call, -1 = untouched."""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    b4 = np.ascontiguousarray(bounds4, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    q = np.ascontiguousarray(queries, TRACKQ_DTYPE); qd = np.ascontiguousarray(query_desc, np.uint8)
    qf = np.ascontiguousarray(query_flags, np.uint8)
    match = np.zeros(max(len(kps), 1), np.int32); nm = np.zeros(1, np.int32)
    F = OrbxLocalPointsFrame()
    F.keypoints = kps.ctypes.data; F.descriptors = desc.ctypes.data
    F.u_right = None if ur is None else ur.ctypes.data; F.occupied = None if occ is None else occ.ctypes.data; F.n = len(kps)
    F.queries = q.ctypes.data; F.query_descriptors = qd.ctypes.data; F.query_flags = qf.ctypes.data; F.nq = len(q)
    F.match = match.ctypes.data; F.nmatches = nm.ctypes.data
    _ck(lib().orbx_search_local_points(C.byref(F), b4.ctypes.data_as(f32p), sf.ctypes.data_as(f32p), len(sf), th, nnratio, device))
    return int(nm[0]), match[:len(kps)]


class OrbxFuseJob(C.Structure):
    """include/orbx.h OrbxFuseJob"""
    _fields_ = [("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("u_right", C.c_void_p), ("n", C.c_int32),
                ("Tcw", C.c_float * 12), ("Ow", C.c_float * 3),
                ("pt_xyz", C.c_void_p), ("pt_normal", C.c_void_p), ("pt_dist", C.c_void_p), ("pt_descriptors", C.c_void_p),
                ("pt_flags", C.c_void_p), ("npts", C.c_int32), ("th", C.c_float), ("mode", C.c_int32),
                ("best_idx", C.c_void_p), ("best_dist", C.c_void_p), ("nfused", C.c_void_p)]


def fuse_search(kps, desc, u_right, Tcw12, Ow3, cam9, scale_factors, inv_level_sigma2, log_scale_factor, pt_xyz, pt_normal,
                pt_dist, pt_desc, pt_flags, th, mode=0, device: int = 0):
    """Search half of ORBmatcher::Fuse (ORBmatcher.cc:918-1092 mode 0, :1094-1236 mode 1) -> (nFused, best_idx, best_dist)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    s2 = np.ascontiguousarray(inv_level_sigma2, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = np.ascontiguousarray(pt_normal, np.float32)
    dst = np.ascontiguousarray(pt_dist, np.float32); pd = np.ascontiguousarray(pt_desc, np.uint8)
    pf = np.ascontiguousarray(pt_flags, np.uint8)
    npts = len(pf)
    bi = np.zeros(max(npts, 1), np.int32); bd = np.zeros(max(npts, 1), np.int32); nf = np.zeros(1, np.int32)
    J = OrbxFuseJob()
    J.keypoints = kps.ctypes.data; J.descriptors = desc.ctypes.data; J.u_right = None if ur is None else ur.ctypes.data; J.n = len(kps)
    J.Tcw = (C.c_float * 12)(*np.asarray(Tcw12, np.float32).ravel().tolist())
    J.Ow = (C.c_float * 3)(*np.asarray(Ow3, np.float32).ravel().tolist())
    J.pt_xyz = xyz.ctypes.data; J.pt_normal = nrm.ctypes.data; J.pt_dist = dst.ctypes.data; J.pt_descriptors = pd.ctypes.data
    J.pt_flags = pf.ctypes.data; J.npts = npts; J.th = th; J.mode = mode
    J.best_idx = bi.ctypes.data; J.best_dist = bd.ctypes.data; J.nfused = nf.ctypes.data
    _ck(lib().orbx_fuse_search(C.byref(J), cam.ctypes.data_as(f32p), sf.ctypes.data_as(f32p), s2.ctypes.data_as(f32p), len(sf),
                               float(log_scale_factor), device))
    return int(nf[0]), bi[:npts], bd[:npts]


class OrbxProjectionJob(C.Structure):
    """include/orbx.h OrbxProjectionJob"""
    _fields_ = [("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("occupied", C.c_void_p), ("n", C.c_int32),
                ("Tcw", C.c_float * 12), ("Ow", C.c_float * 3),
                ("pt_xyz", C.c_void_p), ("pt_normal", C.c_void_p), ("pt_dist", C.c_void_p), ("pt_descriptors", C.c_void_p),
                ("pt_flags", C.c_void_p), ("pt_angle", C.c_void_p), ("npts", C.c_int32), ("th", C.c_float),
                ("max_dist", C.c_int32), ("mode", C.c_int32), ("match", C.c_void_p), ("nmatches", C.c_void_p)]


def search_by_projection_kf(kps, desc, occupied, Tcw12, Ow3, cam9, scale_factors, log_scale_factor, pt_xyz, pt_normal, pt_dist,
                            pt_desc, pt_flags, pt_angle, th, max_dist, mode, check_orientation=True, device: int = 0):
    """mode 0: ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1648-1795);
    mode 1: ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (:327-440). -> (nmatches, match)"""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    xyz = np.ascontiguousarray(pt_xyz, np.float32); nrm = None if pt_normal is None else np.ascontiguousarray(pt_normal, np.float32)
    dst = np.ascontiguousarray(pt_dist, np.float32); pd = np.ascontiguousarray(pt_desc, np.uint8)
    pf = np.ascontiguousarray(pt_flags, np.uint8); pa = None if pt_angle is None else np.ascontiguousarray(pt_angle, np.float32)
    match = np.zeros(max(len(kps), 1), np.int32); nm = np.zeros(1, np.int32)
    J = OrbxProjectionJob()
    J.keypoints = kps.ctypes.data; J.descriptors = desc.ctypes.data; J.occupied = None if occ is None else occ.ctypes.data; J.n = len(kps)
    J.Tcw = (C.c_float * 12)(*np.asarray(Tcw12, np.float32).ravel().tolist())
    J.Ow = (C.c_float * 3)(*np.asarray(Ow3, np.float32).ravel().tolist())
    J.pt_xyz = xyz.ctypes.data; J.pt_normal = None if nrm is None else nrm.ctypes.data; J.pt_dist = dst.ctypes.data
    J.pt_descriptors = pd.ctypes.data; J.pt_flags = pf.ctypes.data; J.pt_angle = None if pa is None else pa.ctypes.data
    J.npts = len(pf); J.th = th; J.max_dist = max_dist; J.mode = mode
    J.match = match.ctypes.data; J.nmatches = nm.ctypes.data
    _ck(lib().orbx_search_by_projection_kf(C.byref(J), cam.ctypes.data_as(f32p), sf.ctypes.data_as(f32p), len(sf),
                                           float(log_scale_factor), int(check_orientation), device))
    return int(nm[0]), match[:len(kps)]


class OrbxSim3KeyFrame(C.Structure):
    """include/orbx.h OrbxSim3KeyFrame"""
    _fields_ = [("keypoints", C.c_void_p), ("descriptors", C.c_void_p), ("n", C.c_int32), ("mp_xyz", C.c_void_p),
                ("mp_dist", C.c_void_p), ("mp_descriptors", C.c_void_p), ("mp_flags", C.c_void_p), ("Tcw", C.c_float * 12)]


def search_by_sim3(kf1, kf2, S12, S21, cam9, scale_factors, log_scale_factor, th, device: int = 0):
    """ORBmatcher::SearchBySim3 (ORBmatcher.cc:1238-1487). kf1 / kf2: dicts with kps, desc, mp_xyz, mp_dist, mp_desc, mp_flags,
    Tcw12. S12 = (sR12, t12), S21 = (sR21, t21): 12 floats each. -> (nFound, match12)"""
    keep = []

    def pack(k):
        K = OrbxSim3KeyFrame()
        a = [np.ascontiguousarray(k["kps"], KP_DTYPE), np.ascontiguousarray(k["desc"], np.uint8), np.ascontiguousarray(k["mp_xyz"], np.float32),
             np.ascontiguousarray(k["mp_dist"], np.float32), np.ascontiguousarray(k["mp_desc"], np.uint8), np.ascontiguousarray(k["mp_flags"], np.uint8)]
        keep.append(a)
        K.keypoints, K.descriptors, K.mp_xyz, K.mp_dist, K.mp_descriptors, K.mp_flags = [x.ctypes.data for x in a]
        K.n = len(a[0]); K.Tcw = (C.c_float * 12)(*np.asarray(k["Tcw12"], np.float32).ravel().tolist())
        return K
    K1, K2 = pack(kf1), pack(kf2)
    cam = np.ascontiguousarray(cam9, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    s12 = np.ascontiguousarray(S12, np.float32); s21 = np.ascontiguousarray(S21, np.float32)
    match = np.zeros(max(K1.n, 1), np.int32); nf = np.zeros(1, np.int32)
    _ck(lib().orbx_search_by_sim3(C.byref(K1), C.byref(K2), s12.ctypes.data_as(f32p), s21.ctypes.data_as(f32p), cam.ctypes.data_as(f32p),
                                  sf.ctypes.data_as(f32p), len(sf), float(log_scale_factor), th, match.ctypes.data, nf.ctypes.data, device))
    return int(nf[0]), match[:K1.n]


class OrbxInitPair(C.Structure):
    """include/orbx.h OrbxInitPair"""
    _fields_ = [("keypoints1", C.c_void_p), ("descriptors1", C.c_void_p), ("n1", C.c_int32),
                ("keypoints2", C.c_void_p), ("descriptors2", C.c_void_p), ("n2", C.c_int32),
                ("prev_matched", C.c_void_p), ("prev_matched_out", C.c_void_p), ("window_size", C.c_int32),
                ("match12", C.c_void_p), ("nmatches", C.c_void_p)]


def search_for_initialization(kps1, desc1, kps2, desc2, bounds4, prev_matched, window_size=100, nnratio=0.9, check_orientation=True,
                              device: int = 0):
    """ORBmatcher(nnratio, checkOri).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)
    (ORBmatcher.cc:442-587) -> (nmatches, vnMatches12, updated vbPrevMatched)."""
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    b4 = np.ascontiguousarray(bounds4, np.float32)
    prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    match = np.zeros(max(len(kps1), 1), np.int32); nm = np.zeros(1, np.int32)
    Q = OrbxInitPair()
    Q.keypoints1 = kps1.ctypes.data; Q.descriptors1 = desc1.ctypes.data; Q.n1 = len(kps1)
    Q.keypoints2 = kps2.ctypes.data; Q.descriptors2 = desc2.ctypes.data; Q.n2 = len(kps2)
    Q.prev_matched = prev.ctypes.data; Q.prev_matched_out = prev.ctypes.data; Q.window_size = window_size
    Q.match12 = match.ctypes.data; Q.nmatches = nm.ctypes.data
    _ck(lib().orbx_search_for_initialization(C.byref(Q), b4.ctypes.data_as(f32p), nnratio, int(check_orientation), device))
    return int(nm[0]), match[:len(kps1)], prev


class ORBVocabulary:
    """DBoW2 vocabulary on the GPU (include/ORBVocabulary.h: TemplatedVocabulary<FORB::TDescriptor, FORB>).
    parent / is_leaf / desc / weight: one entry per non-root node in ORBvoc.txt order (loadFromTextFile)."""

    def __init__(self, k, L, parent, is_leaf, desc, weight, scoring=0, weighting=0, device=0):
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        self._L = lib(); self._h = C.c_void_p()
        _ck(self._L.orbx_vocab_create(k, L, scoring, weighting, len(parent), parent.ctypes.data_as(i32p), _u8(is_leaf), _u8(desc),
                                      weight.ctypes.data_as(C.POINTER(C.c_double)), device, C.byref(self._h)))
        self.nwords = self._L.orbx_vocab_words(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orbx_vocab_destroy(self._h); self._h = None

    def transform_batch(self, descs, levelsup=4):
        """descs: list of (n_i, 32) uint8 arrays -> list of dicts (word, node, bow_id, bow_val, fv_node, fv_off, fv_feat)."""
        frames = len(descs)
        cap = max(1, max(len(d) for d in descs))
        buf = np.zeros((frames, cap, 32), np.uint8); counts = np.zeros(frames, np.int32)
        for i, d in enumerate(descs):
            buf[i, :len(d)] = d; counts[i] = len(d)
        # transform -> get is one critical section on the handle (it keeps the results of the LAST transform)
        _ck(self._L.orbx_vocab_lock(self._h))
        try:
            _ck(self._L.orbx_bow_transform(self._h, buf.ctypes.data, counts.ctypes.data, frames, cap, levelsup))
            return [self.get(i, int(counts[i])) for i in range(frames)]
        finally:
            self._L.orbx_vocab_unlock(self._h)

    def transform(self, desc, levelsup=4):
        """Frame::ComputeBoW: mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)."""
        return self.transform_batch([np.ascontiguousarray(desc, np.uint8)], levelsup)[0]

    def transform_device(self, d_desc, d_counts, frames, cap, levelsup=4, stream=0):
        _ck(self._L.orbx_bow_transform_device(self._h, d_desc, d_counts, frames, cap, levelsup, stream))

    def get(self, frame, n):
        word = np.zeros(n, np.int32); node = np.zeros(n, np.int32); bow_id = np.zeros(n, np.int32); bow_val = np.zeros(n, np.float64)
        fv_node = np.zeros(n, np.int32); fv_off = np.zeros(n + 1, np.int32); fv_feat = np.zeros(n, np.int32)
        nb = C.c_int32(0); nf = C.c_int32(0)
        _ck(self._L.orbx_bow_get(self._h, frame, word.ctypes.data, node.ctypes.data, bow_id.ctypes.data, bow_val.ctypes.data,
                                 C.addressof(nb), fv_node.ctypes.data, fv_off.ctypes.data, fv_feat.ctypes.data, C.addressof(nf)))
        return dict(word=word, node=node, bow_id=bow_id[:nb.value], bow_val=bow_val[:nb.value], fv_node=fv_node[:nf.value],
                    fv_off=fv_off[:nf.value + 1], fv_feat=fv_feat[:fv_off[nf.value]])

    def score(self, frame_a, frame_b):
        """L1Scoring::score between frames of the last transform_batch (KeyFrameDatabase.cc:145)."""
        a = np.ascontiguousarray(frame_a, np.int32); b = np.ascontiguousarray(frame_b, np.int32)
        out = np.zeros(len(a), np.float64)
        _ck(self._L.orbx_bow_score(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return out

    def score_device(self, d_a, d_b, npairs, d_score, stream=0):
        _ck(self._L.orbx_bow_score_device(self._h, d_a, d_b, npairs, d_score, stream))

    def search_by_bow(self, kf_kps, kf_desc, kf_valid, f_kps, f_desc, levelsup=4, nnratio=0.7, check_orientation=True):
        """ORBmatcher(nnratio, checkOri).SearchByBoW(pKF, F, vpMapPointMatches) -> (nmatches, match_f)."""
        kf_kps = np.ascontiguousarray(kf_kps, KP_DTYPE); f_kps = np.ascontiguousarray(f_kps, KP_DTYPE)
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8); f_desc = np.ascontiguousarray(f_desc, np.uint8)
        val = None if kf_valid is None else np.ascontiguousarray(kf_valid, np.uint8)
        match = np.zeros(max(len(f_kps), 1), np.int32); nm = C.c_int32(0)
        _ck(self._L.orbx_search_by_bow(self._h, kf_kps.ctypes.data, kf_desc.ctypes.data, len(kf_kps),
                                       None if val is None else val.ctypes.data, f_kps.ctypes.data, f_desc.ctypes.data, len(f_kps),
                                       levelsup, nnratio, int(check_orientation), match.ctypes.data, C.addressof(nm)))
        return nm.value, match[:len(f_kps)]

    def search_by_bow_kf(self, kps1, desc1, valid1, kps2, desc2, valid2, levelsup=4, nnratio=0.75, check_orientation=True):
        """ORBmatcher(nnratio, checkOri).SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:589-736) -> (nmatches, match12)."""
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        v1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8)
        v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
        match = np.zeros(max(len(kps1), 1), np.int32); nm = C.c_int32(0)
        _ck(self._L.orbx_search_by_bow_kf(self._h, kps1.ctypes.data, desc1.ctypes.data, len(kps1), None if v1 is None else v1.ctypes.data,
                                          kps2.ctypes.data, desc2.ctypes.data, len(kps2), None if v2 is None else v2.ctypes.data,
                                          levelsup, nnratio, int(check_orientation), match.ctypes.data, C.addressof(nm)))
        return nm.value, match[:len(kps1)]

    def search_for_triangulation(self, kps1, desc1, has_mp1, u_right1, kps2, desc2, has_mp2, u_right2, geom28, scale_factors,
                                 level_sigma2, levelsup=4, only_stereo=False, check_orientation=True):
        """ORBmatcher(0.6, checkOri).SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (ORBmatcher.cc:738-916)
        -> (nmatches, match12); vMatchedPairs = [(i, match12[i]) for i where match12[i] >= 0]."""
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
        desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
        m1 = None if has_mp1 is None else np.ascontiguousarray(has_mp1, np.uint8)
        m2 = None if has_mp2 is None else np.ascontiguousarray(has_mp2, np.uint8)
        r1 = None if u_right1 is None else np.ascontiguousarray(u_right1, np.float32)
        r2 = None if u_right2 is None else np.ascontiguousarray(u_right2, np.float32)
        g = np.ascontiguousarray(geom28, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
        s2 = np.ascontiguousarray(level_sigma2, np.float32)
        match = np.zeros(max(len(kps1), 1), np.int32); nm = C.c_int32(0)
        p = lambda a: None if a is None else a.ctypes.data
        _ck(self._L.orbx_search_for_triangulation(self._h, kps1.ctypes.data, desc1.ctypes.data, len(kps1), p(m1), p(r1), kps2.ctypes.data,
                                                  desc2.ctypes.data, len(kps2), p(m2), p(r2), g.ctypes.data_as(f32p),
                                                  sf.ctypes.data_as(f32p), s2.ctypes.data_as(f32p), len(sf), levelsup, int(only_stereo),
                                                  int(check_orientation), match.ctypes.data, C.addressof(nm)))
        return nm.value, match[:len(kps1)]

    def search_by_bow_device(self, npairs, d_kf_frame, d_f_frame, d_kps, d_desc, d_kf_valid, nnratio, check_orientation, d_match,
                             d_nmatches, stream=0):
        _ck(self._L.orbx_search_by_bow_device(self._h, npairs, d_kf_frame, d_f_frame, d_kps, d_desc, d_kf_valid, nnratio,
                                              int(check_orientation), d_match, d_nmatches, stream))


class ORBmatcher:
    """The part of ORB_SLAM2::ORBmatcher that is on the hot path (include/ORBmatcher.h:50, ORBmatcher.cc:37-38)."""
    TH_LOW = 50
    TH_HIGH = 100

    @staticmethod
    def DescriptorDistance(a: np.ndarray, b: np.ndarray, device: int = 0) -> int:
        """static int DescriptorDistance(const cv::Mat&, const cv::Mat&) — one 256-bit pair, on the GPU."""
        _, d1, _ = hamming_top2(np.asarray(a, np.uint8).reshape(1, 32), np.asarray(b, np.uint8).reshape(1, 32), device)
        return int(d1[0])
