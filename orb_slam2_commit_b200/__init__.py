"""B200-native ORB feature front-end — drop-in for ORB-SLAM2's ORBextractor / Hamming best-second-best search.

The product is `csrc/liborbx.so` (hand-written sm_100a CUDA kernels behind the C ABI of include/orbx.h).
This package is only the thin host mirror used by the tests and bench.py: same names and argument meaning as the
reference class (include/ORBextractor.h:51-145). There is no CPU fallback: loading fails loudly when the
library is missing and every compute call fails when no CUDA device is visible.
"""
from .api import (KP_DTYPE, Frame, OrbxError, ORBextractor, ORBmatcher, ORBVocabulary, build_library, hamming_top2, lib, library_path,
                  image_bounds, is_in_frustum, fuse_search, search_by_projection_kf, search_by_sim3, search_for_initialization, search_by_projection_frame, search_local_points, TRACKQ_DTYPE, stereo_extract_host, stereo_extract_host_rectified, init_undistort_rectify_map, stereo_extract_host_begin, stereo_extract_host_end, stereo_hamming, stereo_match, stereo_match_device, undistort_keypoints, window_top2)

__all__ = ["KP_DTYPE", "Frame", "OrbxError", "ORBextractor", "ORBmatcher", "ORBVocabulary", "build_library", "hamming_top2", "lib",
           "library_path", "image_bounds", "is_in_frustum", "fuse_search", "search_by_projection_kf", "search_by_sim3", "search_for_initialization", "search_by_projection_frame", "search_local_points", "TRACKQ_DTYPE", "stereo_extract_host", "stereo_extract_host_rectified", "init_undistort_rectify_map", "stereo_extract_host_begin", "stereo_extract_host_end", "stereo_hamming", "stereo_match", "stereo_match_device", "undistort_keypoints",
           "window_top2"]
