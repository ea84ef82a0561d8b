#!/usr/bin/env python3
"""Writes one synthetic SearchLocalPoints scene (the bench's: 1500 local map points x ~1631 keypoints) as a flat binary for
tools/native/bench_frame_threads.cu:  int32 n, nq, nlevels; then kps (28 B x n), desc (32 x n), u_right (f32 x n),
occupied (u8 x n), queries (20 B x nq), qdesc (32 x nq), qflags (u8 x nq), bounds4 (4 f32), scale_factors (f32 x nlevels)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from orb_slam2_commit_b200 import synth
sc = synth.synth_local_points_scene(21)
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/lp_scene.bin"
n, nq = len(sc["kps"]), len(sc["queries"]); sf = np.ascontiguousarray(sc["scale_factors"], np.float32)
with open(out, "wb") as f:
    f.write(np.array([n, nq, len(sf)], np.int32).tobytes())
    f.write(np.ascontiguousarray(sc["kps"]).tobytes()); f.write(np.ascontiguousarray(sc["desc"], np.uint8).tobytes())
    f.write(np.ascontiguousarray(sc["u_right"], np.float32).tobytes()); f.write(np.ascontiguousarray(sc["occupied"], np.uint8).tobytes())
    f.write(np.ascontiguousarray(sc["queries"]).tobytes()); f.write(np.ascontiguousarray(sc["query_desc"], np.uint8).tobytes())
    f.write(np.ascontiguousarray(sc["query_flags"], np.uint8).tobytes())
    f.write(np.ascontiguousarray(sc["bounds4"], np.float32).tobytes()); f.write(sf.tobytes())
print(out, n, nq, len(sf), sc["kps"].dtype.itemsize, sc["queries"].dtype.itemsize)
