"""CPU-only: the C++ host shim that mirrors the reference class (orb_slam2_commit_b200/host/ORBextractor.{h,cc})
compiles against a minimal OpenCV stand-in, links to liborbx.so, and keeps the reference's getter values and its
silent-return-on-empty-image behaviour. (Compute needs a GPU and is covered by tests/test_gpu_parity.py.)"""
import os
import subprocess
import textwrap

import __graft_entry__ as graft
from orb_slam2_commit_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "orb_slam2_commit_b200", "host")

MAIN = textwrap.dedent(r"""
    #include "ORBextractor.h"
    #include "HammingTop2.h"
    #include "StereoMatch.h"
    #include "FrameOps.h"
    #include "MatcherOps.h"
    #include <cstdio>
    int main() {
        ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
        std::vector<float> sf = ex.GetScaleFactors(), inv = ex.GetInverseScaleFactors();
        std::vector<float> s2 = ex.GetScaleSigmaSquares(), is2 = ex.GetInverseScaleSigmaSquares();
        std::printf("%d %.9g %.9g %.9g %.9g %.9g %zu\n", ex.GetLevels(), ex.GetScaleFactor(), sf[7], inv[7], s2[7], is2[7],
                    ex.mvImagePyramid.size());
        std::vector<cv::KeyPoint> kps(3);
        cv::Mat empty, desc;
        ex(empty, cv::Mat(), kps, desc);                 // ORBextractor.cc:1141: silent return, outputs untouched
        std::printf("%zu %d\n", kps.size(), (int)desc.empty());
        // the Frame / ORBmatcher wrappers: empty inputs return without touching the GPU
        std::vector<cv::KeyPoint> none, un; std::vector<int> match; std::vector<float> sf2(8, 1.f), xyz; std::vector<unsigned char> d8, fl;
        const float K4[4] = {500, 500, 320, 240}, D[5] = {0, 0, 0, 0, 0}, T[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}, cam[9] = {500, 500, 320, 240, 40, 0, 640, 0, 480};
        bool ok = ORB_SLAM2::UndistortKeyPointsGPU(none, K4, D, 5, un);
        float a, b, c, d; ok = ok && ORB_SLAM2::ComputeImageBoundsGPU(640, 480, K4, D, 5, a, b, c, d);
        int nm = ORB_SLAM2::SearchByProjectionGPU(none, desc, 0, 0, T, cam, sf2, none, xyz, d8, fl, 7.f, 0, true, match);
        std::map<unsigned int, double> bow; std::map<unsigned int, std::vector<unsigned int> > fv;
        ok = ok && ORB_SLAM2::ComputeBoWGPU(0, desc, bow, fv) && ORB_SLAM2::ExtractRectified(&ex, empty, kps, desc);
        std::printf("%d %d %.0f %.0f %.0f %.0f\n", (int)ok, nm, a, b, c, d);
        // the ORBmatcher wrappers (MatcherOps.h): empty inputs return 0 without touching the GPU
        ORB_SLAM2::MapPointTable pts; ORB_SLAM2::KeyFrameCamera kc; std::memcpy(kc.camera9, cam, sizeof cam);
        kc.mvScaleFactors = sf2; kc.mvLevelSigma2 = sf2; kc.mvInvLevelSigma2 = sf2; kc.mfLogScaleFactor = 0.18f;
        std::vector<OrbxTrackQuery> tq; std::vector<unsigned char> hm; std::vector<std::pair<size_t, size_t> > vp; std::vector<cv::Point2f> prev;
        const float b4[4] = {0, 640, 0, 480}, ow[3] = {0, 0, 0}, g28[28] = {0};
        int r = ORB_SLAM2::SearchByProjectionGPU(none, desc, 0, 0, b4, sf2, tq, d8, fl, 1.f, 0.8f, match);
        std::vector<unsigned char> inview; r += ORB_SLAM2::IsInFrustumGPU(T, ow, kc, pts, 0.5f, tq, inview) ? 0 : 100;
        r += ORB_SLAM2::FuseSearchGPU(none, desc, 0, T, ow, kc, pts, 3.f, false, match);
        r += ORB_SLAM2::SearchByProjectionGPU(none, desc, 0, T, ow, kc, pts, 0, 10.f, 100, false, true, match);
        r += ORB_SLAM2::SearchForTriangulationGPU(0, none, desc, hm, 0, none, desc, hm, 0, g28, kc, false, true, vp);
        r += ORB_SLAM2::SearchForInitializationGPU(none, desc, none, desc, b4, prev, match, 100, 0.9f, true);
        std::printf("%d %zu\n", r, sizeof(cv::Point2f));
        (void)&ORB_SLAM2::SearchBySim3GPU;
        return 0;
    }
""")


def test_shim_compiles_links_and_matches_getters(tmp_path):
    graft.build()
    src = tmp_path / "main.cc"
    src.write_text(MAIN)
    exe = tmp_path / "shim_test"
    libdir = os.path.dirname(api.library_path())
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-I", HOST, "-I", os.path.join(HOST, "cvmini"),
                           str(src), os.path.join(HOST, "ORBextractor.cc"), "-L", libdir, "-lorbx",
                           f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    f = out[0].split()
    assert f[0] == "8" and f[-1] == "8"
    import numpy as np
    from oracle import binding as ob
    t = ob.Extractor(1000, 1.2, 8, 20, 7).tables()
    assert np.float32(float(f[1])) == np.float32(1.2)
    assert np.float32(float(f[2])) == t["scale_factors"][7] and np.float32(float(f[3])) == t["inv_scale_factors"][7]
    assert np.float32(float(f[4])) == t["sigma2"][7] and np.float32(float(f[5])) == t["inv_sigma2"][7]
    assert out[1].split() == ["3", "1"]
    assert out[2].split() == ["1", "0", "0", "640", "0", "480"]       # Frame.cc:530-536 without distortion
    assert out[3].split() == ["0", "8"]                               # five matcher wrappers on empty inputs; Point2f = 2 packed floats
