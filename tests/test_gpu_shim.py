"""The real drop-in on a GPU (-m gpu): orb_slam2_commit_b200/host/ORBextractor.{h,cc} — the class Frame.cc / Tracking.cc
would link against — is compiled against the minimal OpenCV stand-in (cvmini), linked to liborbx.so and RUN on the seeded
TUM1 and KITTI frames. Its operator() outputs (std::vector<cv::KeyPoint>, the CV_8U N x 32 descriptor Mat) and every
mvImagePyramid[l] — read the way Frame::ComputeStereoMatches reads it (Frame.cc:681-700): through the cv::Mat header,
INCLUDING the 19-px apron around the ROI — are compared with the committed fixtures of the verbatim reference
(tests/golden/ref_*.npz: keypoints in the reference's order, descriptors, CRC of every level with its apron)."""
import json
import os
import subprocess
import textwrap
import zlib

import numpy as np
import pytest

from orb_slam2_commit_b200 import api, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "orb_slam2_commit_b200", "host")

MAIN = textwrap.dedent(r"""
    #include "ORBextractor.h"
    #include <chrono>
    #include <cstdio>
    #include <cstddef>
    #include <cstdlib>
    #include <string>
    #include <vector>
    // usage: shim_run <in.raw> <w> <h> <nfeatures> <scale> <nlevels> <ini> <min> <out-prefix> <download-pyramid>
    int main(int argc, char** argv) {
        if (argc < 11) return 2;
        const int w = atoi(argv[2]), h = atoi(argv[3]);
        std::vector<unsigned char> buf((size_t)w * h);
        FILE* f = fopen(argv[1], "rb"); if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) return 3; fclose(f);
        ORB_SLAM2::ORBextractor ex(atoi(argv[4]), (float)atof(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]));
        ex.mbDownloadPyramid = atoi(argv[10]) != 0;
        cv::Mat image(h, w, CV_8UC1, buf.data(), (size_t)w);
        std::vector<cv::KeyPoint> kps; cv::Mat desc;
        for (int rep = 0; rep < 3; rep++) ex(image, cv::Mat(), kps, desc);     // the third call replays the recorded CUDA graph
        const int reps = 200;
        const auto t0 = std::chrono::steady_clock::now();
        for (int rep = 0; rep < reps; rep++) ex(image, cv::Mat(), kps, desc);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        std::string p(argv[9]);
        f = fopen((p + ".kps").c_str(), "wb"); fwrite(kps.data(), sizeof(cv::KeyPoint), kps.size(), f); fclose(f);
        f = fopen((p + ".desc").c_str(), "wb");
        for (int i = 0; i < desc.rows; i++) fwrite(desc.ptr(i), 1, 32, f);
        fclose(f);
        if (ex.mbDownloadPyramid)
            for (int l = 0; l < ex.GetLevels(); l++) {
                const cv::Mat& m = ex.mvImagePyramid[l];
                char name[512]; snprintf(name, sizeof name, "%s.lvl%d", p.c_str(), l);
                f = fopen(name, "wb");
                int hdr[2] = {m.cols, m.rows}; fwrite(hdr, 4, 2, f);
                // Frame.cc:681-700 style: pointer arithmetic through the header, apron included
                for (int y = -19; y < m.rows + 19; y++) fwrite(m.data + (std::ptrdiff_t)y * (std::ptrdiff_t)m.step - 19, 1, (size_t)m.cols + 38, f);
                fclose(f);
            }
        std::printf("%zu %d %.4f\n", kps.size(), desc.rows, ms);
        return 0;
    }
""")


@pytest.fixture(scope="module")
def shim_exe(tmp_path_factory):
    d = tmp_path_factory.mktemp("shim")
    src = d / "main.cc"
    src.write_text(MAIN)
    exe = d / "shim_run"
    libdir = os.path.dirname(api.library_path())
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-Wall", "-I", HOST, "-I", os.path.join(HOST, "cvmini"), str(src),
                           os.path.join(HOST, "ORBextractor.cc"), "-L", libdir, "-lorbx", f"-Wl,-rpath,{libdir}", "-o", str(exe)])
    return str(exe)


def _run(exe, tmp_path, name, seed, pyramid):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
    c = json.loads(str(g["cfg"]))
    img = synth.synth_image(c["width"], c["height"], seed)
    assert zlib.crc32(img.tobytes()) == int(g["img_crc"])
    raw = tmp_path / f"{name}.raw"
    img.tofile(raw)
    prefix = str(tmp_path / f"{name}_out")
    out = subprocess.check_output([exe, str(raw), str(c["width"]), str(c["height"]), str(c["nfeatures"]), str(c["scale"]),
                                   str(c["nlevels"]), str(c["ini_th"]), str(c["min_th"]), prefix, "1" if pyramid else "0"]).decode().split()
    return g, c, prefix, int(out[0]), int(out[1]), float(out[2])


@pytest.mark.parametrize("name,seed", [("tum1", 1), ("kitti", 2)])
def test_shim_operator_matches_reference_fixtures(shim_exe, tmp_path, name, seed):
    g, c, prefix, n, nd, ms = _run(shim_exe, tmp_path, name, seed, True)
    gk, gd = g["keypoints"], g["descriptors"]
    assert n == len(gk) == nd
    kps = np.fromfile(prefix + ".kps", dtype=api.KP_DTYPE)
    desc = np.fromfile(prefix + ".desc", dtype=np.uint8).reshape(-1, 32)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], gk[f]), f
    assert np.abs(kps["angle"] - gk["angle"]).max() <= 1e-3            # north_star tolerance, degrees
    assert np.array_equal(desc, gd)
    for l in range(c["nlevels"]):
        blob = np.fromfile(f"{prefix}.lvl{l}", dtype=np.uint8)
        w, h = np.frombuffer(blob[:8].tobytes(), np.int32)
        whole = blob[8:].reshape(h + 38, w + 38)
        assert tuple(whole.shape) == tuple(g["level_whole_shape"][l])
        assert zlib.crc32(whole.tobytes()) == int(g["level_crc"][l]), f"mvImagePyramid[{l}] read through the cv::Mat header (apron included)"
    print(f"shim latency {name}: {ms:.3f} ms per operator() call with the pyramid mirror")


def test_shim_without_pyramid_download(shim_exe, tmp_path):
    g, c, prefix, n, nd, ms = _run(shim_exe, tmp_path, "tum1", 1, False)
    assert n == len(g["keypoints"]) == nd
    assert np.array_equal(np.fromfile(prefix + ".desc", dtype=np.uint8).reshape(-1, 32), g["descriptors"])
    assert not os.path.exists(prefix + ".lvl0")
    print(f"shim latency tum1: {ms:.3f} ms per operator() call without the pyramid")
