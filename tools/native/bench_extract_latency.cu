// tools/native/bench_extract_latency.cu — latency and concurrency of the drop-in's own call, ORBextractor::operator() on ONE
// frame (orbx_extract: host image in, keypoints + descriptors out, synchronous), from native host threads: what a tracking
// thread of ORB-SLAM2 sees per frame, and what T such threads (left / right extractor, several cameras) sustain together.
// Build:  nvcc -O2 -std=c++17 -o /tmp/bench_extract_latency tools/native/bench_extract_latency.cu -Iinclude \
//              -Lorb_slam2_commit_b200/csrc -lorbx -Xlinker -rpath -Xlinker $PWD/orb_slam2_commit_b200/csrc
// Run:    /tmp/bench_extract_latency frame.bin width height [threads] [calls] [mirror]     (frame.bin = width*height gray bytes)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "orbx.h"

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s frame.bin width height [threads] [calls] [mirror]\n", argv[0]); return 2; }
    const int W = atoi(argv[2]), H = atoi(argv[3]);
    const int T = argc > 4 ? atoi(argv[4]) : 1, calls = argc > 5 ? atoi(argv[5]) : 1000, mirror = argc > 6 ? atoi(argv[6]) : 0;
    uint8_t* img = nullptr;
    cudaMallocHost(&img, (size_t)W * H);                        // pinned, as a capture pipeline would hand it over
    FILE* fp = fopen(argv[1], "rb");
    if (!fp || fread(img, 1, (size_t)W * H, fp) != (size_t)W * H) { fprintf(stderr, "cannot read frame\n"); return 2; }
    fclose(fp);
    std::vector<orbx_extractor*> ex(T);
    std::vector<int> nk(T, 0);
    int cap = 0;
    for (int t = 0; t < T; t++) {
        if (orbx_create(1000, 1.2f, 8, 20, 7, 0, &ex[t]) != ORBX_OK || orbx_reserve(ex[t], W, H, 1) != ORBX_OK) { fprintf(stderr, "%s\n", orbx_last_error()); return 1; }
        if (mirror) orbx_set_pyramid_mirror(ex[t], 1);
        cap = orbx_max_keypoints(ex[t]);
    }
    auto worker = [&](int t, int reps) {
        OrbxKeyPoint* k = nullptr; uint8_t* d = nullptr;
        cudaMallocHost(&k, sizeof(OrbxKeyPoint) * cap); cudaMallocHost(&d, (size_t)cap * 32);
        for (int i = 0; i < reps; i++)
            if (orbx_extract(ex[t], img, W, H, W, k, cap, &nk[t], d) != ORBX_OK) { fprintf(stderr, "%s\n", orbx_last_error()); exit(1); }
        cudaFreeHost(k); cudaFreeHost(d);
    };
    for (int t = 0; t < T; t++) worker(t, 20);                  // warm-up: includes the CUDA-graph capture of the single-frame path
    auto t0 = std::chrono::steady_clock::now();
    worker(0, calls);
    const double one = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<std::thread> th;
    t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++) th.emplace_back(worker, t, calls);
    for (auto& x : th) x.join();
    const double all = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"width\": %d, \"height\": %d, \"keypoints\": %d, \"pyramid_mirror\": %d, \"one_thread_ms_per_frame\": %.4f, \"threads\": %d, "
           "\"all_threads_frames_per_s\": %.1f, \"all_threads_ms_per_frame_per_thread\": %.4f}\n", W, H, nk[0], mirror, one / calls * 1e3, T,
           (double)T * calls / all, all / calls * 1e3);
    for (auto e : ex) orbx_destroy(e);
    return 0;
}
