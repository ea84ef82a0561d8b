// tools/native/bench_frame_threads.cu — how many one-frame SearchLocalPoints calls per second the library sustains when the
// callers are NATIVE host threads (one device-resident Frame handle and stream each), without the Python harness between the
// calls (16 Python threads share one interpreter lock and spend ~25 us per call in it: bench.py's frame_handle leg tops out
// near 40 k calls/s for that reason). Build on the GPU box:
//   nvcc -O2 -std=c++17 -o gpurun_out/bench_frame_threads tools/native/bench_frame_threads.cu -Iinclude \
//        -Lorb_slam2_commit_b200/csrc -lorbx -Xlinker -rpath -Xlinker $PWD/orb_slam2_commit_b200/csrc
// Run: gpurun_out/bench_frame_threads gpurun_out/lp_scene.bin [threads] [calls per thread]     (scene: tools/native/dump_scene.py)
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "orbx.h"

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s scene.bin [threads] [calls]\n", argv[0]); return 2; }
    const int T = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
    const int calls = argc > 3 ? atoi(argv[3]) : 2000;
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) { perror("scene"); return 2; }
    int hdr[3];
    if (fread(hdr, 4, 3, fp) != 3) return 2;
    const int n = hdr[0], nq = hdr[1], nl = hdr[2];
    std::vector<OrbxKeyPoint> kps(n); std::vector<uint8_t> desc((size_t)n * 32), occ(n), qdesc((size_t)nq * 32), qflags(nq);
    std::vector<float> ur(n), sf(nl); std::vector<OrbxTrackQuery> q(nq); float bounds4[4];
    bool ok = fread(kps.data(), sizeof(OrbxKeyPoint), n, fp) == (size_t)n && fread(desc.data(), 32, n, fp) == (size_t)n &&
              fread(ur.data(), 4, n, fp) == (size_t)n && fread(occ.data(), 1, n, fp) == (size_t)n &&
              fread(q.data(), sizeof(OrbxTrackQuery), nq, fp) == (size_t)nq && fread(qdesc.data(), 32, nq, fp) == (size_t)nq &&
              fread(qflags.data(), 1, nq, fp) == (size_t)nq && fread(bounds4, 4, 4, fp) == 4 && fread(sf.data(), 4, nl, fp) == (size_t)nl;
    fclose(fp);
    if (!ok) { fprintf(stderr, "short scene file\n"); return 2; }
    OrbxKeyPoint* d_k; uint8_t* d_d;
    cudaMalloc(&d_k, sizeof(OrbxKeyPoint) * n); cudaMalloc(&d_d, (size_t)n * 32);
    cudaMemcpy(d_k, kps.data(), sizeof(OrbxKeyPoint) * n, cudaMemcpyHostToDevice); cudaMemcpy(d_d, desc.data(), (size_t)n * 32, cudaMemcpyHostToDevice);
    std::vector<orbx_frame*> fr(T);
    for (int t = 0; t < T; t++) {
        if (orbx_frame_create(0, n, nq, &fr[t]) != ORBX_OK || orbx_frame_from_device(fr[t], d_k, d_d, n, nullptr, nullptr, 0, nullptr) != ORBX_OK ||
            orbx_frame_set_stereo(fr[t], ur.data(), 0) != ORBX_OK) { fprintf(stderr, "setup: %s\n", orbx_last_error()); return 1; }
    }
    cudaDeviceSynchronize();
    std::vector<int> nm(T, -1);
    auto worker = [&](int t, int reps) {
        std::vector<int32_t> match(n);
        for (int i = 0; i < reps; i++)
            if (orbx_frame_search_local_points(fr[t], q.data(), qdesc.data(), qflags.data(), nq, occ.data(), bounds4, sf.data(), nl, 1.0f, 0.8f,
                                               match.data(), &nm[t]) != ORBX_OK) { fprintf(stderr, "call: %s\n", orbx_last_error()); exit(1); }
    };
    worker(0, 50);                                              // warm-up
    auto t0 = std::chrono::steady_clock::now();
    worker(0, calls);
    const double one = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<std::thread> th;
    t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; t++) th.emplace_back(worker, t, calls);
    for (auto& x : th) x.join();
    const double all = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"n\": %d, \"nq\": %d, \"matches\": %d, \"one_thread_calls_per_s\": %.1f, \"one_thread_ms_per_call\": %.4f, \"threads\": %d, "
           "\"all_threads_calls_per_s\": %.1f}\n", n, nq, nm[0], calls / one, one / calls * 1e3, T, (double)T * calls / all);
    for (auto f : fr) orbx_frame_destroy(f);
    return 0;
}
