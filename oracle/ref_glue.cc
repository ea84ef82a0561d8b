/*
 * ref_glue.cc — C entry points around the VERBATIM reference ORB_SLAM2::ORBextractor
 * (/root/reference/src/ORBextractor.cc compiled against oracle/cvshim). Test infrastructure only.
 *
 * Determinism: DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cc:733), i.e. breaks
 * count ties by list-node ADDRESS. In deterministic mode this library replaces global operator new with a
 * per-thread monotonic bump arena for the duration of one operator() call, so address order == creation
 * order — the tie rule the oracle and the CUDA path implement. With orbref_set_deterministic(0) the
 * allocator is plain malloc (what a stock build does; used for CPU-baseline timing).
 * Linked with -Wl,-Bsymbolic so the replacement binds inside this library only.
 */
#include <sys/mman.h>
#include <chrono>
#include <new>
#include <thread>
#include <vector>
#include "ORBextractor.h"

static int g_deterministic = 1;
static const size_t ARENA_BYTES = (size_t)16 << 30; /* virtual reservation, touched lazily */
static thread_local char* t_arena = 0;
static thread_local size_t t_off = 0;
static thread_local int t_arena_on = 0;

static inline bool in_arena(void* p) { return t_arena && (char*)p >= t_arena && (char*)p < t_arena + ARENA_BYTES; }

void* operator new(size_t n)
{
    if (t_arena_on) {
        size_t a = (t_off + 15) & ~(size_t)15;
        if (a + n <= ARENA_BYTES) { t_off = a + n; return t_arena + a; }
    }
    void* p = malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !in_arena(p)) free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

namespace {
struct ArenaScope {
    ArenaScope()
    {
        if (!g_deterministic) return;
        if (!t_arena) {
            void* m = mmap(0, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (m == MAP_FAILED) return;
            t_arena = (char*)m;
        }
        t_off = 0; t_arena_on = 1;
    }
    ~ArenaScope()
    {
        if (t_arena_on) { t_arena_on = 0; if (t_off > ((size_t)64 << 20)) madvise(t_arena, t_off, MADV_DONTNEED); t_off = 0; }
    }
};
int run(ORB_SLAM2::ORBextractor* e, const uint8_t* img, int w, int h, int stride, OcKeyPoint* kps, int cap, uint8_t* desc)
{
    cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
    cv::Mat descriptors;
    int n;
    {
        ArenaScope scope;
        {
            std::vector<cv::KeyPoint> keypoints;
            (*e)(image, cv::Mat(), keypoints, descriptors);
            n = (int)keypoints.size();
            if (n <= cap) {
                if (kps && n) memcpy(kps, keypoints.data(), sizeof(cv::KeyPoint) * (size_t)n);
                if (desc) for (int i = 0; i < n; i++) memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
            } else n = -1;
        }
    }
    return n;
}
} // namespace

extern "C" {
void orbref_set_deterministic(int on) { g_deterministic = on; }
void* orbref_create(int nfeatures, float scale, int nlevels, int ini, int min) { return new ORB_SLAM2::ORBextractor(nfeatures, scale, nlevels, ini, min); }
void orbref_destroy(void* h) { delete (ORB_SLAM2::ORBextractor*)h; }
int orbref_extract(void* h, const uint8_t* img, int w, int hgt, int stride, OcKeyPoint* kps, int cap, uint8_t* desc)
{
    return run((ORB_SLAM2::ORBextractor*)h, img, w, hgt, stride, kps, cap, desc);
}
int orbref_level(void* h, int level, int* w, int* hgt, int* stride, void** ptr)
{
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    if (level < 0 || level >= e->GetLevels() || e->mvImagePyramid[level].empty()) return -1;
    const cv::Mat& m = e->mvImagePyramid[level];
    *w = m.cols; *hgt = m.rows; *stride = (int)m.step; *ptr = m.data;
    return 0;
}
void orbref_tables(void* h, float* sf, float* inv, float* s2, float* is2)
{
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (size_t i = 0; i < a.size(); i++) { sf[i] = a[i]; inv[i] = b[i]; s2[i] = c[i]; is2[i] = d[i]; }
}
/* CPU baseline: `iters` frames (cycling over `nimgs` images of w x h, contiguous) spread over `nthreads`
 * host threads, one extractor instance per thread (frames are independent; the reference itself runs one
 * instance per thread for stereo, Frame.cc:80-84). Returns wall seconds (steady_clock, like the examples). */
double orbref_bench(int nfeatures, float scale, int nlevels, int ini, int min, const uint8_t* imgs, int w, int h,
                    int nimgs, int iters, int nthreads, long long* total_kp)
{
    int det = g_deterministic; g_deterministic = 0;
    std::vector<long long> kp((size_t)nthreads, 0);
    std::vector<ORB_SLAM2::ORBextractor*> ex;
    for (int t = 0; t < nthreads; t++) ex.push_back(new ORB_SLAM2::ORBextractor(nfeatures, scale, nlevels, ini, min));
    auto work = [&](int t) {
        std::vector<OcKeyPoint> k((size_t)nfeatures * 2 + 8192);
        std::vector<uint8_t> d(k.size() * 32);
        for (int i = t; i < iters; i += nthreads) {
            int n = run(ex[t], imgs + (size_t)(i % nimgs) * w * h, w, h, w, k.data(), (int)k.size(), d.data());
            if (n > 0) kp[t] += n;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long long tot = 0; for (auto v : kp) tot += v;
    if (total_kp) *total_kp = tot;
    for (auto p : ex) delete p;
    g_deterministic = det;
    return s;
}
}
