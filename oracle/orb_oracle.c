/*
 * orb_oracle.c — CPU ORACLE (test infrastructure, NOT product code; see orb_oracle.h).
 *
 * Restates, function by function, the reference hot path (file:line relative to /root/reference):
 *   src/ORBextractor.cc  IC_Angle :77-105, computeOrbDescriptor :110-152, ctor :416-490,
 *                        DivideNode :501-560, DistributeOctTree :562-815,
 *                        ComputeKeyPointsOctTree :818-946, operator() :1138-1211, ComputePyramid :1215-1250
 *   src/ORBmatcher.cc    DescriptorDistance :1844-1860, best/second-best idiom :84-126
 *   src/Frame.cc         ComputeStereoMatches Hamming stage :554-663
 * and the un-vendored third-party arithmetic those lines call:
 *   OpenCV 4.13 (cv2 4.13.0 is the only OpenCV in the authoring image; reference says "2.4.3 or 3.x",
 *   CMakeLists.txt:31-37): resize INTER_LINEAR 8U, copyMakeBorder REFLECT_101, FAST 9-16 + NMS,
 *   GaussianBlur 7x7 sigma 2 (fixed-point path), fastAtan2, cvRound.
 *   glibc 2.39 x86_64 cosf/sinf (FMA ifunc variant), restated from its double-precision polynomial.
 *
 * Deterministic choice: the reference breaks count ties in DistributeOctTree's
 * careful pass by comparing list-node ADDRESSES (ORBextractor.cc:733 sorts pair<int,ExtractorNode*>).
 * The oracle uses the node CREATION SEQUENCE instead; oracle/_ref reproduces exactly that by running the
 * verbatim reference under a monotonic bump allocator.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off: the pinned semantics are the uncontracted ones (contraction note).
 */
#include "orb_oracle.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define PATCH_SIZE 31
#define HALF_PATCH_SIZE 15
#define EDGE_THRESHOLD 19

static const int8_t g_pattern[1024] = {
#include "orb_pattern.inc"
};

/* ------------------------------------------------------------------ cvRound / fastAtan2 / cosf / sinf */

int oc_round_f(float v) { return (int)lrintf(v); } /* SSE cvtss2si: round-half-even */

static int oc_round_d(double v) { return (int)lrint(v); }

/* cv::fastAtan2 scalar path (OpenCV core mathfuncs_core: atanImpl<float>) */
float oc_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* glibc 2.39 sincosf tables (sysdeps/ieee754/flt-32/s_sincosf_data.c), as IEEE-754 bit patterns.
 * Entry 1 is entry 0 with the cosine coefficients negated. */
static double dbits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
typedef struct { double sign[4], hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3; } sincos_tab;
static sincos_tab g_sc[2];
static pthread_once_t g_sc_once = PTHREAD_ONCE_INIT;
static void sc_init(void)
{
    for (int t = 0; t < 2; t++) {
        double sg = t ? -1.0 : 1.0;
        g_sc[t].sign[0] = 1.0; g_sc[t].sign[1] = -1.0; g_sc[t].sign[2] = -1.0; g_sc[t].sign[3] = 1.0;
        g_sc[t].hpi_inv = dbits(0x41645f306dc9c883ull); /* 2/pi * 2^24 */
        g_sc[t].hpi = dbits(0x3ff921fb54442d18ull);     /* pi/2 */
        g_sc[t].c0 = sg * 1.0;
        g_sc[t].c1 = sg * dbits(0xbfdffffffd0c621cull);
        g_sc[t].c2 = sg * dbits(0x3fa55553e1068f19ull);
        g_sc[t].c3 = sg * dbits(0xbf56c087e89a359dull);
        g_sc[t].c4 = sg * dbits(0x3ef99343027bf8c3ull);
        g_sc[t].s1 = dbits(0xbfc555545995a603ull);
        g_sc[t].s2 = dbits(0x3f81107605230bc4ull);
        g_sc[t].s3 = dbits(0xbf2994eb3774cf24ull);
    }
}
static inline double sc_cos_poly(double x2, const sincos_tab* p)
{
    double x4 = x2 * x2;
    double c2 = fma(p->c4, x2, p->c3);
    double c1 = fma(p->c1, x2, p->c0);
    double x6 = x2 * x4;
    double c = fma(x4, p->c2, c1);
    return fma(c2, x6, c);
}
static inline double sc_sin_poly(double x, double x2, const sincos_tab* p)
{
    double s1 = fma(p->s3, x2, p->s2);
    double x3 = x2 * x;
    double x7 = x2 * x3;
    double s = fma(x3, p->s1, x);
    return fma(s1, x7, s);
}
static inline uint32_t abstop12(float x) { uint32_t u; memcpy(&u, &x, 4); return (u >> 20) & 0x7ff; }
/* valid for |x| < 120 (the descriptor angle is in [0, 2*pi]) */
static double sc_reduce(double x, int* np)
{
    double r = x * g_sc[0].hpi_inv;
    int n = ((int32_t)r + 0x800000) >> 24;
    *np = n;
    return fma(-(double)n, g_sc[0].hpi, x);
}
float oc_cosf(float y)
{
    pthread_once(&g_sc_once, sc_init);
    double x = (double)y;
    uint32_t t = abstop12(y);
    if (t <= 0x3f3) {
        if (t <= 0x397) return 1.0f;
        return (float)sc_cos_poly(x * x, &g_sc[0]);
    }
    int n;
    x = sc_reduce(x, &n);
    const sincos_tab* p = &g_sc[(n >> 1) & 1];
    double x2 = x * x;
    if ((n & 1) == 0) return (float)sc_cos_poly(x2, p);
    return (float)sc_sin_poly(x * g_sc[0].sign[n & 3], x2, p);
}
float oc_sinf(float y)
{
    pthread_once(&g_sc_once, sc_init);
    double x = (double)y;
    uint32_t t = abstop12(y);
    if (t <= 0x3f3) {
        if (t <= 0x397) return y;
        return (float)sc_sin_poly(x, x * x, &g_sc[0]);
    }
    int n;
    x = sc_reduce(x, &n);
    const sincos_tab* p = &g_sc[(n >> 1) & 1];
    double x2 = x * x;
    if ((n & 1) == 0) return (float)sc_sin_poly(x * g_sc[0].sign[n & 3], x2, p);
    return (float)sc_cos_poly(x2, p);
}

/* ------------------------------------------------------------------ cv::resize INTER_LINEAR, CV_8UC1 */
/* OpenCV imgproc/resize.cpp: coefficient tables in float from a double scale, INTER_RESIZE_COEF_BITS=11,
 * HResizeLinear (int accumulators) + VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>. */
static void resize_axis_tab(int ssize, int dsize, int* ofs, short* coef)
{
    double scale = (double)ssize / (double)dsize;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        float c0 = 1.f - f, c1 = f;
        coef[2 * d] = (short)oc_round_f(c0 * 2048.f);
        coef[2 * d + 1] = (short)oc_round_f(c1 * 2048.f);
    }
}
void oc_resize_linear_8u(const uint8_t* src, int sw, int sh, int sstride,
                         uint8_t* dst, int dw, int dh, int dstride)
{
    int* xofs = (int*)malloc(sizeof(int) * (size_t)(dw + dh));
    int* yofs = xofs + dw;
    short* ca = (short*)malloc(sizeof(short) * 2 * (size_t)(dw + dh));
    short* cb = ca + 2 * dw;
    int* rows = (int*)malloc(sizeof(int) * 2 * (size_t)dw);
    resize_axis_tab(sw, dw, xofs, ca);
    resize_axis_tab(sh, dh, yofs, cb);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = yofs[dy], sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const uint8_t* r0 = src + (size_t)sy0 * sstride;
        const uint8_t* r1 = src + (size_t)sy1 * sstride;
        int* S0 = rows; int* S1 = rows + dw;
        for (int dx = 0; dx < dw; dx++) {
            int sx0 = xofs[dx], sx1 = sx0 + 1 < sw ? sx0 + 1 : sw - 1;
            int a0 = ca[2 * dx], a1 = ca[2 * dx + 1];
            S0[dx] = r0[sx0] * a0 + r0[sx1] * a1;
            S1[dx] = r1[sx0] * a0 + r1[sx1] * a1;
        }
        int b0 = cb[2 * dy], b1 = cb[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = (uint8_t)((((b0 * (S0[dx] >> 4)) >> 16) + ((b1 * (S1[dx] >> 4)) >> 16) + 2) >> 2);
    }
    free(rows); free(ca); free(xofs);
}

/* cv::copyMakeBorder(..., BORDER_REFLECT_101 [+ISOLATED]) in place: `whole` is the (w+2b)x(h+2b) buffer whose
 * payload already sits at (b,b). */
static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * (len - 1) - p; }
    return p;
}
void oc_border_reflect101(uint8_t* whole, int w, int h, int stride, int b)
{
    for (int y = 0; y < h; y++) {
        uint8_t* row = whole + (size_t)(y + b) * stride;
        for (int x = 0; x < b; x++) {
            row[x] = row[b + reflect101(x - b, w)];
            row[b + w + x] = row[b + reflect101(w + x, w)];
        }
    }
    for (int y = 0; y < b; y++) {
        memcpy(whole + (size_t)y * stride, whole + (size_t)(b + reflect101(y - b, h)) * stride, (size_t)(w + 2 * b));
        memcpy(whole + (size_t)(b + h + y) * stride, whole + (size_t)(b + reflect101(h + y, h)) * stride, (size_t)(w + 2 * b));
    }
}

/* ------------------------------------------------------------------ cv::GaussianBlur(7x7, 2, 2, REFLECT_101), CV_8U */
/* OpenCV >= 3.4.1 / 4.x takes the fixed-point path (imgproc/smooth.simd.hpp, ufixedpoint16 kernel
 * [18,34,48,56,48,34,18]/256): horizontal sums exact in 16 bits, vertical (sum + 2^15) >> 16. */
void oc_gaussian7x7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
    uint16_t* hb = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint16_t* o = hb + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            if (x >= 3 && x < w - 3) {
                o[x] = (uint16_t)(18 * (s[x - 3] + s[x + 3]) + 34 * (s[x - 2] + s[x + 2]) + 48 * (s[x - 1] + s[x + 1]) + 56 * s[x]);
            } else {
                int acc = 0;
                for (int k = -3; k <= 3; k++) acc += K[k + 3] * s[reflect101(x + k, w)];
                o[x] = (uint16_t)acc;
            }
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* r[7];
        for (int k = -3; k <= 3; k++) r[k + 3] = hb + (size_t)reflect101(y + k, h) * w;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 18u * ((uint32_t)r[0][x] + r[6][x]) + 34u * ((uint32_t)r[1][x] + r[5][x]) +
                           48u * ((uint32_t)r[2][x] + r[4][x]) + 56u * r[3][x];
            acc = (acc + 32768u) >> 16;
            d[x] = (uint8_t)(acc > 255 ? 255 : acc);
        }
    }
    free(hb);
}

/* ------------------------------------------------------------------ cv::cvtColor to gray, CV_8U (Tracking.cc:174-199) */
/* OpenCV 4.x imgproc/color_rgb: fixed point with 15 fractional bits, R 9798, G 19235, B 3735, round to nearest. */
void oc_cvt_gray(const uint8_t* src, int w, int h, int sstride, int channels, int rgb, uint8_t* dst, int dstride)
{
    const int c0 = rgb ? 9798 : 3735, c2 = rgb ? 3735 : 9798;
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++, s += channels)
            d[x] = (uint8_t)((s[0] * c0 + s[1] * 19235 + s[2] * c2 + (1 << 14)) >> 15);
    }
}

/* ------------------------------------------------------------------ cv::FAST TYPE_9_16 (features2d/fast.cpp, fast_score.cpp) */
static const int g_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int g_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* cornerScore<16>(ptr, pixel, threshold): max(threshold, best bright arc, best dark arc) - 1 */
static int corner_score(const uint8_t* p, int stride, int threshold)
{
    int d[25], v = p[0];
    for (int k = 0; k < 25; k++) d[k] = v - p[g_ring_dy[k & 15] * stride + g_ring_dx[k & 15]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        for (int m = 3; m <= 8; m++) if (d[k + m] < a) a = d[k + m];
        int t = a < d[k] ? a : d[k]; if (t > a0) a0 = t;
        t = a < d[k + 9] ? a : d[k + 9]; if (t > a0) a0 = t;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        for (int m = 3; m <= 8; m++) if (d[k + m] > b) b = d[k + m];
        int t = b > d[k] ? b : d[k]; if (t < b0) b0 = t;
        t = b > d[k + 9] ? b : d[k + 9]; if (t < b0) b0 = t;
    }
    return -b0 - 1;
}
int oc_fast_score(const uint8_t* p, int stride) { return corner_score(p, stride, 0); }

/* FAST_t<16> corner test: >= 9 contiguous ring pixels all darker than v-T or all brighter than v+T.
 * Same early-outs as cv::FAST_t: every 9-arc contains one pixel of each opposite pair (k, k+8). */
static inline int ring_class(int x, int lo, int hi) { return (x < lo ? 1 : 0) | (x > hi ? 2 : 0); }
static int is_corner(const uint8_t* p, const int* off, int T)
{
    const int v = p[0], lo = v - T, hi = v + T;
    int d = ring_class(p[off[0]], lo, hi) | ring_class(p[off[8]], lo, hi);
    if (!d) return 0;
    d &= ring_class(p[off[2]], lo, hi) | ring_class(p[off[10]], lo, hi);
    d &= ring_class(p[off[4]], lo, hi) | ring_class(p[off[12]], lo, hi);
    d &= ring_class(p[off[6]], lo, hi) | ring_class(p[off[14]], lo, hi);
    if (!d) return 0;
    d &= ring_class(p[off[1]], lo, hi) | ring_class(p[off[9]], lo, hi);
    d &= ring_class(p[off[3]], lo, hi) | ring_class(p[off[11]], lo, hi);
    d &= ring_class(p[off[5]], lo, hi) | ring_class(p[off[13]], lo, hi);
    d &= ring_class(p[off[7]], lo, hi) | ring_class(p[off[15]], lo, hi);
    if (!d) return 0;
    if (d & 1) {
        int n = 0;
        for (int k = 0; k < 25; k++) { if (p[off[k & 15]] < lo) { if (++n > 8) return 1; } else n = 0; }
    }
    if (d & 2) {
        int n = 0;
        for (int k = 0; k < 25; k++) { if (p[off[k & 15]] > hi) { if (++n > 8) return 1; } else n = 0; }
    }
    return 0;
}

int oc_fast9_16(const uint8_t* roi, int w, int h, int stride, int threshold, int nms, OcKeyPoint* out, int cap)
{
    if (threshold < 0) threshold = 0;
    if (threshold > 255) threshold = 255;
    if (w < 7 || h < 7) return 0;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = g_ring_dy[k] * stride + g_ring_dx[k];
    /* score buffer: 0 = not a corner; bit 8 marks a corner whose score is 0 (possible only when threshold == 0) */
    uint16_t stackbuf[64 * 64];
    const size_t npx = (size_t)w * (size_t)h;
    uint16_t* sc = npx <= 64 * 64 ? stackbuf : (uint16_t*)malloc(npx * sizeof(uint16_t));
    memset(sc, 0, npx * sizeof(uint16_t));
    for (int y = 3; y < h - 3; y++) {
        const uint8_t* row = roi + (size_t)y * stride;
        uint16_t* srow = sc + (size_t)y * w;
        for (int x = 3; x < w - 3; x++)
            if (is_corner(row + x, off, threshold)) srow[x] = (uint16_t)(0x100 | corner_score(row + x, stride, threshold));
    }
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint16_t* c = sc + (size_t)y * w + x;
            if (!c[0]) continue;
            const int s = c[0] & 0xff;
            if (nms) {
                if (!(s > (c[-1] & 0xff) && s > (c[1] & 0xff) && s > (c[-w - 1] & 0xff) && s > (c[-w] & 0xff) &&
                      s > (c[-w + 1] & 0xff) && s > (c[w - 1] & 0xff) && s > (c[w] & 0xff) && s > (c[w + 1] & 0xff))) continue;
            }
            if (n < cap) {
                OcKeyPoint k = {(float)x, (float)y, 7.f, -1.f, nms ? (float)s : 0.f, 0, -1};
                out[n] = k;
            }
            n++;
        }
    if (sc != stackbuf) free(sc);
    return n;
}

/* ------------------------------------------------------------------ IC_Angle / rBRIEF (ORBextractor.cc:77-152) */
float oc_ic_angle(const uint8_t* center, int step, const int* umax)
{
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = center[u + v * step], vm = center[u - v * step];
            v_sum += (vp - vm);
            m_10 += u * (vp + vm);
        }
        m_01 += v * v_sum;
    }
    return oc_fast_atan2((float)m_01, (float)m_10);
}

void oc_orb_descriptor(const uint8_t* center, int step, float angle_deg, uint8_t* desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factorPI;
    float a = oc_cosf(angle), b = oc_sinf(angle);
    const int8_t* pat = g_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
            int t0 = center[oc_round_f((float)x0 * b + (float)y0 * a) * step + oc_round_f((float)x0 * a - (float)y0 * b)];
            int t1 = center[oc_round_f((float)x1 * b + (float)y1 * a) * step + oc_round_f((float)x1 * a - (float)y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ------------------------------------------------------------------ DistributeOctTree (ORBextractor.cc:501-815) */
typedef struct {
    int ULx, ULy, URx, BRy;    /* UL=(ULx,ULy) UR=(URx,ULy) BL=(ULx,BRy) BR=(URx,BRy) */
    int* keys; int nkeys;      /* indices into the input vector, in input order */
    int nomore, prev, next, seq;
} QNode;
typedef struct { QNode* n; int cnt, cap, head, tail, size, seq; } QList;

static int ql_alloc(QList* L)
{
    if (L->cnt == L->cap) { L->cap = L->cap ? L->cap * 2 : 256; L->n = (QNode*)realloc(L->n, sizeof(QNode) * (size_t)L->cap); }
    QNode* q = &L->n[L->cnt];
    memset(q, 0, sizeof(*q));
    q->prev = q->next = -1;
    q->seq = L->seq++;
    return L->cnt++;
}
static void ql_push_front(QList* L, int id)
{
    L->n[id].prev = -1; L->n[id].next = L->head;
    if (L->head >= 0) L->n[L->head].prev = id; else L->tail = id;
    L->head = id; L->size++;
}
static void ql_push_back(QList* L, int id)
{
    L->n[id].next = -1; L->n[id].prev = L->tail;
    if (L->tail >= 0) L->n[L->tail].next = id; else L->head = id;
    L->tail = id; L->size++;
}
static int ql_erase(QList* L, int id) /* returns next */
{
    int p = L->n[id].prev, nx = L->n[id].next;
    if (p >= 0) L->n[p].next = nx; else L->head = nx;
    if (nx >= 0) L->n[nx].prev = p; else L->tail = p;
    free(L->n[id].keys); L->n[id].keys = NULL;
    L->size--;
    return nx;
}
/* ExtractorNode::DivideNode: children ids returned in c[0..3] (n1..n4), not yet linked */
static void q_divide(QList* L, int id, const OcKeyPoint* kp, int c[4])
{
    for (int k = 0; k < 4; k++) c[k] = ql_alloc(L); /* may realloc: take parent pointer afterwards */
    QNode* P = &L->n[id];
    int halfX = (int)ceilf((float)(P->URx - P->ULx) / 2);
    int halfY = (int)ceilf((float)(P->BRy - P->ULy) / 2);
    int midX = P->ULx + halfX, midY = P->ULy + halfY;
    QNode* n1 = &L->n[c[0]]; QNode* n2 = &L->n[c[1]]; QNode* n3 = &L->n[c[2]]; QNode* n4 = &L->n[c[3]];
    n1->ULx = P->ULx; n1->URx = midX;   n1->ULy = P->ULy; n1->BRy = midY;
    n2->ULx = midX;   n2->URx = P->URx; n2->ULy = P->ULy; n2->BRy = midY;
    n3->ULx = P->ULx; n3->URx = midX;   n3->ULy = midY;   n3->BRy = P->BRy;
    n4->ULx = midX;   n4->URx = P->URx; n4->ULy = midY;   n4->BRy = P->BRy;
    for (int k = 0; k < 4; k++) L->n[c[k]].keys = (int*)malloc(sizeof(int) * (size_t)(P->nkeys ? P->nkeys : 1));
    for (int i = 0; i < P->nkeys; i++) {
        const OcKeyPoint* k = &kp[P->keys[i]];
        QNode* t;
        if (k->x < (float)midX) t = (k->y < (float)midY) ? n1 : n3;
        else t = (k->y < (float)midY) ? n2 : n4;
        t->keys[t->nkeys++] = P->keys[i];
    }
    for (int k = 0; k < 4; k++) if (L->n[c[k]].nkeys == 1) L->n[c[k]].nomore = 1;
}
typedef struct { int size, seq, id; } QSizeNode;
static int qsn_cmp(const void* a, const void* b)
{
    const QSizeNode* x = (const QSizeNode*)a; const QSizeNode* y = (const QSizeNode*)b;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq); /* stand-in for the reference's pointer compare */
}
/* push the non-empty children (n1..n4 order) to the list front; record the expandable ones */
static void q_link_children(QList* L, const int c[4], QSizeNode* v, int* nv, int* nToExpand)
{
    for (int k = 0; k < 4; k++) {
        QNode* ch = &L->n[c[k]];
        if (ch->nkeys > 0) {
            /* creation order of the surviving list nodes = push order: re-stamp so dropped children don't count */
            ch->seq = L->seq++;
            ql_push_front(L, c[k]);
            if (ch->nkeys > 1) {
                if (nToExpand) (*nToExpand)++;
                v[*nv].size = ch->nkeys; v[*nv].seq = ch->seq; v[*nv].id = c[k]; (*nv)++;
            }
        } else { free(ch->keys); ch->keys = NULL; }
    }
}

int oc_distribute_octtree(const OcKeyPoint* in, int n, int minX, int maxX, int minY, int maxY,
                          int N, OcKeyPoint* out, int cap)
{
    const int nIni = (int)roundf((float)(maxX - minX) / (float)(maxY - minY));
    if (nIni < 1) return -2; /* reference divides by zero here (portrait level) */
    const float hX = (float)(maxX - minX) / (float)nIni;
    QList L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    int* ini = (int*)malloc(sizeof(int) * (size_t)nIni);
    for (int i = 0; i < nIni; i++) {
        int id = ql_alloc(&L);
        QNode* q = &L.n[id];
        q->ULx = (int)(hX * (float)i); q->URx = (int)(hX * (float)(i + 1));
        q->ULy = 0; q->BRy = maxY - minY;
        q->keys = (int*)malloc(sizeof(int) * (size_t)(n ? n : 1));
        ql_push_back(&L, id);
        ini[i] = id;
    }
    for (int i = 0; i < n; i++) {
        QNode* q = &L.n[ini[(size_t)(in[i].x / hX)]];
        q->keys[q->nkeys++] = i;
    }
    for (int it = L.head; it >= 0;) {
        if (L.n[it].nkeys == 1) { L.n[it].nomore = 1; it = L.n[it].next; }
        else if (L.n[it].nkeys == 0) it = ql_erase(&L, it);
        else it = L.n[it].next;
    }
    int finish = 0;
    QSizeNode* v = (QSizeNode*)malloc(sizeof(QSizeNode) * (size_t)(4 * (n + nIni) + 16));
    QSizeNode* vprev = (QSizeNode*)malloc(sizeof(QSizeNode) * (size_t)(4 * (n + nIni) + 16));
    int nv = 0;
    while (!finish) {
        int prevSize = L.size, nToExpand = 0;
        nv = 0;
        for (int it = L.head; it >= 0;) {
            if (L.n[it].nomore) { it = L.n[it].next; continue; }
            int c[4];
            q_divide(&L, it, in, c);
            q_link_children(&L, c, v, &nv, &nToExpand);
            it = ql_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) finish = 1;
        else if (L.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                int nprev = nv;
                memcpy(vprev, v, sizeof(QSizeNode) * (size_t)nv);
                nv = 0;
                qsort(vprev, (size_t)nprev, sizeof(QSizeNode), qsn_cmp);
                for (int j = nprev - 1; j >= 0; j--) {
                    int c[4];
                    q_divide(&L, vprev[j].id, in, c);
                    q_link_children(&L, c, v, &nv, NULL);
                    ql_erase(&L, vprev[j].id);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
    int m = 0;
    for (int it = L.head; it >= 0; it = L.n[it].next) {
        const QNode* q = &L.n[it];
        int best = q->keys[0];
        float maxR = in[best].response;
        for (int k = 1; k < q->nkeys; k++)
            if (in[q->keys[k]].response > maxR) { best = q->keys[k]; maxR = in[best].response; }
        if (m < cap) out[m] = in[best];
        m++;
    }
    for (int i = 0; i < L.cnt; i++) free(L.n[i].keys);
    free(L.n); free(ini); free(v); free(vprev);
    return m;
}

/* ------------------------------------------------------------------ Hamming (ORBmatcher.cc:1844-1860, :84-126) */
int oc_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb; memcpy(&pa, a + 4 * i, 4); memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24);
    }
    return dist;
}
static inline int dist_fast(const uint64_t* a, const uint64_t* b)
{
    return __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
           __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
}
typedef struct { const uint8_t* q; const uint8_t* t; int q0, q1, nt; int32_t *idx1, *d1, *d2; } Top2Job;
static void* top2_worker(void* arg)
{
    Top2Job* J = (Top2Job*)arg;
    for (int i = J->q0; i < J->q1; i++) {
        uint64_t qa[4]; memcpy(qa, J->q + 32 * (size_t)i, 32);
        int best = 256, second = 256, bi = -1;
        for (int j = 0; j < J->nt; j++) {
            uint64_t tb[4]; memcpy(tb, J->t + 32 * (size_t)j, 32);
            int d = dist_fast(qa, tb); /* == oc_descriptor_distance (tests check) */
            if (d < best) { second = best; best = d; bi = j; }
            else if (d < second) second = d;
        }
        J->idx1[i] = bi; J->d1[i] = best; J->d2[i] = second;
    }
    return NULL;
}
void oc_hamming_top2(const uint8_t* q, int nq, const uint8_t* t, int nt,
                     int32_t* idx1, int32_t* d1, int32_t* d2, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256]; Top2Job jobs[256];
    int per = (nq + nthreads - 1) / nthreads;
    for (int k = 0; k < nthreads; k++) {
        int q0 = k * per, q1 = q0 + per > nq ? nq : q0 + per;
        if (q0 > nq) q0 = nq;
        Top2Job j = {q, t, q0, q1, nt, idx1, d1, d2};
        jobs[k] = j;
        if (nthreads == 1) top2_worker(&jobs[k]); else pthread_create(&th[k], NULL, top2_worker, &jobs[k]);
    }
    if (nthreads > 1) for (int k = 0; k < nthreads; k++) pthread_join(th[k], NULL);
}

/* Frame::ComputeStereoMatches, Hamming stage (Frame.cc:554-663): row-band table of right keypoints,
 * octave +-1 and disparity gates, best distance with init TH_HIGH=100, strict <. Output per left keypoint:
 * best right index (or -1 when no candidate beat 100) and its distance; callers accept dist < 75. */
void oc_stereo_hamming(const OcKeyPoint* kl, const uint8_t* dl, int nl,
                       const OcKeyPoint* kr, const uint8_t* dr, int nr,
                       int rows, const float* sf, float minD, float maxD,
                       int32_t* best_idx_r, int32_t* best_dist)
{
    /* vRowIndices as CSR; filling in ascending iR keeps each row's push_back order (:564-590) */
    int* start = (int*)calloc((size_t)rows + 1, sizeof(int));
    for (int pass = 0; pass < 2; pass++) {
        int* fill = pass ? (int*)calloc((size_t)rows, sizeof(int)) : NULL;
        static int dummy;
        int* tab = pass ? (int*)malloc(sizeof(int) * (size_t)(start[rows] ? start[rows] : 1)) : &dummy;
        for (int iR = 0; iR < nr; iR++) {
            float kpY = kr[iR].y;
            float r = 2.0f * sf[kr[iR].octave];
            int maxr = (int)ceilf(kpY + r), minr = (int)floorf(kpY - r);
            for (int yi = minr; yi <= maxr; yi++) {
                if (yi < 0 || yi >= rows) continue; /* the reference would index out of bounds; never happens for extractor output */
                if (!pass) start[yi + 1]++; else tab[start[yi] + fill[yi]++] = iR;
            }
        }
        if (!pass) { for (int y = 0; y < rows; y++) start[y + 1] += start[y]; continue; }
        for (int iL = 0; iL < nl; iL++) {
            best_idx_r[iL] = -1; best_dist[iL] = 100; /* ORBmatcher::TH_HIGH */
            const int levelL = kl[iL].octave;
            const float vL = kl[iL].y, uL = kl[iL].x;
            const int row = (int)vL;
            if (row < 0 || row >= rows || start[row + 1] == start[row]) continue;
            const float minU = uL - maxD, maxU = uL - minD;
            if (maxU < 0) continue;
            int bestDist = 100, bestIdx = -1;
            for (int c = start[row]; c < start[row + 1]; c++) {
                const int iR = tab[c];
                if (kr[iR].octave < levelL - 1 || kr[iR].octave > levelL + 1) continue;
                const float uR = kr[iR].x;
                if (uR >= minU && uR <= maxU) {
                    int d = oc_descriptor_distance(dl + 32 * (size_t)iL, dr + 32 * (size_t)iR);
                    if (d < bestDist) { bestDist = d; bestIdx = iR; }
                }
            }
            best_idx_r[iL] = bestIdx; best_dist[iL] = bestDist;
        }
        free(tab); free(fill);
    }
    free(start);
}

/* ------------------------------------------------------------------ extractor object */
#define OC_MAX_LEVELS 32
struct OcExtractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;                      /* ORBextractor.h: `double scaleFactor` */
    float sf[OC_MAX_LEVELS], inv_sf[OC_MAX_LEVELS], sigma2[OC_MAX_LEVELS], inv_sigma2[OC_MAX_LEVELS];
    int per_level[OC_MAX_LEVELS], umax[HALF_PATCH_SIZE + 1];
    /* per-call state */
    uint8_t* whole[OC_MAX_LEVELS]; int lw[OC_MAX_LEVELS], lh[OC_MAX_LEVELS];
    uint8_t* blur[OC_MAX_LEVELS];
    OcKeyPoint* cand[OC_MAX_LEVELS]; int ncand[OC_MAX_LEVELS];
    int nkp[OC_MAX_LEVELS];
};

OcExtractor* oc_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh)
{
    if (nlevels < 1 || nlevels > OC_MAX_LEVELS) return NULL;
    OcExtractor* e = (OcExtractor*)calloc(1, sizeof(*e));
    e->nfeatures = nfeatures; e->scaleFactor = scaleFactor; e->nlevels = nlevels; e->iniTh = iniTh; e->minTh = minTh;
    e->sf[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->sf[i] = (float)(e->sf[i - 1] * e->scaleFactor);
        e->sigma2[i] = e->sf[i] * e->sf[i];
    }
    for (int i = 0; i < nlevels; i++) { e->inv_sf[i] = 1.0f / e->sf[i]; e->inv_sigma2[i] = 1.0f / e->sigma2[i]; }
    float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->per_level[l] = oc_round_f(nDesired);
        sum += e->per_level[l];
        nDesired *= factor;
    }
    e->per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    int v, v0, vmax = (int)floor(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceil(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = oc_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}
static void oc_release_frame(OcExtractor* e)
{
    for (int l = 0; l < OC_MAX_LEVELS; l++) {
        free(e->whole[l]); e->whole[l] = NULL;
        free(e->blur[l]); e->blur[l] = NULL;
        free(e->cand[l]); e->cand[l] = NULL;
        e->ncand[l] = 0; e->nkp[l] = 0;
    }
}
void oc_destroy(OcExtractor* e) { if (e) { oc_release_frame(e); free(e); } }
int oc_levels(const OcExtractor* e) { return e->nlevels; }
void oc_tables(const OcExtractor* e, float* sf, float* inv_sf, float* s2, float* inv_s2, int* fpl, int* umax16)
{
    for (int i = 0; i < e->nlevels; i++) {
        if (sf) sf[i] = e->sf[i];
        if (inv_sf) inv_sf[i] = e->inv_sf[i];
        if (s2) s2[i] = e->sigma2[i];
        if (inv_s2) inv_s2[i] = e->inv_sigma2[i];
        if (fpl) fpl[i] = e->per_level[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}
int oc_level_size(const OcExtractor* e, int l, int* w, int* h, int* stride)
{
    if (l < 0 || l >= e->nlevels || !e->whole[l]) return -1;
    *w = e->lw[l]; *h = e->lh[l]; *stride = e->lw[l] + 2 * EDGE_THRESHOLD;
    return 0;
}
const uint8_t* oc_level_ptr(const OcExtractor* e, int l)
{
    return e->whole[l] ? e->whole[l] + (size_t)EDGE_THRESHOLD * (e->lw[l] + 2 * EDGE_THRESHOLD) + EDGE_THRESHOLD : NULL;
}
const uint8_t* oc_level_blur_ptr(const OcExtractor* e, int l) { return e->blur[l]; }
int oc_level_candidates(const OcExtractor* e, int l, OcKeyPoint* out, int cap)
{
    int n = e->ncand[l] < cap ? e->ncand[l] : cap;
    if (out && n > 0) memcpy(out, e->cand[l], sizeof(OcKeyPoint) * (size_t)n);
    return e->ncand[l];
}
int oc_level_nkeypoints(const OcExtractor* e, int l) { return e->nkp[l]; }

int oc_extract(OcExtractor* e, const uint8_t* img, int w, int h, int stride,
               OcKeyPoint* kps, int cap, uint8_t* desc)
{
    if (!img || w <= 0 || h <= 0) return 0; /* :1141 empty image -> silent return */
    oc_release_frame(e);
    const int B = EDGE_THRESHOLD;
    /* ---- ComputePyramid :1215-1250 ---- */
    for (int l = 0; l < e->nlevels; l++) {
        float scale = e->inv_sf[l];
        int lw = oc_round_f((float)w * scale), lh = oc_round_f((float)h * scale);
        int ws = lw + 2 * B;
        e->lw[l] = lw; e->lh[l] = lh;
        e->whole[l] = (uint8_t*)malloc((size_t)ws * (size_t)(lh + 2 * B));
        uint8_t* pay = e->whole[l] + (size_t)B * ws + B;
        if (l != 0) {
            const uint8_t* prev = e->whole[l - 1] + (size_t)B * (e->lw[l - 1] + 2 * B) + B;
            oc_resize_linear_8u(prev, e->lw[l - 1], e->lh[l - 1], e->lw[l - 1] + 2 * B, pay, lw, lh, ws);
        } else {
            for (int y = 0; y < h; y++) memcpy(pay + (size_t)y * ws, img + (size_t)y * stride, (size_t)w);
        }
        oc_border_reflect101(e->whole[l], lw, lh, ws, B);
    }
    /* ---- ComputeKeyPointsOctTree :818-946 ---- */
    OcKeyPoint** lvl = (OcKeyPoint**)calloc((size_t)e->nlevels, sizeof(OcKeyPoint*));
    int status = 0;
    const float W = 30;
    for (int l = 0; l < e->nlevels && status == 0; l++) {
        const int ws = e->lw[l] + 2 * B;
        const uint8_t* pay = e->whole[l] + (size_t)B * ws + B;
        const int minBorderX = B - 3, minBorderY = minBorderX;
        const int maxBorderX = e->lw[l] - B + 3, maxBorderY = e->lh[l] - B + 3;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) { status = -2; break; } /* reference divides by zero */
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        int capc = 1024, nc = 0;
        OcKeyPoint* cand = (OcKeyPoint*)malloc(sizeof(OcKeyPoint) * (size_t)capc);
        OcKeyPoint* cell = (OcKeyPoint*)malloc(sizeof(OcKeyPoint) * (size_t)((wCell + 6) * (hCell + 6)));
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                const int y0 = (int)iniY, y1 = (int)maxY, x0 = (int)iniX, x1 = (int)maxX;
                const uint8_t* roi = pay + (size_t)y0 * ws + x0;
                int n = oc_fast9_16(roi, x1 - x0, y1 - y0, ws, e->iniTh, 1, cell, (wCell + 6) * (hCell + 6));
                if (n == 0) n = oc_fast9_16(roi, x1 - x0, y1 - y0, ws, e->minTh, 1, cell, (wCell + 6) * (hCell + 6));
                for (int k = 0; k < n; k++) {
                    cell[k].x += j * wCell; cell[k].y += i * hCell;
                    if (nc == capc) { capc *= 2; cand = (OcKeyPoint*)realloc(cand, sizeof(OcKeyPoint) * (size_t)capc); }
                    cand[nc++] = cell[k];
                }
            }
        }
        free(cell);
        e->cand[l] = cand; e->ncand[l] = nc;
        int capl = nc > 0 ? nc : 1;
        lvl[l] = (OcKeyPoint*)malloc(sizeof(OcKeyPoint) * (size_t)capl);
        int m = oc_distribute_octtree(cand, nc, minBorderX, maxBorderX, minBorderY, maxBorderY, e->per_level[l], lvl[l], capl);
        if (m < 0) { status = m; break; }
        e->nkp[l] = m;
        const int scaledPatchSize = (int)(PATCH_SIZE * e->sf[l]);
        for (int k = 0; k < m; k++) {
            lvl[l][k].x += minBorderX; lvl[l][k].y += minBorderY;
            lvl[l][k].octave = l; lvl[l][k].size = (float)scaledPatchSize;
        }
    }
    if (status == 0)
        for (int l = 0; l < e->nlevels; l++) { /* computeOrientation :492-499 on the un-blurred level */
            const int ws = e->lw[l] + 2 * B;
            const uint8_t* pay = e->whole[l] + (size_t)B * ws + B;
            for (int k = 0; k < e->nkp[l]; k++)
                lvl[l][k].angle = oc_ic_angle(pay + (size_t)oc_round_f(lvl[l][k].y) * ws + oc_round_f(lvl[l][k].x), ws, e->umax);
        }
    /* ---- operator() :1160-1210 ---- */
    int total = 0;
    if (status == 0) {
        for (int l = 0; l < e->nlevels; l++) total += e->nkp[l];
        if (total > cap) status = -1;
    }
    if (status == 0) {
        int offset = 0;
        for (int l = 0; l < e->nlevels; l++) {
            int n = e->nkp[l];
            if (n == 0) continue;
            const int ws = e->lw[l] + 2 * B;
            const uint8_t* pay = e->whole[l] + (size_t)B * ws + B;
            e->blur[l] = (uint8_t*)malloc((size_t)e->lw[l] * (size_t)e->lh[l]);
            oc_gaussian7x7_s2(pay, e->lw[l], e->lh[l], ws, e->blur[l], e->lw[l]);
            for (int k = 0; k < n; k++) {
                OcKeyPoint* kp = &lvl[l][k];
                const uint8_t* c = e->blur[l] + (size_t)oc_round_f(kp->y) * e->lw[l] + oc_round_f(kp->x);
                oc_orb_descriptor(c, e->lw[l], kp->angle, desc + 32 * (size_t)(offset + k));
            }
            if (l != 0) {
                float scale = e->sf[l];
                for (int k = 0; k < n; k++) { lvl[l][k].x *= scale; lvl[l][k].y *= scale; }
            }
            memcpy(kps + offset, lvl[l], sizeof(OcKeyPoint) * (size_t)n);
            offset += n;
        }
    }
    for (int l = 0; l < e->nlevels; l++) free(lvl[l]);
    free(lvl);
    return status == 0 ? total : status;
}

/* ------------------------------------------------------------------ Frame::ComputeStereoMatches (Frame.cc:547-788) */
typedef struct { int dist, idx; } DistIdx;
static int distidx_cmp(const void* a, const void* b)
{
    const DistIdx* x = (const DistIdx*)a; const DistIdx* y = (const DistIdx*)b;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
void oc_stereo_match(const OcExtractor* EL, const OcExtractor* ER,
                     const OcKeyPoint* kl, const uint8_t* dl, int nl,
                     const OcKeyPoint* kr, const uint8_t* dr, int nr,
                     float mbf, float fx, float* u_right, float* depth)
{
    const int B = EDGE_THRESHOLD;
    for (int i = 0; i < nl; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (100 + 50) / 2;                  /* (TH_HIGH+TH_LOW)/2 */
    const int nRows = EL->lh[0];
    const float mb = mbf / fx;                             /* Frame.cc:120 */
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    int32_t* bidx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nl > 0 ? nl : 1));
    int32_t* bdist = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nl > 0 ? nl : 1));
    oc_stereo_hamming(kl, dl, nl, kr, dr, nr, nRows, EL->sf, minD, maxD, bidx, bdist);   /* :554-663 */
    DistIdx* v = (DistIdx*)malloc(sizeof(DistIdx) * (size_t)(nl > 0 ? nl : 1));
    int nv = 0;
    for (int iL = 0; iL < nl; iL++) {
        if (bidx[iL] < 0 || !(bdist[iL] < thOrbDist)) continue;
        const OcKeyPoint* kpL = &kl[iL];
        const float uL = kpL->x;
        const float uR0 = kr[bidx[iL]].x;
        const float scaleFactor = EL->inv_sf[kpL->octave];
        const float scaleduL = roundf(kpL->x * scaleFactor);
        const float scaledvL = roundf(kpL->y * scaleFactor);
        const float scaleduR0 = roundf(uR0 * scaleFactor);
        const int w = 5, L = 5;
        const int lv = kpL->octave;
        const int wsL = EL->lw[lv] + 2 * B, wsR = ER->lw[lv] + 2 * B;
        const uint8_t* pL = EL->whole[lv] + (size_t)B * wsL + B;     /* payload origins; the apron makes small */
        const uint8_t* pR = ER->whole[lv] + (size_t)B * wsR + B;     /* excursions valid memory, as in the reference */
        const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
        const float iniu = scaleduR0 + L - w;
        const float endu = scaleduR0 + L + w + 1;
        if (iniu < 0 || endu >= ER->lw[lv]) continue;
        const int cL = pL[(size_t)cv * wsL + cu];
        int bestDist = 2147483647, bestincR = 0;
        float vDists[11];
        for (int incR = -L; incR <= L; incR++) {
            const int cR = pR[(size_t)cv * wsR + cr + incR];
            int sad = 0;
            for (int dy = -w; dy <= w; dy++)
                for (int dx = -w; dx <= w; dx++) {
                    const int a = pL[(size_t)(cv + dy) * wsL + cu + dx] - cL;
                    const int b = pR[(size_t)(cv + dy) * wsR + cr + incR + dx] - cR;
                    sad += a > b ? a - b : b - a;
                }
            const float dist = (float)sad;                 /* cv::norm(IL,IR,NORM_L1) of integer-valued floats */
            if (dist < (float)bestDist) { bestDist = (int)dist; bestincR = incR; }
            vDists[L + incR] = dist;
        }
        if (bestincR == -L || bestincR == L) continue;
        const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = EL->sf[lv] * ((float)scaleduR0 + (float)bestincR + deltaR);
        float disparity = (uL - bestuR);
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
            depth[iL] = mbf / disparity;
            u_right[iL] = bestuR;
            v[nv].dist = bestDist; v[nv].idx = iL; nv++;
        }
    }
    if (nv > 0) {                                          /* the reference indexes an empty vector here */
        qsort(v, (size_t)nv, sizeof(DistIdx), distidx_cmp);
        const float median = (float)v[nv / 2].dist;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nv - 1; i >= 0; i--) {
            if ((float)v[i].dist < thDist) break;
            u_right[v[i].idx] = -1; depth[v[i].idx] = -1;
        }
    }
    free(v); free(bidx); free(bdist);
}

/* ------------------------------------------------------------------ windowed top-2 (Frame.cc:254-271,388-460; ORBmatcher.cc:46-142) */
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
void oc_window_top2(const OcKeyPoint* kps, const uint8_t* desc, int n, const uint8_t* occupied, const float* u_right,
                    float mnMinX, float mnMinY, float invW, float invH,
                    const OcWindowQuery* q, const uint8_t* qdesc, int nq,
                    int32_t* best_idx, int32_t* best_dist, int32_t* best_level, int32_t* best_dist2, int32_t* best_level2)
{
    /* AssignFeaturesToGrid: mGrid[posX][posY] holds keypoint indices in ascending order */
    int* cnt = (int*)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, sizeof(int));
    int* cell = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        int posX = (int)roundf((kps[i].x - mnMinX) * invW), posY = (int)roundf((kps[i].y - mnMinY) * invH);
        cell[i] = (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) ? -1 : posX * FRAME_GRID_ROWS + posY;
        if (cell[i] >= 0) cnt[cell[i] + 1]++;
    }
    for (int c = 0; c < FRAME_GRID_COLS * FRAME_GRID_ROWS; c++) cnt[c + 1] += cnt[c];
    int* tab = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* fill = (int*)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS, sizeof(int));
    for (int i = 0; i < n; i++) if (cell[i] >= 0) tab[cnt[cell[i]] + fill[cell[i]]++] = i;
    for (int k = 0; k < nq; k++) {
        const float x = q[k].x, y = q[k].y, r = q[k].r;
        const int minLevel = q[k].min_level, maxLevel = q[k].max_level;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        /* GetFeaturesInArea */
        int nMinCellX = (int)floorf((x - mnMinX - r) * invW); if (nMinCellX < 0) nMinCellX = 0;
        int nMaxCellX = (int)ceilf((x - mnMinX + r) * invW); if (nMaxCellX > FRAME_GRID_COLS - 1) nMaxCellX = FRAME_GRID_COLS - 1;
        int nMinCellY = (int)floorf((y - mnMinY - r) * invH); if (nMinCellY < 0) nMinCellY = 0;
        int nMaxCellY = (int)ceilf((y - mnMinY + r) * invH); if (nMaxCellY > FRAME_GRID_ROWS - 1) nMaxCellY = FRAME_GRID_ROWS - 1;
        const int empty = nMinCellX >= FRAME_GRID_COLS || nMaxCellX < 0 || nMinCellY >= FRAME_GRID_ROWS || nMaxCellY < 0;
        const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        if (!empty)
            for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
                for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                    const int c = ix * FRAME_GRID_ROWS + iy;
                    for (int j = cnt[c]; j < cnt[c + 1]; j++) {
                        const int idx = tab[j];
                        const OcKeyPoint* kp = &kps[idx];
                        if (bCheckLevels) {
                            if (kp->octave < minLevel) continue;
                            if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                        }
                        const float distx = kp->x - x, disty = kp->y - y;
                        if (!(fabsf(distx) < r && fabsf(disty) < r)) continue;
                        /* the SearchByProjection loop body (:86-126) */
                        if (occupied && occupied[idx]) continue;
                        if (u_right && u_right[idx] > 0) {
                            const float er = fabsf(q[k].xr - u_right[idx]);
                            if (er > r) continue;
                        }
                        const int dist = oc_descriptor_distance(qdesc + 32 * (size_t)k, desc + 32 * (size_t)idx);
                        if (dist < bestDist) {
                            bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kp->octave; bestIdx = idx;
                        } else if (dist < bestDist2) {
                            bestLevel2 = kp->octave; bestDist2 = dist;
                        }
                    }
                }
        best_idx[k] = bestIdx; best_dist[k] = bestDist; best_level[k] = bestLevel; best_dist2[k] = bestDist2; best_level2[k] = bestLevel2;
    }
    free(cnt); free(cell); free(tab); free(fill);
}

/* ------------------------------------------------------------------ cv::remap INTER_LINEAR, CV_8UC1, BORDER_CONSTANT(0)
 * (the rectification of Examples/Stereo/stereo_euroc.cc:136-137; maps from cv::initUndistortRectifyMap(.., CV_32F, ..), :97-98).
 * OpenCV 4.x imgproc/imgwarp.cpp: the CV_32FC1 map pair is converted to fixed point with INTER_BITS = 5,
 *   sx = cvRound(map1 * 32), sy = cvRound(map2 * 32)      (f32 product, round half even)
 *   (ix, iy) = saturate_cast<short>(sx >> 5, sy >> 5),   a = sx & 31, b = sy & 31
 * and remapBilinear<FixedPtCast<int, uchar, 15>> blends with the short table w = {(32-a)(32-b), a(32-b), (32-a)b, ab} * 32
 * (exact: the float table entries are multiples of 2^-10, so the table needs no sum correction):
 *   dst = (p00*w00 + p01*w01 + p10*w10 + p11*w11 + 2^14) >> 15,   taps outside the source read the constant 0.
 * Pinned bit-exact against cv2 4.13 (tests/golden/prims2_cv2.npz). */
static int sat_short(int v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; }
void oc_remap_linear_8u(const uint8_t* src, int sw, int sh, int sstride, const float* map1, const float* map2,
                        int map_stride, int dw, int dh, uint8_t* dst, int dstride)
{
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            const int sx = oc_round_f(map1[(size_t)y * map_stride + x] * 32.0f);
            const int sy = oc_round_f(map2[(size_t)y * map_stride + x] * 32.0f);
            const int ix = sat_short(sx >> 5), iy = sat_short(sy >> 5), a = sx & 31, b = sy & 31;
            int p[4];
            for (int k = 0; k < 4; k++) {
                const int xx = ix + (k & 1), yy = iy + (k >> 1);
                p[k] = (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstride + xx] : 0;
            }
            const int v = p[0] * ((32 - a) * (32 - b) * 32) + p[1] * (a * (32 - b) * 32) +
                          p[2] * ((32 - a) * b * 32) + p[3] * (a * b * 32);
            dst[(size_t)y * dstride + x] = (uint8_t)((v + (1 << 14)) >> 15);
        }
}

/* ------------------------------------------------------------------ cv::undistortPoints(src, dst, K, D, Mat(), K)
 * as called by Frame::UndistortKeyPoints (Frame.cc:471-506) and Frame::ComputeImageBounds (:508-538).
 * OpenCV 4.x calib3d/undistort: double arithmetic, 5 fixed-point iterations (default TermCriteria(MAX_ITER, 5, 0.01)),
 * identity R, P = K, no tilt; k = (k1, k2, p1, p2, k3) with k3 = 0 when only four coefficients are given.
 * K4 = (fx, fy, cx, cy) as f32 (mK is CV_32F, Frame.cc:91-103). Pinned bit-exact against cv2 4.13. */
void oc_undistort_points(const float* xy, int n, const float* K4, const float* dist, int ndist, float* out_xy)
{
    double k[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < ndist && i < 5; i++) k[i] = (double)dist[i];
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3], ifx = 1.0 / fx, ify = 1.0 / fy;
    for (int i = 0; i < n; i++) {
        const double u = xy[2 * i], v = xy[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((0.0 * r2 + 0.0) * r2 + 0.0) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        out_xy[2 * i] = (float)(fx * x + cx);
        out_xy[2 * i + 1] = (float)(fy * y + cy);
    }
}

/* ------------------------------------------------------------------ cv::initUndistortRectifyMap(K, D, R, P, size, CV_32F, M1, M2)
 * as called once per camera by Examples/Stereo/stereo_euroc.cc:96-97 (the maps cv::remap then uses at :136-137).
 * OpenCV 4.x calib3d/undistort: everything in double; iR = (P[:, :3] * R)^-1 with cv::Matx's closed 3x3 cofactor inverse;
 * for the pixel (j, i): (_x, _y, _w) = iR * (j, i, 1) evaluated as (i*ir[1] + ir[2]) + j*ir[0] (the per-lane positions of the
 * vectorised loop — the scalar tail's running sum `_x += ir[0]` differs in the last bit and is NOT what cv2 4.13 returns),
 * then the Brown-Conrady model with k1..k6, p1, p2, s1..s4 (no tilt) and the projection with K. D has 4, 5, 8 or 12 entries.
 * Pinned bit-exact against cv2 4.13 (tests/golden/prims3_cv2.npz). K9 / R9 row-major 3x3, Ar9 = the left 3x3 block of P. */
void oc_init_undistort_rectify_map(const double* K9, const double* D, int nD, const double* R9, const double* Ar9, int w, int h,
                                   float* map1, float* map2)
{
    double k[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nD && i < 12; i++) k[i] = D[i];
    const double k1 = k[0], k2 = k[1], p1 = k[2], p2 = k[3], k3 = k[4], k4 = k[5], k5 = k[6], k6 = k[7], s1 = k[8], s2 = k[9], s3 = k[10], s4 = k[11];
    const double fx = K9[0], fy = K9[4], u0 = K9[2], v0 = K9[5];
    double a[9], ir[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int q = 0; q < 3; q++) s = s + Ar9[3 * r + q] * R9[3 * q + c];
            a[3 * r + c] = s;
        }
    {
        double d = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
        d = 1.0 / d;
        ir[0] = (a[4] * a[8] - a[5] * a[7]) * d; ir[1] = (a[2] * a[7] - a[1] * a[8]) * d; ir[2] = (a[1] * a[5] - a[2] * a[4]) * d;
        ir[3] = (a[5] * a[6] - a[3] * a[8]) * d; ir[4] = (a[0] * a[8] - a[2] * a[6]) * d; ir[5] = (a[2] * a[3] - a[0] * a[5]) * d;
        ir[6] = (a[3] * a[7] - a[4] * a[6]) * d; ir[7] = (a[1] * a[6] - a[0] * a[7]) * d; ir[8] = (a[0] * a[4] - a[1] * a[3]) * d;
    }
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const double _x = ((double)i * ir[1] + ir[2]) + (double)j * ir[0];
            const double _y = ((double)i * ir[4] + ir[5]) + (double)j * ir[3];
            const double _w = ((double)i * ir[7] + ir[8]) + (double)j * ir[6];
            const double iw = 1.0 / _w, x = _x * iw, y = _y * iw;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
            const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2;
            const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2;
            map1[(size_t)i * w + j] = (float)(fx * xd + u0);
            map2[(size_t)i * w + j] = (float)(fy * yd + v0);
        }
}

/* ================================================================== bag of words (DBoW2)
 * Frame::ComputeBoW (Frame.cc:462-469) = mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) with
 * ORBVocabulary = TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h). The reference vendors only
 * DBoW2's HEADERS (the .h files of Thirdparty/DBoW2/DBoW2): TemplatedVocabulary.h holds transform (:1127-1196, :1216-1260) and the
 * text loader that defines node / word numbering (:1338-1420); FORB.cpp, BowVector.cpp, FeatureVector.cpp and
 * ScoringObject.cpp are ABSENT from the snapshot, so FORB::distance (the same 32-bit SWAR Hamming as
 * ORBmatcher::DescriptorDistance), BowVector::addWeight / addIfNotExist / normalize, FeatureVector::addFeature and
 * L1Scoring::score are restated from DBoW2's published algorithm (the un-versioned copy ORB-SLAM2 ships).
 * Pinned to the reference's verbatim TemplatedVocabulary.h header (oracle/bow_glue.cc); the leaf functions of DBoW2's absent
 * .cpp files follow DBoW2's published algorithm. */
struct OcVocabulary {
    int k, L, scoring, weighting, n;      /* n = nodes incl. root (node 0) */
    int32_t* child_off; int32_t* child;   /* CSR children lists in push_back (= ascending id) order */
    uint8_t* desc; double* weight; int32_t* word; int nwords;
};

OcVocabulary* oc_vocab_create(int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                              const uint8_t* is_leaf, const uint8_t* desc, const double* weight)
{
    if (n_nodes < 0 || n_nodes > (1 << 28)) return NULL;
    OcVocabulary* v = (OcVocabulary*)calloc(1, sizeof *v);
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->n = n_nodes + 1;
    v->child_off = (int32_t*)calloc((size_t)v->n + 1, sizeof(int32_t));
    v->child = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_nodes > 0 ? n_nodes : 1));
    v->desc = (uint8_t*)calloc((size_t)v->n, 32);
    v->weight = (double*)calloc((size_t)v->n, sizeof(double));
    v->word = (int32_t*)malloc(sizeof(int32_t) * (size_t)v->n);
    v->word[0] = -1;
    for (int i = 0; i < n_nodes; i++) {                 /* node id = line number (TemplatedVocabulary.h:1385-1392) */
        if (parent[i] < 0 || parent[i] > i) { oc_vocab_destroy(v); return NULL; }
        v->child_off[parent[i] + 1]++;
        memcpy(v->desc + 32 * (size_t)(i + 1), desc + 32 * (size_t)i, 32);
        v->weight[i + 1] = weight[i];
        v->word[i + 1] = is_leaf[i] ? v->nwords++ : -1; /* :1407-1414 */
    }
    for (int i = 0; i < v->n; i++) v->child_off[i + 1] += v->child_off[i];
    int32_t* fill = (int32_t*)calloc((size_t)n_nodes + 1, sizeof(int32_t));
    for (int i = 0; i < n_nodes; i++) v->child[v->child_off[parent[i]] + fill[parent[i]]++] = i + 1;
    free(fill);
    return v;
}
void oc_vocab_destroy(OcVocabulary* v)
{
    if (!v) return;
    free(v->child_off); free(v->child); free(v->desc); free(v->weight); free(v->word); free(v);
}
int oc_vocab_words(const OcVocabulary* v) { return v->nwords; }

/* transform(feature, word_id, weight, nid, levelsup), TemplatedVocabulary.h:1216-1260 */
static void vocab_descend(const OcVocabulary* v, const uint8_t* f, int levelsup, int32_t* word, double* w, int32_t* nid)
{
    const int nid_level = v->L - levelsup;
    if (nid_level <= 0) *nid = 0;
    int final_id = 0, current_level = 0;
    do {
        ++current_level;
        const int32_t* ch = v->child + v->child_off[final_id];
        const int nch = v->child_off[final_id + 1] - v->child_off[final_id];
        final_id = ch[0];
        double best_d = oc_descriptor_distance(f, v->desc + 32 * (size_t)final_id);
        for (int j = 1; j < nch; j++) {
            const double d = oc_descriptor_distance(f, v->desc + 32 * (size_t)ch[j]);
            if (d < best_d) { best_d = d; final_id = ch[j]; }
        }
        if (current_level == nid_level) *nid = final_id;
    } while (v->child_off[final_id + 1] != v->child_off[final_id]);
    *word = v->word[final_id]; *w = v->weight[final_id];
}

/* transform(features, v, fv, levelsup), TemplatedVocabulary.h:1127-1196. Outputs: per feature word / node (always
 * written); the BowVector as (id ascending, value); the FeatureVector as CSR (node ascending, features in push order). */
void oc_vocab_transform(const OcVocabulary* v, const uint8_t* desc, int n, int levelsup,
                        int32_t* word, int32_t* node, int32_t* bow_id, double* bow_val, int32_t* n_bow,
                        int32_t* fv_node, int32_t* fv_off, int32_t* fv_feat, int32_t* n_fv)
{
    *n_bow = 0; *n_fv = 0; fv_off[0] = 0;
    if (v->n <= 1 || n <= 0) return;
    double* acc = (double*)calloc((size_t)(v->nwords > 0 ? v->nwords : 1), sizeof(double));
    uint8_t* present = (uint8_t*)calloc((size_t)(v->nwords > 0 ? v->nwords : 1), 1);
    double* wts = (double*)malloc(sizeof(double) * (size_t)n);
    int32_t* cnt = (int32_t*)calloc((size_t)v->n + 1, sizeof(int32_t));
    const int tf = v->weighting == 0 || v->weighting == 1;            /* TF_IDF, TF: addWeight; IDF, BINARY: addIfNotExist */
    for (int i = 0; i < n; i++) {
        vocab_descend(v, desc + 32 * (size_t)i, levelsup, &word[i], &wts[i], &node[i]);
        if (wts[i] > 0) {
            if (!present[word[i]]) { present[word[i]] = 1; acc[word[i]] = wts[i]; }
            else if (tf) acc[word[i]] += wts[i];
            cnt[node[i] + 1]++;
        }
    }
    int nb = 0;
    for (int id = 0; id < v->nwords; id++) if (present[id]) { bow_id[nb] = id; bow_val[nb] = acc[id]; nb++; }
    const int must = v->scoring != 5;                                  /* DOT_PRODUCT does not normalise */
    if (tf && nb > 0 && !must) { const double nd = nb; for (int j = 0; j < nb; j++) bow_val[j] /= nd; }
    if (must) {                                                        /* BowVector::normalize */
        double norm = 0.0;
        if (v->scoring != 1) { for (int j = 0; j < nb; j++) norm += fabs(bow_val[j]); }
        else { for (int j = 0; j < nb; j++) norm += bow_val[j] * bow_val[j]; norm = sqrt(norm); }
        if (norm > 0.0) for (int j = 0; j < nb; j++) bow_val[j] /= norm;
    }
    *n_bow = nb;
    int nf = 0;
    int32_t* start = (int32_t*)malloc(sizeof(int32_t) * ((size_t)v->n + 1));
    int run = 0;
    for (int id = 0; id < v->n; id++) { start[id] = run; if (cnt[id + 1]) { fv_node[nf] = id; fv_off[nf] = run; nf++; } run += cnt[id + 1]; }
    fv_off[nf] = run;
    for (int i = 0; i < n; i++) if (wts[i] > 0) fv_feat[start[node[i]]++] = i;
    *n_fv = nf;
    free(acc); free(present); free(wts); free(cnt); free(start);
}

/* L1Scoring::score (DBoW2 ScoringObject.cpp), the score of KeyFrameDatabase.cc:145,274 and LoopClosing.cc:152 */
double oc_bow_score_l1(const int32_t* id1, const double* v1, int n1, const int32_t* id2, const double* v2, int n2)
{
    double score = 0;
    int a = 0, b = 0;
    while (a < n1 && b < n2) {
        if (id1[a] == id2[b]) { score += fabs(v1[a] - v2[b]) - fabs(v1[a]) - fabs(v2[b]); a++; b++; }
        else if (id1[a] < id2[b]) a++;      /* lower_bound walk == advancing the smaller side */
        else b++;
    }
    return -score / 2.0;
}

/* ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:1797-1839) on bin counts */
static void three_maxima(const int* count, int L, int* ind1, int* ind2, int* ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    *ind1 = *ind2 = *ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = count[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { *ind3 = -1; }
}
#define HISTO_LENGTH 30
static int rot_bin(float a1, float a2)      /* ORBmatcher.cc:268-277 (the 1/30 factor is the reference's) */
{
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (ORBmatcher.cc:175-325).
 * kf_valid[i] != 0 <=> the keyframe's feature i has a map point that is not bad (:212-218).
 * match_f[j] = keyframe feature matched to frame feature j, or -1 (stands for vpMapPointMatches). Returns nmatches. */
int oc_search_by_bow(const int32_t* kf_fv_node, const int32_t* kf_fv_off, const int32_t* kf_fv_feat, int kf_nfv,
                     const int32_t* f_fv_node, const int32_t* f_fv_off, const int32_t* f_fv_feat, int f_nfv,
                     const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                     const uint8_t* f_desc, const float* f_angle, int f_n,
                     float nnratio, int check_orientation, int32_t* match_f)
{
    int nmatches = 0;
    int* hist_idx = (int*)malloc(sizeof(int) * (size_t)(f_n > 0 ? f_n : 1)), * hist_bin = (int*)malloc(sizeof(int) * (size_t)(f_n > 0 ? f_n : 1));
    int nh = 0, count[HISTO_LENGTH] = {0};
    for (int j = 0; j < f_n; j++) match_f[j] = -1;
    int a = 0, b = 0;
    while (a < kf_nfv && b < f_nfv) {
        if (kf_fv_node[a] == f_fv_node[b]) {
            for (int ik = kf_fv_off[a]; ik < kf_fv_off[a + 1]; ik++) {
                const int realIdxKF = kf_fv_feat[ik];
                if (!kf_valid[realIdxKF]) continue;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int jf = f_fv_off[b]; jf < f_fv_off[b + 1]; jf++) {
                    const int realIdxF = f_fv_feat[jf];
                    if (match_f[realIdxF] >= 0) continue;
                    const int dist = oc_descriptor_distance(kf_desc + 32 * (size_t)realIdxKF, f_desc + 32 * (size_t)realIdxF);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= 50 && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match_f[bestIdxF] = realIdxKF;
                    if (check_orientation) {
                        const int bin = rot_bin(kf_angle[realIdxKF], f_angle[bestIdxF]);
                        hist_idx[nh] = bestIdxF; hist_bin[nh] = bin; nh++; count[bin]++;
                    }
                    nmatches++;
                }
            }
            a++; b++;
        } else if (kf_fv_node[a] < f_fv_node[b]) a++;
        else b++;
    }
    if (check_orientation) {
        int i1, i2, i3;
        three_maxima(count, HISTO_LENGTH, &i1, &i2, &i3);
        for (int t = 0; t < nh; t++)
            if (hist_bin[t] != i1 && hist_bin[t] != i2 && hist_bin[t] != i3) { match_f[hist_idx[t]] = -1; nmatches--; }
    }
    free(hist_idx); free(hist_bin);
    return nmatches;
}

/* ------------------------------------------------------------------ ORBmatcher::SearchByProjection(Frame &CurrentFrame,
 * const Frame &LastFrame, th, bMono) (ORBmatcher.cc:1489-1646), the matcher of Tracking::TrackWithMotionModel.
 * cv::Mat arithmetic restated from OpenCV 4.13 (pinned with cv2.gemm): `Rcw*x3Dw+tcw` is one gemm on CV_32F 3x3 * 3x1
 * operands evaluated in f32, left to right, the addend last. `1.0/z` is a double division rounded to float. The
 * reference arithmetic is taken uncontracted (contraction note).
 * mode: 0 = neither (levels octave-1..octave+1), 1 = bForward, 2 = bBackward (:1510-1511; decided by the caller from
 * tlc and mb). cam = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY). last_flags bit 0: LastFrame.mvpMapPoints[i]
 * exists and !mvbOutlier[i]; bit 1: that map point has Observations() > 0 (what :1574-1576 tests once it has been
 * assigned to a current keypoint). cur_occupied: the same test on CurrentFrame.mvpMapPoints before the call.
 * match_cur[k] stands for CurrentFrame.mvpMapPoints[k]: index of the last-frame keypoint whose map point it holds. */
int oc_search_by_projection_frame(const OcKeyPoint* cur_kps, const uint8_t* cur_desc, int n_cur, const float* cur_u_right,
                                  const uint8_t* cur_occupied, const float* Tcw12, const float* cam9,
                                  const float* scale_factors,
                                  const OcKeyPoint* last_kps, const float* last_xyz, const uint8_t* last_desc,
                                  const uint8_t* last_flags, int n_last, float th, int mode, int check_orientation,
                                  int32_t* match_cur)
{
    const float fx = cam9[0], fy = cam9[1], cx = cam9[2], cy = cam9[3], mbf = cam9[4];
    const float mnMinX = cam9[5], mnMaxX = cam9[6], mnMinY = cam9[7], mnMaxY = cam9[8];
    const float invW = (float)FRAME_GRID_COLS / (mnMaxX - mnMinX), invH = (float)FRAME_GRID_ROWS / (mnMaxY - mnMinY);
    /* AssignFeaturesToGrid */
    int* cnt = (int*)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, sizeof(int));
    int* cell = (int*)malloc(sizeof(int) * (size_t)(n_cur > 0 ? n_cur : 1));
    int* tab = (int*)malloc(sizeof(int) * (size_t)(n_cur > 0 ? n_cur : 1));
    int* fill = (int*)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS, sizeof(int));
    uint8_t* occ = (uint8_t*)calloc((size_t)(n_cur > 0 ? n_cur : 1), 1);
    for (int i = 0; i < n_cur; i++) {
        int posX = (int)roundf((cur_kps[i].x - mnMinX) * invW), posY = (int)roundf((cur_kps[i].y - mnMinY) * invH);
        cell[i] = (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) ? -1 : posX * FRAME_GRID_ROWS + posY;
        if (cell[i] >= 0) cnt[cell[i] + 1]++;
        match_cur[i] = -1;
        occ[i] = cur_occupied ? (cur_occupied[i] != 0) : 0;
    }
    for (int c = 0; c < FRAME_GRID_COLS * FRAME_GRID_ROWS; c++) cnt[c + 1] += cnt[c];
    for (int i = 0; i < n_cur; i++) if (cell[i] >= 0) tab[cnt[cell[i]] + fill[cell[i]]++] = i;
    int nmatches = 0, nh = 0, count[HISTO_LENGTH] = {0};
    int* hist_idx = (int*)malloc(sizeof(int) * (size_t)(n_last > 0 ? n_last : 1));
    int* hist_bin = (int*)malloc(sizeof(int) * (size_t)(n_last > 0 ? n_last : 1));
    for (int i = 0; i < n_last; i++) {
        if (!(last_flags[i] & 1)) continue;
        const float X = last_xyz[3 * i], Y = last_xyz[3 * i + 1], Z = last_xyz[3 * i + 2];
        float c3[3];
        for (int r = 0; r < 3; r++) {
            float s = Tcw12[3 * r] * X;
            s = s + Tcw12[3 * r + 1] * Y;
            s = s + Tcw12[3 * r + 2] * Z;
            c3[r] = s + Tcw12[9 + r];
        }
        const float xc = c3[0], yc = c3[1];
        const float invzc = (float)(1.0 / (double)c3[2]);
        if (invzc < 0) continue;
        const float u = fx * xc * invzc + cx, v = fy * yc * invzc + cy;
        if (u < mnMinX || u > mnMaxX) continue;
        if (v < mnMinY || v > mnMaxY) continue;
        const int nLastOctave = last_kps[i].octave;
        const float radius = th * scale_factors[nLastOctave];
        int minLevel, maxLevel;
        if (mode == 1) { minLevel = nLastOctave; maxLevel = -1; }
        else if (mode == 2) { minLevel = 0; maxLevel = nLastOctave; }
        else { minLevel = nLastOctave - 1; maxLevel = nLastOctave + 1; }
        int nMinCellX = (int)floorf((u - mnMinX - radius) * invW); if (nMinCellX < 0) nMinCellX = 0;
        if (nMinCellX >= FRAME_GRID_COLS) continue;
        int nMaxCellX = (int)ceilf((u - mnMinX + radius) * invW); if (nMaxCellX > FRAME_GRID_COLS - 1) nMaxCellX = FRAME_GRID_COLS - 1;
        if (nMaxCellX < 0) continue;
        int nMinCellY = (int)floorf((v - mnMinY - radius) * invH); if (nMinCellY < 0) nMinCellY = 0;
        if (nMinCellY >= FRAME_GRID_ROWS) continue;
        int nMaxCellY = (int)ceilf((v - mnMinY + radius) * invH); if (nMaxCellY > FRAME_GRID_ROWS - 1) nMaxCellY = FRAME_GRID_ROWS - 1;
        if (nMaxCellY < 0) continue;
        const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        int bestDist = 256, bestIdx2 = -1;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const int c = ix * FRAME_GRID_ROWS + iy;
                for (int j = cnt[c]; j < cnt[c + 1]; j++) {
                    const int i2 = tab[j];
                    const OcKeyPoint* kp = &cur_kps[i2];
                    if (bCheckLevels) {
                        if (kp->octave < minLevel) continue;
                        if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                    }
                    if (!(fabsf(kp->x - u) < radius && fabsf(kp->y - v) < radius)) continue;
                    if (occ[i2]) continue;                                   /* :1574-1576 */
                    if (cur_u_right && cur_u_right[i2] > 0) {
                        const float ur = u - mbf * invzc;
                        const float er = fabsf(ur - cur_u_right[i2]);
                        if (er > radius) continue;
                    }
                    const int dist = oc_descriptor_distance(last_desc + 32 * (size_t)i, cur_desc + 32 * (size_t)i2);
                    if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
                }
            }
        if (bestDist <= 100) {                                               /* TH_HIGH */
            match_cur[bestIdx2] = i;
            occ[bestIdx2] = (last_flags[i] & 2) != 0;                        /* the new holder decides later skips */
            nmatches++;
            if (check_orientation) {
                const int bin = rot_bin(last_kps[i].angle, cur_kps[bestIdx2].angle);
                hist_idx[nh] = bestIdx2; hist_bin[nh] = bin; nh++; count[bin]++;
            }
        }
    }
    if (check_orientation) {
        int i1, i2, i3;
        three_maxima(count, HISTO_LENGTH, &i1, &i2, &i3);
        for (int t = 0; t < nh; t++)
            if (hist_bin[t] != i1 && hist_bin[t] != i2 && hist_bin[t] != i3) { match_cur[hist_idx[t]] = -1; nmatches--; }
    }
    free(cnt); free(cell); free(tab); free(fill); free(occ); free(hist_idx); free(hist_bin);
    return nmatches;
}

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (ORBmatcher.cc:589-736).
 * valid1 / valid2: the keyframe's feature has a map point that is not bad. match12[idx1] = feature of keyframe 2 whose
 * map point vpMatches12[idx1] holds, -1 = none. Returns nmatches. */
int oc_search_by_bow_kf(const int32_t* fv1_node, const int32_t* fv1_off, const int32_t* fv1_feat, int nfv1,
                        const int32_t* fv2_node, const int32_t* fv2_off, const int32_t* fv2_feat, int nfv2,
                        const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                        const uint8_t* desc2, const float* angle2, const uint8_t* valid2, int n2,
                        float nnratio, int check_orientation, int32_t* match12)
{
    int nmatches = 0, nh = 0, count[HISTO_LENGTH] = {0};
    int* hist_idx = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1)), * hist_bin = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
    uint8_t* matched2 = (uint8_t*)calloc((size_t)(n2 > 0 ? n2 : 1), 1);
    for (int i = 0; i < n1; i++) match12[i] = -1;
    int a = 0, b = 0;
    while (a < nfv1 && b < nfv2) {
        if (fv1_node[a] == fv2_node[b]) {
            for (int i1 = fv1_off[a]; i1 < fv1_off[a + 1]; i1++) {
                const int idx1 = fv1_feat[i1];
                if (!valid1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = fv2_off[b]; i2 < fv2_off[b + 1]; i2++) {
                    const int idx2 = fv2_feat[i2];
                    if (matched2[idx2] || !valid2[idx2]) continue;
                    const int dist = oc_descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < 50 && (float)bestDist1 < nnratio * (float)bestDist2) {
                    match12[idx1] = bestIdx2;
                    matched2[bestIdx2] = 1;
                    if (check_orientation) {
                        const int bin = rot_bin(angle1[idx1], angle2[bestIdx2]);
                        hist_idx[nh] = idx1; hist_bin[nh] = bin; nh++; count[bin]++;
                    }
                    nmatches++;
                }
            }
            a++; b++;
        } else if (fv1_node[a] < fv2_node[b]) a++;
        else b++;
    }
    if (check_orientation) {
        int i1, i2, i3;
        three_maxima(count, HISTO_LENGTH, &i1, &i2, &i3);
        for (int t = 0; t < nh; t++)
            if (hist_bin[t] != i1 && hist_bin[t] != i2 && hist_bin[t] != i3) { match12[hist_idx[t]] = -1; nmatches--; }
    }
    free(hist_idx); free(hist_bin); free(matched2);
    return nmatches;
}

/* ------------------------------------------------------------------ Frame grid helper shared by the matchers below */
typedef struct { int* cnt; int* tab; } OcGrid;
static OcGrid oc_grid_build(const OcKeyPoint* kps, int n, float mnMinX, float mnMinY, float invW, float invH)
{
    OcGrid g;
    g.cnt = (int*)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, sizeof(int));
    g.tab = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* cell = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    int* fill = (int*)calloc(FRAME_GRID_COLS * FRAME_GRID_ROWS, sizeof(int));
    for (int i = 0; i < n; i++) {
        int posX = (int)roundf((kps[i].x - mnMinX) * invW), posY = (int)roundf((kps[i].y - mnMinY) * invH);
        cell[i] = (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) ? -1 : posX * FRAME_GRID_ROWS + posY;
        if (cell[i] >= 0) g.cnt[cell[i] + 1]++;
    }
    for (int c = 0; c < FRAME_GRID_COLS * FRAME_GRID_ROWS; c++) g.cnt[c + 1] += g.cnt[c];
    for (int i = 0; i < n; i++) if (cell[i] >= 0) g.tab[g.cnt[cell[i]] + fill[cell[i]]++] = i;
    free(cell); free(fill);
    return g;
}
static void oc_grid_free(OcGrid* g) { free(g->cnt); free(g->tab); }
/* cell rectangle of GetFeaturesInArea (Frame.cc:394-408, KeyFrame.cc:713-727); 0 = empty result */
static int oc_grid_window(float x, float y, float r, float mnMinX, float mnMinY, float invW, float invH, int* x0, int* x1, int* y0, int* y1)
{
    *x0 = (int)floorf((x - mnMinX - r) * invW); if (*x0 < 0) *x0 = 0;
    if (*x0 >= FRAME_GRID_COLS) return 0;
    *x1 = (int)ceilf((x - mnMinX + r) * invW); if (*x1 > FRAME_GRID_COLS - 1) *x1 = FRAME_GRID_COLS - 1;
    if (*x1 < 0) return 0;
    *y0 = (int)floorf((y - mnMinY - r) * invH); if (*y0 < 0) *y0 = 0;
    if (*y0 >= FRAME_GRID_ROWS) return 0;
    *y1 = (int)ceilf((y - mnMinY + r) * invH); if (*y1 > FRAME_GRID_ROWS - 1) *y1 = FRAME_GRID_ROWS - 1;
    if (*y1 < 0) return 0;
    return 1;
}

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th) (ORBmatcher.cc:46-142),
 * called by Tracking::SearchLocalPoints. q[i] carries what Frame::isInFrustum left in the map point (mTrackProjX,
 * mTrackProjY, mTrackProjXR, mTrackViewCos, mnTrackScaleLevel); qflags bit 0: mbTrackInView && !isBad(), bit 1:
 * Observations() > 0. occupied[k]: F.mvpMapPoints[k] exists with Observations() > 0 before the call.
 * match[k] = index into vpMapPoints that keypoint k holds after the call, -1 = untouched. Returns nmatches. */
int oc_search_local_points(const OcKeyPoint* kps, const uint8_t* desc, int n, const float* u_right, const uint8_t* occupied,
                           const float* bounds4, const float* scale_factors, int nlevels,
                           const OcTrackQuery* q, const uint8_t* qdesc, const uint8_t* qflags, int nq,
                           float th, float nnratio, int32_t* match)
{
    const float mnMinX = bounds4[0], mnMaxX = bounds4[1], mnMinY = bounds4[2], mnMaxY = bounds4[3];
    const float invW = (float)FRAME_GRID_COLS / (mnMaxX - mnMinX), invH = (float)FRAME_GRID_ROWS / (mnMaxY - mnMinY);
    OcGrid g = oc_grid_build(kps, n, mnMinX, mnMinY, invW, invH);
    uint8_t* occ = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int i = 0; i < n; i++) { match[i] = -1; occ[i] = occupied ? (occupied[i] != 0) : 0; }
    int nmatches = 0;
    const int bFactor = th != 1.0;
    (void)nlevels;
    for (int iMP = 0; iMP < nq; iMP++) {
        if (!(qflags[iMP] & 1)) continue;
        const int nPredictedLevel = q[iMP].level;
        float r = q[iMP].view_cos > 0.998 ? 2.5 : 4.0;                          /* RadiusByViewingCos (:144-150) */
        if (bFactor) r *= th;
        const float rad = r * scale_factors[nPredictedLevel];
        const int minLevel = nPredictedLevel - 1, maxLevel = nPredictedLevel;
        int x0, x1, y0, y1;
        if (!oc_grid_window(q[iMP].x, q[iMP].y, rad, mnMinX, mnMinY, invW, invH, &x0, &x1, &y0, &y1)) continue;
        const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const int c = ix * FRAME_GRID_ROWS + iy;
                for (int j = g.cnt[c]; j < g.cnt[c + 1]; j++) {
                    const int idx = g.tab[j];
                    const OcKeyPoint* kp = &kps[idx];
                    if (bCheckLevels) {
                        if (kp->octave < minLevel) continue;
                        if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                    }
                    if (!(fabsf(kp->x - q[iMP].x) < rad && fabsf(kp->y - q[iMP].y) < rad)) continue;
                    if (occ[idx]) continue;                                       /* :90-92 */
                    if (u_right && u_right[idx] > 0) {
                        const float er = fabsf(q[iMP].xr - u_right[idx]);
                        if (er > r * scale_factors[nPredictedLevel]) continue;
                    }
                    const int dist = oc_descriptor_distance(qdesc + 32 * (size_t)iMP, desc + 32 * (size_t)idx);
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kp->octave; bestIdx = idx; }
                    else if (dist < bestDist2) { bestLevel2 = kp->octave; bestDist2 = dist; }
                }
            }
        if (bestDist <= 100) {                                                    /* TH_HIGH */
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            match[bestIdx] = iMP;                                                 /* F.mvpMapPoints[bestIdx] = pMP */
            occ[bestIdx] = (qflags[iMP] & 2) != 0;
            nmatches++;
        }
    }
    oc_grid_free(&g); free(occ);
    return nmatches;
}

/* MapPoint::PredictScale (MapPoint.cc:407-422); `log` resolves to std::log(float) because KeyFrame.h pulls in DBoW2's
 * global `using namespace std` (TemplatedVocabulary.h:36), so the expression is evaluated in f32. */
int oc_predict_scale(float max_distance, float current_dist, float log_scale_factor, int nlevels)
{
    const float ratio = max_distance / current_dist;
    int nScale = (int)ceilf(logf(ratio) / log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= nlevels) nScale = nlevels - 1;
    return nScale;
}

/* The search half of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (ORBmatcher.cc:918-1092; mode 0) and of
 * Fuse(KeyFrame*, cv::Mat Scw, ...) (:1094-1236; mode 1, Rcw / tcw / Ow already taken out of Scw by the caller as
 * :1101-1106 do). cv::Mat arithmetic restated from OpenCV 4.13: `Rcw*p3Dw + tcw` is one f32 gemm (products summed left
 * to right, addend last), cv::norm / Mat::dot of 3-vectors accumulate in f64 in element order.
 * cam = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY). pt_dist[i] = (GetMinDistanceInvariance,
 * GetMaxDistanceInvariance, mfMaxDistance). pt_flags bit 0: the point is non-NULL, not bad and not already in the keyframe.
 * best_idx[i] = keyframe feature to fuse with (bestDist <= TH_LOW) or -1; best_dist[i] = bestDist. Returns nFused. */
int oc_fuse_search(const OcKeyPoint* kps, const uint8_t* desc, int n, const float* u_right,
                   const float* Tcw12, const float* Ow3, const float* cam9, const float* scale_factors,
                   const float* inv_level_sigma2, int nlevels, float log_scale_factor,
                   const float* pt_xyz, const float* pt_normal, const float* pt_dist, const uint8_t* pt_desc,
                   const uint8_t* pt_flags, int npts, float th, int mode, int32_t* best_idx, int32_t* best_dist)
{
    const float fx = cam9[0], fy = cam9[1], cx = cam9[2], cy = cam9[3], bf = cam9[4];
    const float mnMinX = cam9[5], mnMaxX = cam9[6], mnMinY = cam9[7], mnMaxY = cam9[8];
    const float invW = (float)FRAME_GRID_COLS / (mnMaxX - mnMinX), invH = (float)FRAME_GRID_ROWS / (mnMaxY - mnMinY);
    OcGrid g = oc_grid_build(kps, n, mnMinX, mnMinY, invW, invH);      /* the Frame's grid, copied by KeyFrame.cc:52-57 */
    /* KeyFrame.h:236-239 keeps the bounds as const int: IsInImage and GetFeaturesInArea see the truncated values */
    const int kMinX = (int)mnMinX, kMaxX = (int)mnMaxX, kMinY = (int)mnMinY, kMaxY = (int)mnMaxY;
    int nFused = 0;
    for (int i = 0; i < npts; i++) {
        int bestDist = mode == 0 ? 256 : INT_MAX, bestIdx = -1;
        best_idx[i] = -1; best_dist[i] = bestDist;
        if (!(pt_flags[i] & 1)) continue;
        const float X = pt_xyz[3 * i], Y = pt_xyz[3 * i + 1], Z = pt_xyz[3 * i + 2];
        float c3[3];
        for (int r = 0; r < 3; r++) {
            float s = Tcw12[3 * r] * X;
            s = s + Tcw12[3 * r + 1] * Y;
            s = s + Tcw12[3 * r + 2] * Z;
            c3[r] = s + Tcw12[9 + r];
        }
        if (c3[2] < 0.0f) continue;
        const float invz = mode == 0 ? 1 / c3[2] : (float)(1.0 / c3[2]);          /* :954 / :1146 */
        const float x = c3[0] * invz, y = c3[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!(u >= kMinX && u < kMaxX && v >= kMinY && v < kMaxY)) continue;      /* KeyFrame::IsInImage */
        const float ur = u - bf * invz;
        const float maxDistance = pt_dist[3 * i + 1], minDistance = pt_dist[3 * i];
        const float PO[3] = {X - Ow3[0], Y - Ow3[1], Z - Ow3[2]};
        double s2 = 0.0;
        for (int k = 0; k < 3; k++) s2 += (double)PO[k] * (double)PO[k];
        const float dist3D = (float)sqrt(s2);
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        double dot = 0.0;
        for (int k = 0; k < 3; k++) dot += (double)PO[k] * (double)pt_normal[3 * i + k];
        if (dot < 0.5 * dist3D) continue;
        const int nPredictedLevel = oc_predict_scale(pt_dist[3 * i + 2], dist3D, log_scale_factor, nlevels);
        const float radius = th * scale_factors[nPredictedLevel];
        int x0, x1, y0, y1;
        if (!oc_grid_window(u, v, radius, (float)kMinX, (float)kMinY, invW, invH, &x0, &x1, &y0, &y1)) continue;
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const int c = ix * FRAME_GRID_ROWS + iy;
                for (int j = g.cnt[c]; j < g.cnt[c + 1]; j++) {
                    const int idx = g.tab[j];
                    const OcKeyPoint* kp = &kps[idx];
                    if (!(fabsf(kp->x - u) < radius && fabsf(kp->y - v) < radius)) continue;   /* KeyFrame::GetFeaturesInArea */
                    const int kpLevel = kp->octave;
                    if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
                    if (mode == 0) {
                        if (u_right && u_right[idx] >= 0) {
                            const float ex = u - kp->x, ey = v - kp->y, er = ur - u_right[idx];
                            const float e2 = ex * ex + ey * ey + er * er;
                            if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
                        } else {
                            const float ex = u - kp->x, ey = v - kp->y;
                            const float e2 = ex * ex + ey * ey;
                            if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
                        }
                    }
                    const int dist = oc_descriptor_distance(pt_desc + 32 * (size_t)i, desc + 32 * (size_t)idx);
                    if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
                }
            }
        best_dist[i] = bestDist;
        if (bestDist <= 50) { best_idx[i] = bestIdx; nFused++; }                   /* TH_LOW */
    }
    oc_grid_free(&g);
    return nFused;
}

/* ORBmatcher::SearchForTriangulation (ORBmatcher.cc:738-916). skip1 / skip2: the feature already has a map point
 * (GetMapPoint != NULL). u_right1 / u_right2: mvuRight (NULL = monocular keyframes, i.e. all -1). F12 row-major 3x3.
 * pose2 = (R2w row-major 9, t2w 3), Cw1 = pKF1->GetCameraCenter(), K2 = (fx, fy, cx, cy) of pKF2.
 * Note that the reference never sets vbMatched2, so several features of keyframe 1 may share one of keyframe 2.
 * match12[idx1] = idx2 or -1 (vMatchedPairs = the non-negative entries in ascending idx1). Returns nmatches. */
int oc_search_for_triangulation(const int32_t* fv1_node, const int32_t* fv1_off, const int32_t* fv1_feat, int nfv1,
                                const int32_t* fv2_node, const int32_t* fv2_off, const int32_t* fv2_feat, int nfv2,
                                const OcKeyPoint* kps1, const uint8_t* desc1, const uint8_t* skip1, const float* u_right1, int n1,
                                const OcKeyPoint* kps2, const uint8_t* desc2, const uint8_t* skip2, const float* u_right2, int n2,
                                const float* F12, const float* Cw1, const float* pose2, const float* K2,
                                const float* scale_factors2, const float* level_sigma2_2,
                                int only_stereo, int check_orientation, int32_t* match12)
{
    (void)n2;
    float C2[3];
    for (int r = 0; r < 3; r++) {
        float s = pose2[3 * r] * Cw1[0];
        s = s + pose2[3 * r + 1] * Cw1[1];
        s = s + pose2[3 * r + 2] * Cw1[2];
        C2[r] = s + pose2[9 + r];
    }
    const float invz = 1.0f / C2[2];
    const float ex = K2[0] * C2[0] * invz + K2[2];
    const float ey = K2[1] * C2[1] * invz + K2[3];
    int nmatches = 0, nh = 0, count[HISTO_LENGTH] = {0};
    int* hist_idx = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1)), * hist_bin = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
    for (int i = 0; i < n1; i++) match12[i] = -1;
    int a = 0, b = 0;
    while (a < nfv1 && b < nfv2) {
        if (fv1_node[a] == fv2_node[b]) {
            for (int i1 = fv1_off[a]; i1 < fv1_off[a + 1]; i1++) {
                const int idx1 = fv1_feat[i1];
                if (skip1 && skip1[idx1]) continue;
                const int bStereo1 = u_right1 ? u_right1[idx1] >= 0 : 0;
                if (only_stereo && !bStereo1) continue;
                const OcKeyPoint* kp1 = &kps1[idx1];
                int bestDist = 50, bestIdx2 = -1;                                 /* TH_LOW */
                for (int i2 = fv2_off[b]; i2 < fv2_off[b + 1]; i2++) {
                    const int idx2 = fv2_feat[i2];
                    if (skip2 && skip2[idx2]) continue;
                    const int bStereo2 = u_right2 ? u_right2[idx2] >= 0 : 0;
                    if (only_stereo && !bStereo2) continue;
                    const int dist = oc_descriptor_distance(desc1 + 32 * (size_t)idx1, desc2 + 32 * (size_t)idx2);
                    if (dist > 50 || dist > bestDist) continue;
                    const OcKeyPoint* kp2 = &kps2[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2->x, distey = ey - kp2->y;
                        if (distex * distex + distey * distey < 100 * scale_factors2[kp2->octave]) continue;
                    }
                    /* CheckDistEpipolarLine (:153-173) */
                    const float fa = kp1->x * F12[0] + kp1->y * F12[3] + F12[6];
                    const float fb = kp1->x * F12[1] + kp1->y * F12[4] + F12[7];
                    const float fc = kp1->x * F12[2] + kp1->y * F12[5] + F12[8];
                    const float num = fa * kp2->x + fb * kp2->y + fc;
                    const float den = fa * fa + fb * fb;
                    if (den == 0) continue;
                    const float dsqr = num * num / den;
                    if (dsqr < 3.84 * level_sigma2_2[kp2->octave]) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    nmatches++;
                    if (check_orientation) {
                        const int bin = rot_bin(kp1->angle, kps2[bestIdx2].angle);
                        hist_idx[nh] = idx1; hist_bin[nh] = bin; nh++; count[bin]++;
                    }
                }
            }
            a++; b++;
        } else if (fv1_node[a] < fv2_node[b]) a++;
        else b++;
    }
    if (check_orientation) {
        int i1, i2, i3;
        three_maxima(count, HISTO_LENGTH, &i1, &i2, &i3);
        for (int t = 0; t < nh; t++)
            if (hist_bin[t] != i1 && hist_bin[t] != i2 && hist_bin[t] != i3) { match12[hist_idx[t]] = -1; nmatches--; }
    }
    free(hist_idx); free(hist_bin);
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
 * (ORBmatcher.cc:1648-1795; mode 0, the relocalisation matcher) and
 * ORBmatcher::SearchByProjection(KeyFrame *pKF, cv::Mat Scw, vpPoints, vpMatched, int th) (:327-440; mode 1, loop closing,
 * Rcw / tcw / Ow taken out of Scw by the caller as :333-339 do; th_dist = TH_LOW).
 * occupied[k]: the frame / keyframe feature already holds a map point before the call (CurrentFrame.mvpMapPoints[k] /
 * vpMatched[k] non-NULL). pt_flags bit 0: the point is non-NULL, not bad and not in sAlreadyFound / spAlreadyFound.
 * pt_angle[i] = pKF->mvKeysUn[i].angle (mode 0 with check_orientation). match[k] = point index the feature now holds,
 * -1 = untouched. Returns nmatches. */
int oc_search_by_projection_seq(const OcKeyPoint* kps, const uint8_t* desc, int n, const uint8_t* occupied,
                                const float* Tcw12, const float* Ow3, const float* cam9, const float* scale_factors,
                                int nlevels, float log_scale_factor,
                                const float* pt_xyz, const float* pt_normal, const float* pt_dist, const uint8_t* pt_desc,
                                const uint8_t* pt_flags, const float* pt_angle, int npts,
                                float th, int th_dist, int mode, int check_orientation, int32_t* match)
{
    const float fx = cam9[0], fy = cam9[1], cx = cam9[2], cy = cam9[3];
    const float mnMinX = cam9[5], mnMaxX = cam9[6], mnMinY = cam9[7], mnMaxY = cam9[8];
    const float invW = (float)FRAME_GRID_COLS / (mnMaxX - mnMinX), invH = (float)FRAME_GRID_ROWS / (mnMaxY - mnMinY);
    OcGrid g = oc_grid_build(kps, n, mnMinX, mnMinY, invW, invH);
    const int kMinX = (int)mnMinX, kMaxX = (int)mnMaxX, kMinY = (int)mnMinY, kMaxY = (int)mnMaxY;   /* KeyFrame bounds (mode 1) */
    uint8_t* occ = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int i = 0; i < n; i++) { match[i] = -1; occ[i] = occupied ? (occupied[i] != 0) : 0; }
    int nmatches = 0, nh = 0, count[HISTO_LENGTH] = {0};
    int* hist_idx = (int*)malloc(sizeof(int) * (size_t)(npts > 0 ? npts : 1)), * hist_bin = (int*)malloc(sizeof(int) * (size_t)(npts > 0 ? npts : 1));
    for (int i = 0; i < npts; i++) {
        if (!(pt_flags[i] & 1)) continue;
        const float X = pt_xyz[3 * i], Y = pt_xyz[3 * i + 1], Z = pt_xyz[3 * i + 2];
        float c3[3];
        for (int r = 0; r < 3; r++) {
            float s = Tcw12[3 * r] * X;
            s = s + Tcw12[3 * r + 1] * Y;
            s = s + Tcw12[3 * r + 2] * Z;
            c3[r] = s + Tcw12[9 + r];
        }
        float u, v;
        if (mode == 0) {
            const float xc = c3[0], yc = c3[1];
            const float invzc = (float)(1.0 / c3[2]);
            u = fx * xc * invzc + cx; v = fy * yc * invzc + cy;
            if (u < mnMinX || u > mnMaxX) continue;
            if (v < mnMinY || v > mnMaxY) continue;
        } else {
            if (c3[2] < 0.0) continue;
            const float invz = 1 / c3[2];
            const float x = c3[0] * invz, y = c3[1] * invz;
            u = fx * x + cx; v = fy * y + cy;
            if (!(u >= kMinX && u < kMaxX && v >= kMinY && v < kMaxY)) continue;
        }
        const float PO[3] = {X - Ow3[0], Y - Ow3[1], Z - Ow3[2]};
        double s2 = 0.0;
        for (int k = 0; k < 3; k++) s2 += (double)PO[k] * (double)PO[k];
        const float dist3D = (float)sqrt(s2);
        if (dist3D < pt_dist[3 * i] || dist3D > pt_dist[3 * i + 1]) continue;
        if (mode == 1) {
            double dot = 0.0;
            for (int k = 0; k < 3; k++) dot += (double)PO[k] * (double)pt_normal[3 * i + k];
            if (dot < 0.5 * dist3D) continue;
        }
        const int nPredictedLevel = oc_predict_scale(pt_dist[3 * i + 2], dist3D, log_scale_factor, nlevels);
        const float radius = th * scale_factors[nPredictedLevel];
        const int minLevel = nPredictedLevel - 1, maxLevel = mode == 0 ? nPredictedLevel + 1 : nPredictedLevel;
        int x0, x1, y0, y1;
        if (!oc_grid_window(u, v, radius, mode == 0 ? mnMinX : (float)kMinX, mode == 0 ? mnMinY : (float)kMinY, invW, invH, &x0, &x1, &y0, &y1)) continue;
        int bestDist = 256, bestIdx = -1;
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const int c = ix * FRAME_GRID_ROWS + iy;
                for (int j = g.cnt[c]; j < g.cnt[c + 1]; j++) {
                    const int idx = g.tab[j];
                    const OcKeyPoint* kp = &kps[idx];
                    /* mode 0: Frame::GetFeaturesInArea(.., level-1, level+1): maxLevel >= 0 always, so both bounds are checked;
                       mode 1: KeyFrame::GetFeaturesInArea + the explicit level test (:415-418) */
                    if (kp->octave < minLevel || kp->octave > maxLevel) continue;
                    if (!(fabsf(kp->x - u) < radius && fabsf(kp->y - v) < radius)) continue;
                    if (occ[idx]) continue;
                    const int dist = oc_descriptor_distance(pt_desc + 32 * (size_t)i, desc + 32 * (size_t)idx);
                    if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
                }
            }
        if (bestDist <= th_dist) {
            match[bestIdx] = i; occ[bestIdx] = 1;
            nmatches++;
            if (mode == 0 && check_orientation) {
                const int bin = rot_bin(pt_angle[i], kps[bestIdx].angle);
                hist_idx[nh] = bestIdx; hist_bin[nh] = bin; nh++; count[bin]++;
            }
        }
    }
    if (mode == 0 && check_orientation) {
        int i1, i2, i3;
        three_maxima(count, HISTO_LENGTH, &i1, &i2, &i3);
        for (int t = 0; t < nh; t++)
            if (hist_bin[t] != i1 && hist_bin[t] != i2 && hist_bin[t] != i3) { match[hist_idx[t]] = -1; nmatches--; }
    }
    oc_grid_free(&g); free(occ); free(hist_idx); free(hist_bin);
    return nmatches;
}

/* One direction of ORBmatcher::SearchBySim3 (ORBmatcher.cc:1281-1359 / :1362-1440): points of the source keyframe go
 * through T1 = (R1w, t1w) and T2 = (sR21, t21) into the target camera; vnMatch[i] = target feature or -1. */
static void sim3_direction(const OcKeyPoint* kps, const uint8_t* desc, int n, const float* T1, const float* T2,
                           const float* cam9, const float* scale_factors, int nlevels, float log_scale_factor,
                           const float* pt_xyz, const float* pt_dist, const uint8_t* pt_desc, const uint8_t* pt_flags, int npts,
                           float th, int32_t* vnMatch)
{
    const float fx = cam9[0], fy = cam9[1], cx = cam9[2], cy = cam9[3];
    const float mnMinX = cam9[5], mnMaxX = cam9[6], mnMinY = cam9[7], mnMaxY = cam9[8];
    const float invW = (float)FRAME_GRID_COLS / (mnMaxX - mnMinX), invH = (float)FRAME_GRID_ROWS / (mnMaxY - mnMinY);
    OcGrid g = oc_grid_build(kps, n, mnMinX, mnMinY, invW, invH);
    const int kMinX = (int)mnMinX, kMaxX = (int)mnMaxX, kMinY = (int)mnMinY, kMaxY = (int)mnMaxY;   /* KeyFrame bounds */
    for (int i = 0; i < npts; i++) {
        vnMatch[i] = -1;
        if (!(pt_flags[i] & 1)) continue;                         /* NULL, already matched or bad */
        const float X = pt_xyz[3 * i], Y = pt_xyz[3 * i + 1], Z = pt_xyz[3 * i + 2];
        float a3[3], c3[3];
        for (int r = 0; r < 3; r++) {
            float s = T1[3 * r] * X;
            s = s + T1[3 * r + 1] * Y;
            s = s + T1[3 * r + 2] * Z;
            a3[r] = s + T1[9 + r];
        }
        for (int r = 0; r < 3; r++) {
            float s = T2[3 * r] * a3[0];
            s = s + T2[3 * r + 1] * a3[1];
            s = s + T2[3 * r + 2] * a3[2];
            c3[r] = s + T2[9 + r];
        }
        if (c3[2] < 0.0) continue;
        const float invz = (float)(1.0 / c3[2]);
        const float x = c3[0] * invz, y = c3[1] * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!(u >= kMinX && u < kMaxX && v >= kMinY && v < kMaxY)) continue;
        double s2 = 0.0;
        for (int k = 0; k < 3; k++) s2 += (double)c3[k] * (double)c3[k];
        const float dist3D = (float)sqrt(s2);
        if (dist3D < pt_dist[3 * i] || dist3D > pt_dist[3 * i + 1]) continue;
        const int nPredictedLevel = oc_predict_scale(pt_dist[3 * i + 2], dist3D, log_scale_factor, nlevels);
        const float radius = th * scale_factors[nPredictedLevel];
        int x0, x1, y0, y1;
        if (!oc_grid_window(u, v, radius, (float)kMinX, (float)kMinY, invW, invH, &x0, &x1, &y0, &y1)) continue;
        int bestDist = INT_MAX, bestIdx = -1;
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const int c = ix * FRAME_GRID_ROWS + iy;
                for (int j = g.cnt[c]; j < g.cnt[c + 1]; j++) {
                    const int idx = g.tab[j];
                    const OcKeyPoint* kp = &kps[idx];
                    if (!(fabsf(kp->x - u) < radius && fabsf(kp->y - v) < radius)) continue;
                    if (kp->octave < nPredictedLevel - 1 || kp->octave > nPredictedLevel) continue;
                    const int dist = oc_descriptor_distance(pt_desc + 32 * (size_t)i, desc + 32 * (size_t)idx);
                    if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
                }
            }
        if (bestDist <= 100) vnMatch[i] = bestIdx;                 /* TH_HIGH */
    }
    oc_grid_free(&g);
}

/* ORBmatcher::SearchBySim3 (ORBmatcher.cc:1238-1487). Per keyframe: undistorted keypoints, descriptors, and for every
 * feature its map point (xyz, (min, max distance invariance, mfMaxDistance), descriptor) with flags bit 0 = the feature has
 * a map point that is not bad and is not already matched (vbAlreadyMatched). T1w / T2w = (R, t) of the keyframes,
 * S12 = (sR12, t12), S21 = (sR21, t21) as :1253-1255 build them. match12[i1] = feature of keyframe 2 whose map point
 * vpMatches12[i1] receives, -1 = none. Returns nFound. */
int oc_search_by_sim3(const OcKeyPoint* kps1, const uint8_t* desc1, int n1, const float* xyz1, const float* dist1,
                      const uint8_t* mpdesc1, const uint8_t* flags1,
                      const OcKeyPoint* kps2, const uint8_t* desc2, int n2, const float* xyz2, const float* dist2,
                      const uint8_t* mpdesc2, const uint8_t* flags2,
                      const float* T1w, const float* T2w, const float* S12, const float* S21,
                      const float* cam9, const float* scale_factors, int nlevels, float log_scale_factor, float th,
                      int32_t* match12)
{
    int32_t* m1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n1 > 0 ? n1 : 1));
    int32_t* m2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
    sim3_direction(kps2, desc2, n2, T1w, S21, cam9, scale_factors, nlevels, log_scale_factor, xyz1, dist1, mpdesc1, flags1, n1, th, m1);
    sim3_direction(kps1, desc1, n1, T2w, S12, cam9, scale_factors, nlevels, log_scale_factor, xyz2, dist2, mpdesc2, flags2, n2, th, m2);
    int nFound = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        match12[i1] = -1;
        const int idx2 = m1[i1];
        if (idx2 >= 0 && m2[idx2] == i1) { match12[i1] = idx2; nFound++; }
    }
    free(m1); free(m2);
    return nFound;
}

/* ORBmatcher::SearchForInitialization (ORBmatcher.cc:442-587). prev = vbPrevMatched (2 floats per F1 keypoint, updated in
 * place like :582-584); match12 = vnMatches12. Returns nmatches. */
int oc_search_for_initialization(const OcKeyPoint* kps1, const uint8_t* desc1, int n1,
                                 const OcKeyPoint* kps2, const uint8_t* desc2, int n2, const float* bounds4,
                                 float* prev, int window, float nnratio, int check_orientation, int32_t* match12)
{
    const float mnMinX = bounds4[0], mnMaxX = bounds4[1], mnMinY = bounds4[2], mnMaxY = bounds4[3];
    const float invW = (float)FRAME_GRID_COLS / (mnMaxX - mnMinX), invH = (float)FRAME_GRID_ROWS / (mnMaxY - mnMinY);
    OcGrid g = oc_grid_build(kps2, n2, mnMinX, mnMinY, invW, invH);
    int nmatches = 0, nh = 0, count[HISTO_LENGTH] = {0};
    int* hist_idx = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1)), * hist_bin = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
    int* vMatchedDistance = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    int* vnMatches21 = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    for (int i = 0; i < n2; i++) { vMatchedDistance[i] = INT_MAX; vnMatches21[i] = -1; }
    for (int i = 0; i < n1; i++) match12[i] = -1;
    const float r = (float)window;
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = kps1[i1].octave;
        if (level1 > 0) continue;
        const float x = prev[2 * i1], y = prev[2 * i1 + 1];
        int x0, x1, y0, y1;
        if (!oc_grid_window(x, y, r, mnMinX, mnMinY, invW, invH, &x0, &x1, &y0, &y1)) continue;
        const int minLevel = level1, maxLevel = level1;
        const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int ix = x0; ix <= x1; ix++)
            for (int iy = y0; iy <= y1; iy++) {
                const int c = ix * FRAME_GRID_ROWS + iy;
                for (int j = g.cnt[c]; j < g.cnt[c + 1]; j++) {
                    const int i2 = g.tab[j];
                    const OcKeyPoint* kp = &kps2[i2];
                    if (bCheckLevels) {
                        if (kp->octave < minLevel) continue;
                        if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                    }
                    if (!(fabsf(kp->x - x) < r && fabsf(kp->y - y) < r)) continue;
                    const int dist = oc_descriptor_distance(desc1 + 32 * (size_t)i1, desc2 + 32 * (size_t)i2);
                    if (vMatchedDistance[i2] <= dist) continue;
                    if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
            }
        if (bestDist <= 50) {                                                     /* TH_LOW */
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { match12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                match12[i1] = bestIdx2; vnMatches21[bestIdx2] = i1; vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (check_orientation) {
                    const int bin = rot_bin(kps1[i1].angle, kps2[bestIdx2].angle);
                    hist_idx[nh] = i1; hist_bin[nh] = bin; nh++; count[bin]++;
                }
            }
        }
    }
    if (check_orientation) {
        int i1, i2, i3;
        three_maxima(count, HISTO_LENGTH, &i1, &i2, &i3);
        for (int t = 0; t < nh; t++)
            if (hist_bin[t] != i1 && hist_bin[t] != i2 && hist_bin[t] != i3 && match12[hist_idx[t]] >= 0) { match12[hist_idx[t]] = -1; nmatches--; }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (match12[i1] >= 0) { prev[2 * i1] = kps2[match12[i1]].x; prev[2 * i1 + 1] = kps2[match12[i1]].y; }
    oc_grid_free(&g); free(hist_idx); free(hist_bin); free(vMatchedDistance); free(vnMatches21);
    return nmatches;
}


/* Frame::isInFrustum (Frame.cc:315-378) for npts map points: q[i] (the five mTrack* fields) is written where the function
 * returns true, in_view[i] = its return value. cam9 = (fx, fy, cx, cy, mbf, mnMinX, mnMaxX, mnMinY, mnMaxY). */
void oc_is_in_frustum(const float* Tcw12, const float* Ow3, const float* cam9, int nlevels, float log_scale_factor,
                      const float* pt_xyz, const float* pt_normal, const float* pt_dist, int npts, float viewingCosLimit,
                      OcTrackQuery* q, uint8_t* in_view)
{
    const float fx = cam9[0], fy = cam9[1], cx = cam9[2], cy = cam9[3], mbf = cam9[4];
    const float mnMinX = cam9[5], mnMaxX = cam9[6], mnMinY = cam9[7], mnMaxY = cam9[8];
    for (int i = 0; i < npts; i++) {
        in_view[i] = 0;
        const float X = pt_xyz[3 * i], Y = pt_xyz[3 * i + 1], Z = pt_xyz[3 * i + 2];
        float Pc[3];
        for (int r = 0; r < 3; r++) {
            float s = Tcw12[3 * r] * X;
            s = s + Tcw12[3 * r + 1] * Y;
            s = s + Tcw12[3 * r + 2] * Z;
            Pc[r] = s + Tcw12[9 + r];
        }
        if (Pc[2] < 0.0f) continue;
        const float invz = 1.0f / Pc[2];
        const float u = fx * Pc[0] * invz + cx, v = fy * Pc[1] * invz + cy;
        if (u < mnMinX || u > mnMaxX) continue;
        if (v < mnMinY || v > mnMaxY) continue;
        const float PO[3] = {X - Ow3[0], Y - Ow3[1], Z - Ow3[2]};
        double s2 = 0.0, dot = 0.0;
        for (int k = 0; k < 3; k++) s2 += (double)PO[k] * (double)PO[k];
        const float dist = (float)sqrt(s2);
        if (dist < pt_dist[3 * i] || dist > pt_dist[3 * i + 1]) continue;
        for (int k = 0; k < 3; k++) dot += (double)PO[k] * (double)pt_normal[3 * i + k];
        const float viewCos = dot / dist;
        if (viewCos < viewingCosLimit) continue;
        if (u != u || v != v || viewCos != viewCos) continue;       /* NaN (a point at the camera centre): outside the domain */
        q[i].level = oc_predict_scale(pt_dist[3 * i + 2], dist, log_scale_factor, nlevels);
        q[i].x = u; q[i].xr = u - mbf * invz; q[i].y = v; q[i].view_cos = viewCos;
        in_view[i] = 1;
    }
}
