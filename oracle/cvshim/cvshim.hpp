/*
 * cvshim.hpp — the smallest stand-in for the OpenCV C++ API that lets the UNMODIFIED reference file
 * /root/reference/src/ORBextractor.cc compile (oracle/_ref build, test infrastructure only).
 *
 * OpenCV C++ headers/libs are not installed in this image (only the cv2 4.13 Python wheel), so the
 * reference's own control flow (cell loop, quadtree, list order, descriptor loop) is kept verbatim while
 * the five primitives it calls are routed to oracle/orb_oracle.c, each of which is pinned bit-exactly to
 * cv2 4.13 by tests/test_oracle_golden.py.
 */
#ifndef CVSHIM_HPP
#define CVSHIM_HPP
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "orb_oracle.h"

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

namespace cv {
typedef unsigned char uchar;

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {} };

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt; float size, angle, response; int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct ZerosExpr { int rows, cols; };

/* 8-bit single-channel matrix header with malloc-backed, reference-counted storage (never operator new:
 * pyramid levels outlive a call while the bump allocator of ref_glue.cc is reset per call). */
class Mat {
public:
    int rows, cols; uchar* data; size_t step;
    Mat() : rows(0), cols(0), data(0), step(0), rc(0) {}
    Mat(Size s, int) : rows(0), cols(0), data(0), step(0), rc(0) { create(s.height, s.width, 0); }
    Mat(int r, int c, int) : rows(0), cols(0), data(0), step(0), rc(0) { create(r, c, 0); }
    Mat(int r, int c, int, void* ext, size_t st) : rows(r), cols(c), data((uchar*)ext), step(st), rc(0) {}
    Mat(const Mat& m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), rc(m.rc) { if (rc) __sync_fetch_and_add(rc, 1); }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m)
    {
        if (this != &m) { if (m.rc) __sync_fetch_and_add(m.rc, 1); release(); rows = m.rows; cols = m.cols; data = m.data; step = m.step; rc = m.rc; }
        return *this;
    }
    /* cv: `m = Mat::zeros(r,c,t)` evaluates INTO m: create() keeps a header of matching size (e.g. a row-range
     * view of the output descriptors), then fills with zeros. */
    Mat& operator=(const ZerosExpr& z)
    {
        create(z.rows, z.cols, 0);
        for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, (size_t)cols);
        return *this;
    }
    static ZerosExpr zeros(int r, int c, int) { ZerosExpr z = {r, c}; return z; }
    void create(int r, int c, int)
    {
        if (data && r == rows && c == cols) return;
        release();
        rows = r; cols = c; step = (size_t)c;
        size_t bytes = (size_t)r * (size_t)c;
        rc = (int*)malloc(64 + (bytes ? bytes : 1));
        *rc = 1; data = (uchar*)rc + 64;
    }
    void release()
    {
        if (rc && __sync_sub_and_fetch(rc, 1) == 0) free(rc);
        rc = 0; data = 0; rows = cols = 0; step = 0;
    }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }
    Mat clone() const
    {
        Mat m(rows, cols, 0);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols);
        return m;
    }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data += (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data += a; m.cols = b - a; return m; }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
private:
    int* rc;
};

class _InputArray {
public:
    _InputArray(const Mat& m) : p(const_cast<Mat*>(&m)) {}
    bool empty() const { return p->empty(); }
    Mat getMat() const { return *p; }
protected:
    Mat* p;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int t) const { p->create(r, c, t); }
    void create(Size s, int t) const { p->create(s.height, s.width, t); }
    void release() const { p->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline int cvRound(float v) { return oc_round_f(v); }
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline float fastAtan2(float y, float x) { return oc_fast_atan2(y, x); }

inline void resize(InputArray _src, OutputArray _dst, Size dsize, double = 0, double = 0, int interp = INTER_LINEAR)
{
    assert(interp == INTER_LINEAR);
    Mat src = _src.getMat();
    _dst.create(dsize, src.type());
    Mat dst = _dst.getMat();
    oc_resize_linear_8u(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}
inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int btype)
{
    assert((btype & ~BORDER_ISOLATED) == BORDER_REFLECT_101 && top == bottom && top == left && top == right);
    Mat src = _src.getMat();
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    uchar* pay = dst.data + (size_t)top * dst.step + left;
    if (pay != src.data)
        for (int y = 0; y < src.rows; y++) memmove(pay + (size_t)y * dst.step, src.data + (size_t)y * src.step, (size_t)src.cols);
    oc_border_reflect101(dst.data, src.cols, src.rows, (int)dst.step, top);
}
inline void GaussianBlur(InputArray _src, OutputArray _dst, Size k, double sx, double sy, int btype)
{
    assert(k.width == 7 && k.height == 7 && sx == 2 && sy == 2 && btype == BORDER_REFLECT_101);
    Mat src = _src.getMat();
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    oc_gaussian7x7_s2(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step);
}
inline void FAST(InputArray _img, std::vector<KeyPoint>& kps, int threshold, bool nms = true)
{
    Mat img = _img.getMat();
    kps.clear();
    if (img.rows < 7 || img.cols < 7) return;
    std::vector<OcKeyPoint> tmp((size_t)img.rows * (size_t)img.cols);
    int n = oc_fast9_16(img.data, img.cols, img.rows, (int)img.step, threshold, nms ? 1 : 0, tmp.data(), (int)tmp.size());
    kps.reserve((size_t)n);
    for (int i = 0; i < n; i++)
        kps.push_back(KeyPoint(tmp[i].x, tmp[i].y, tmp[i].size, tmp[i].angle, tmp[i].response, tmp[i].octave, tmp[i].class_id));
}
/* only referenced by the reference's dead ComputeKeyPointsOld (call commented out at ORBextractor.cc:1155) */
struct KeyPointsFilter {
    static bool byResponse(const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; }
    static void retainBest(std::vector<KeyPoint>& k, int n)
    {
        if (n >= 0 && k.size() > (size_t)n) { std::stable_sort(k.begin(), k.end(), byResponse); k.resize((size_t)n); }
    }
};
} // namespace cv
#endif
