// cvmini — NOT OpenCV. A tiny stand-in for <opencv2/core/core.hpp> used only to compile-check the host shim
// (host/ORBextractor.{h,cc}) in images without OpenCV's C++ headers. A real build uses real OpenCV instead.
#ifndef CVMINI_CORE_HPP
#define CVMINI_CORE_HPP
#include <cstdlib>
#include <cstring>
#define CV_8U 0
#define CV_8UC1 0
namespace cv {
typedef unsigned char uchar;
struct Rect { int x, y, width, height; Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} };
typedef Point_<float> Point2f;
class KeyPoint { public: Point2f pt; float size, angle, response; int octave, class_id; KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {} };
class Mat {
public:
    int rows, cols; uchar* data; size_t step;
    Mat() : rows(0), cols(0), data(0), step(0), rc(0) {}
    Mat(int r, int c, int) : rows(0), cols(0), data(0), step(0), rc(0) { create(r, c, 0); }
    Mat(int r, int c, int, void* p, size_t s) : rows(r), cols(c), data((uchar*)p), step(s), rc(0) {}
    Mat(const Mat& m) : rows(m.rows), cols(m.cols), data(m.data), step(m.step), rc(m.rc) { if (rc) ++*rc; }
    Mat& operator=(const Mat& m) { if (this != &m) { if (m.rc) ++*m.rc; release(); rows = m.rows; cols = m.cols; data = m.data; step = m.step; rc = m.rc; } return *this; }
    ~Mat() { release(); }
    void create(int r, int c, int) { if (data && r == rows && c == cols) return; release(); rows = r; cols = c; step = (size_t)c; rc = (int*)std::malloc(64 + (size_t)r * c + 1); *rc = 1; data = (uchar*)rc + 64; }
    void release() { if (rc && --*rc == 0) std::free(rc); rc = 0; data = 0; rows = cols = 0; step = 0; }
    bool empty() const { return !data || !rows || !cols; }
    int type() const { return CV_8UC1; }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    Mat operator()(const Rect& r) const { Mat m(*this); m.data += (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
private:
    int* rc;
};
class _InputArray { public: _InputArray(const Mat& m) : p(const_cast<Mat*>(&m)) {} bool empty() const { return p->empty(); } Mat getMat() const { return *p; } protected: Mat* p; };
class _OutputArray : public _InputArray { public: _OutputArray(Mat& m) : _InputArray(m) {} void create(int r, int c, int t) const { p->create(r, c, t); } void release() const { p->release(); } };
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
}
#endif
