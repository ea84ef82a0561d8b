/*
 * oracle/slamshim — stand-ins that let the UNMODIFIED /root/reference/src/ORBmatcher.cc compile (oracle/_ref build,
 * test infrastructure only): a small cv::Mat for CV_32F matrices and CV_8U descriptor rows, and stub Frame / KeyFrame /
 * MapPoint classes (Frame.h, KeyFrame.h, MapPoint.h next to this directory) that carry exactly the members the matcher
 * reads. OpenCV C++ is not installed in this image; the arithmetic of the expressions the matcher evaluates is the one
 * pinned against cv2 4.13: `A*B` of small CV_32F matrices sums f32 products left to right, `+ C` adds
 * last, element-wise ops are f32, cv::norm / Mat::dot of 3-vectors accumulate in f64 in element order.
 */
#ifndef SLAMSHIM_CORE_HPP
#define SLAMSHIM_CORE_HPP
#include <cassert>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {
typedef unsigned char uchar;
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
class KeyPoint {
public:
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    Point2f pt; float size, angle, response; int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(CV_32F), step_(0), data_(0) {}
    Mat(int r, int c, int t) : rows(r), cols(c), type_(t), step_((size_t)c * (t == CV_32F ? 4 : 1)),
        buf_(new std::vector<uchar>((size_t)r * c * (t == CV_32F ? 4 : 1) + 1, 0)), data_(buf_->data()) {}
    Mat(int r, int c, int t, void* ext, size_t step) : rows(r), cols(c), type_(t), step_(step), data_(static_cast<uchar*>(ext)) {}   /* external data */
    static Mat ones(int r, int c, int t) { Mat m(r, c, t); for (int y = 0; y < r; y++) for (int x = 0; x < c; x++) m.at<float>(y, x) = 1.f; return m; }
    void convertTo(Mat& dst, int t) const                 /* CV_8U -> CV_32F only; dst may be *this */
    {
        assert(type_ == CV_8U && t == CV_32F);
        Mat m(rows, cols, CV_32F);
        for (int y = 0; y < rows; y++) for (int x = 0; x < cols; x++) m.at<float>(y, x) = (float)at<uchar>(y, x);
        dst = m;
    }
    bool empty() const { return rows == 0 || cols == 0; }
    int type() const { return type_; }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data_ + (size_t)r * step_ + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data_ + (size_t)r * step_ + (size_t)c * sizeof(T)); }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data_ + (size_t)r * step_); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data_ + (size_t)r * step_); }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data_ = data_ + (size_t)a * step_; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data_ = data_ + (size_t)a * esz(); m.cols = b - a; return m; }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.data_ + (size_t)r * m.step_, data_ + (size_t)r * step_, (size_t)cols * esz()); return m; }
    Mat t() const { Mat m(cols, rows, type_); for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m.at<float>(c, r) = at<float>(r, c); return m; }
    double dot(const Mat& o) const                       /* dotProd_32f scalar tail: f64 accumulation in element order */
    {
        double s = 0.0;
        for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) s += (double)at<float>(r, c) * (double)o.at<float>(r, c);
        return s;
    }
private:
    size_t esz() const { return type_ == CV_32F ? 4 : 1; }
    int type_; size_t step_;
    std::shared_ptr<std::vector<uchar> > buf_;
    uchar* data_;
};

inline Mat operator*(const Mat& a, const Mat& b)          /* gemm, CV_32F: f32 products summed left to right */
{
    assert(a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_32F);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < b.cols; c++) {
            float s = a.at<float>(r, 0) * b.at<float>(0, c);
            for (int k = 1; k < a.cols; k++) s = s + a.at<float>(r, k) * b.at<float>(k, c);
            m.at<float>(r, c) = s;
        }
    return m;
}
#define SLAMSHIM_EW(OP)                                                                                            \
    inline Mat operator OP(const Mat& a, const Mat& b)                                                             \
    {                                                                                                              \
        Mat m(a.rows, a.cols, CV_32F);                                                                             \
        for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.at<float>(r, c) OP b.at<float>(r, c); \
        return m;                                                                                                  \
    }
SLAMSHIM_EW(+)
SLAMSHIM_EW(-)
#undef SLAMSHIM_EW
/* scaled forms: OpenCV evaluates `s*A`, `A/s`, `-A` as convertTo / gemm with a DOUBLE scale factor: f32(f64(a) * s) */
inline Mat scaled(const Mat& a, double s)
{
    Mat m(a.rows, a.cols, CV_32F);
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = (float)((double)a.at<float>(r, c) * s);
    return m;
}
inline Mat operator-(const Mat& a) { return scaled(a, -1.0); }
inline Mat operator*(double s, const Mat& a) { return scaled(a, s); }
inline Mat operator*(const Mat& a, double s) { return scaled(a, s); }
inline Mat operator/(const Mat& a, double s) { return scaled(a, 1.0 / s); }
inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }
enum { NORM_L1 = 2 };
inline double norm(const Mat& a, const Mat& b, int t)     /* normDiffL1_32f: |a - b| accumulated in f64 */
{
    assert(t == NORM_L1);
    double s = 0.0;
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) s += std::fabs((double)a.at<float>(r, c) - (double)b.at<float>(r, c));
    return s;
}
/* cv::FileStorage / cv::FileNode: only so that the virtual save / load members of DBoW2's TemplatedVocabulary compile
 * (bow_glue.cc loads the vocabulary through the reference's own text loader, never through these). */
class FileNode {
public:
    enum { SEQ = 5 };
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    int type() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](const std::string&) const { return FileNode(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
} // namespace cv
#endif
