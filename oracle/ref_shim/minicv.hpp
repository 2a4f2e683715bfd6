/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * minicv: the slice of the OpenCV C++ API that the reference's own hot-path translation units use
 * (src/ORBextractor.cc, src/ORBmatcher.cc, src/Frame.cc, src/Tracking.cc:1093-1367), written from scratch so those
 * files can be compiled UNCHANGED here (oracle/Makefile target `_ref`), where no OpenCV C++ exists.
 *
 * What is a faithful container and what is restated arithmetic:
 *   - cv::Mat / Point / Rect / KeyPoint / InputArray ...: containers with OpenCV's observable semantics
 *     (reference-counted buffers, ROI views, row/col views, push_back, reshape).
 *   - resize / copyMakeBorder / GaussianBlur / FAST / fastAtan2 / undistortPoints forward to oracle/orc_prims.cpp,
 *     the restatement that tests/test_oracle_prims.py pins bit-for-bit against cv2 4.13.
 *   - the matrix expressions the reference writes (A*B+C, -A.t()*b, s*A, A/s, norm, dot, inv) follow cv::MatExpr's
 *     fusion rules and OpenCV's gemm / convertTo / dotProd / invert arithmetic for CV_32F (small-matrix float path
 *     for 2..4-length products without transposes, double accumulation otherwise); tests/test_oracle_ref.py checks
 *     these against cv2.gemm / cv2.invert / cv2.norm through the `ref_cv_*` probes.
 *   - drawing / file output (rectangle, putText, drawMatches, imwrite) are no-ops: debug side effects of the
 *     reference.
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_PI 3.1415926535897932384626433832795
#define CV_GRAY2BGR 8
#define CV_RANSAC 8

/* cvRound = round-half-to-even (cvtsd2si / cvtss2si), cvFloor, cvCeil: SURVEY A-6 */
static inline int cvRound(double v) { return (int)std::lrint(v); }
static inline int cvRound(float v) { return (int)std::lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U>& o);
    Point_& operator*=(double s) { x = T(x * s); y = T(y * s); return *this; }
    Point_& operator*=(float s) { x = T(x * s); y = T(y * s); return *this; }
    Point_& operator*=(int s) { x = T(x * s); y = T(y * s); return *this; }
};
template <class T, class U> struct PtCast { static T go(U v) { return T(v); } };
template <class U> struct PtCast<int, U> { static int go(U v) { return cvRound(v); } };   /* saturate_cast<int>(float) rounds */
template <> struct PtCast<int, int> { static int go(int v) { return v; } };
template <class T> template <class U> Point_<T>::Point_(const Point_<U>& o) : x(PtCast<T, U>::go(o.x)), y(PtCast<T, U>::go(o.y)) {}
template <class T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
};
typedef Size_<int> Size;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
    Rect_(const Point_<T>& a, const Point_<T>& b)
    {
        x = std::min(a.x, b.x); y = std::min(a.y, b.y);
        width = std::max(a.x, b.x) - x; height = std::max(a.y, b.y) - y;
    }
    template <class U> operator Rect_<U>() const { return Rect_<U>(PtCast<U, T>::go(x), PtCast<U, T>::go(y), PtCast<U, T>::go(width), PtCast<U, T>::go(height)); }
    T area() const { return width * height; }
    bool empty() const { return width <= 0 || height <= 0; }
    /* half-open on both axes: x <= px < x+w (types.hpp) */
    template <class U> bool contains(const Point_<U>& p) const { return x <= p.x && p.x < x + width && y <= p.y && p.y < y + height; }
};
template <class T> Rect_<T> operator&(const Rect_<T>& a, const Rect_<T>& b)
{
    Rect_<T> r;
    T x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    r.width = std::min(a.x + a.width, b.x + b.width) - x1;
    r.height = std::min(a.y + a.height, b.y + b.height) - y1;
    r.x = x1; r.y = y1;
    if (r.width <= 0 || r.height <= 0) r = Rect_<T>();
    return r;
}
template <class T> Rect_<T> operator|(const Rect_<T>& a, const Rect_<T>& b)
{
    if (a.empty()) return b;
    if (b.empty()) return a;
    Rect_<T> r;
    T x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y);
    r.width = std::max(a.x + a.width, b.x + b.width) - x1;
    r.height = std::max(a.y + a.height, b.y + b.height) - y1;
    r.x = x1; r.y = y1;
    return r;
}
template <class T> Rect_<T> operator+(const Rect_<T>& a, const Point_<T>& b) { return Rect_<T>(a.x + b.x, a.y + b.y, a.width, a.height); }
typedef Rect_<int> Rect;
typedef Rect_<float> Rect2f;
typedef Rect_<double> Rect2d;

struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(INT32_MIN, INT32_MAX); }
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4028235e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4, NORM_HAMMING = 6 };
enum { COLOR_GRAY2BGR = 8 };

struct MatStep {
    size_t v;
    MatStep(size_t s = 0) : v(s) {}
    operator size_t() const { return v; }
};

class Mat;
struct MatT;      /* A.t() (optionally scaled) before evaluation */
struct MatMul;    /* alpha * op(A) * B (+ C) before evaluation */
struct MatScaled; /* alpha * A before evaluation */
struct MatInit;   /* Mat::zeros / ones / eye (optionally scaled) before evaluation */

class Mat {
public:
    int rows, cols;
    uchar* data;
    MatStep step;
    int dims;

    Mat() : rows(0), cols(0), data(nullptr), step(0), dims(0), type_(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); setTo(s.val[0]); }
    /* user-data view (no ownership) */
    Mat(int r, int c, int type, void* p, size_t stp = 0) : rows(r), cols(c), data((uchar*)p), step(0), dims(2), type_(type)
    {
        step = stp ? stp : (size_t)c * elemSize();
    }
    Mat(const Mat& m, const Rect& roi) : rows(roi.height), cols(roi.width), data(m.data + roi.y * (size_t)m.step + roi.x * m.elemSize()),
                                          step(m.step), dims(2), type_(m.type_), buf_(m.buf_)
    {
        assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= m.cols && roi.y + roi.height <= m.rows);
    }
    Mat(const MatT& e);
    Mat(const MatMul& e);
    Mat(const MatScaled& e);
    Mat(const MatInit& e);
    Mat& operator=(const MatInit& e);   /* evaluates INTO the existing buffer when size and type agree, like cv::MatExpr */
    Mat& operator=(const MatT& e);
    Mat& operator=(const MatMul& e);
    Mat& operator=(const MatScaled& e);

    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && type_ == type) return;
        rows = r; cols = c; type_ = type; dims = 2;
        step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<std::vector<uchar>>(new std::vector<uchar>((size_t)r * step.v + 64));
        data = r * c ? buf_->data() : nullptr;
        if (!data) buf_.reset();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { rows = cols = 0; data = nullptr; step = 0; buf_.reset(); dims = 0; }

    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize1() const { static const int sz[] = {1, 1, 2, 2, 4, 4, 8, 2}; return sz[depth()]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step.v / elemSize1(); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step.v == (size_t)cols * elemSize(); }

    uchar* ptr(int i = 0) { return data + (size_t)i * step.v; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step.v; }
    template <class T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step.v); }
    template <class T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step.v); }
    template <class T> T& at(int i, int j) { return ((T*)(data + (size_t)i * step.v))[j]; }
    template <class T> const T& at(int i, int j) const { return ((const T*)(data + (size_t)i * step.v))[j]; }
    template <class T> T& at(int i)
    {
        if (isContinuous() || rows == 1) return ((T*)data)[i];
        if (cols == 1) return *(T*)(data + (size_t)i * step.v);
        int r = i / cols;
        return ((T*)(data + (size_t)r * step.v))[i - r * cols];
    }
    template <class T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }

    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat operator()(Range rr, Range cr) const
    {
        if (rr.start == INT32_MIN) rr = Range(0, rows);
        if (cr.start == INT32_MIN) cr = Range(0, cols);
        return Mat(*this, Rect(cr.start, rr.start, cr.end - cr.start, rr.end - rr.start));
    }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat row(int i) const { return rowRange(i, i + 1); }
    Mat col(int j) const { return colRange(j, j + 1); }

    /* copyTo(OutputArray): a destination of the right size and type (e.g. a ROI view, even a temporary) is written in place */
    void copyTo(const Mat& dstView) const
    {
        Mat& dst = const_cast<Mat&>(dstView);
        if (empty()) { dst.release(); return; }
        if (dst.data == data && dst.rows == rows && dst.cols == cols) return;
        dst.create(rows, cols, type_);
        const size_t len = (size_t)cols * elemSize();
        for (int i = 0; i < rows; ++i) std::memcpy(dst.ptr(i), ptr(i), len);
    }
    Mat clone() const { Mat m; copyTo(m); return m; }
    void setTo(double v)
    {
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < cols * channels(); ++j) setElem(i, j, v);
    }
    double getElem(int i, int j) const
    {
        const uchar* p = ptr(i);
        switch (depth()) {
        case CV_8U: return p[j];
        case CV_32S: return ((const int*)p)[j];
        case CV_32F: return ((const float*)p)[j];
        case CV_64F: return ((const double*)p)[j];
        default: assert(!"minicv: depth"); return 0;
        }
    }
    void setElem(int i, int j, double v)
    {
        uchar* p = ptr(i);
        switch (depth()) {
        case CV_8U: { int r = cvRound(v); p[j] = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r); break; }
        case CV_32S: ((int*)p)[j] = cvRound(v); break;
        case CV_32F: ((float*)p)[j] = (float)v; break;
        case CV_64F: ((double*)p)[j] = v; break;
        default: assert(!"minicv: depth");
        }
    }
    /* Mat::convertTo: dst = saturate_cast<rtype>(src*alpha + beta).  8U -> 32F with alpha 1 is exact; 32F -> 32F scales in float
     * (cvtScale's working type for float is float). */
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const
    {
        const int dd = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
        Mat out(rows, cols, CV_MAKETYPE(dd, channels()));
        const int n = cols * channels();
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < n; ++j) {
                if (depth() == CV_32F && dd == CV_32F) {
                    const float s = ((const float*)ptr(i))[j];
                    ((float*)out.ptr(i))[j] = (alpha == 1 && beta == 0) ? s : s * (float)alpha + (float)beta;
                } else if (depth() == CV_64F && dd == CV_32F) {
                    ((float*)out.ptr(i))[j] = (float)(((const double*)ptr(i))[j] * alpha + beta);
                } else {
                    out.setElem(i, j, getElem(i, j) * alpha + beta);
                }
            }
        dst = out;
    }
    Mat reshape(int cn, int /*rows*/ = 0) const
    {
        assert(isContinuous() || rows == 1);
        Mat m = *this;
        const int totalCh = cols * channels();
        assert(totalCh % cn == 0);
        m.cols = totalCh / cn;
        m.type_ = CV_MAKETYPE(depth(), cn);
        return m;
    }
    /* Mat::push_back(const Mat&): append rows; an empty matrix becomes a copy of the argument. */
    void push_back(const Mat& m)
    {
        if (m.empty()) return;
        if (empty()) { *this = m.clone(); return; }
        assert(m.type_ == type_ && m.cols == cols);
        Mat out(rows + m.rows, cols, type_);
        const size_t len = (size_t)cols * elemSize();
        for (int i = 0; i < rows; ++i) std::memcpy(out.ptr(i), ptr(i), len);
        for (int i = 0; i < m.rows; ++i) std::memcpy(out.ptr(rows + i), m.ptr(i), len);
        *this = out;
    }

    static MatInit zeros(int r, int c, int type);
    static MatInit ones(int r, int c, int type);
    static MatInit eye(int r, int c, int type);

    MatT t() const;
    Mat inv(int method = 0) const;
    double dot(const Mat& m) const;

    int type_;
    std::shared_ptr<std::vector<uchar>> buf_;
};

template <class T> struct DepthOf;
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };
template <> struct DepthOf<int> { enum { value = CV_32S }; };
template <> struct DepthOf<uchar> { enum { value = CV_8U }; };

template <class T> struct MatCommaInitializer_ {
    Mat m; int idx;
    MatCommaInitializer_(const Mat& m_, T v) : m(m_), idx(0) { put(v); }
    void put(T v) { m.at<T>(idx / m.cols, idx % m.cols) = v; ++idx; }
    template <class U> MatCommaInitializer_& operator,(U v) { put((T)v); return *this; }
    operator Mat() const { return m; }
};
template <class T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, DepthOf<T>::value) {}
    T& operator()(int i, int j) { return at<T>(i, j); }
};
template <class T, class U> MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) { return MatCommaInitializer_<T>(m, (T)v); }

/* ---- matrix expressions (cv::MatExpr fusion rules for the forms the reference writes) ---- */
struct MatT { Mat a; double alpha; };
struct MatScaled { Mat a; double alpha; };
struct MatMul { Mat a, b, c; double alpha, beta; int flags; /* bit0: A transposed */ };
struct MatInit { int rows, cols, type; double value; bool eye; };
inline MatInit Mat::zeros(int r, int c, int type) { return MatInit{r, c, type, 0.0, false}; }
inline MatInit Mat::ones(int r, int c, int type) { return MatInit{r, c, type, 1.0, false}; }
inline MatInit Mat::eye(int r, int c, int type) { return MatInit{r, c, type, 1.0, true}; }
inline Mat& Mat::operator=(const MatInit& e)
{
    create(e.rows, e.cols, e.type);
    if (e.eye) { setTo(0); for (int i = 0; i < std::min(rows, cols); ++i) setElem(i, i, e.value); }
    else setTo(e.value);
    return *this;
}
inline Mat::Mat(const MatInit& e) : Mat() { *this = e; }
inline MatInit operator*(double s, const MatInit& e) { MatInit r = e; r.value *= s; return r; }
inline MatInit operator*(const MatInit& e, double s) { MatInit r = e; r.value *= s; return r; }

Mat minicv_gemm(const Mat& a, const Mat& b, double alpha, const Mat& c, double beta, int flags);
Mat minicv_transpose(const Mat& a);
Mat minicv_addsub(const Mat& a, const Mat& b, int sign);

inline MatT Mat::t() const { return MatT{*this, 1.0}; }
inline Mat::Mat(const MatT& e) : Mat() { *this = e; }
inline Mat::Mat(const MatMul& e) : Mat() { *this = e; }
inline Mat::Mat(const MatScaled& e) : Mat() { *this = e; }
inline Mat& Mat::operator=(const MatT& e)
{
    Mat r = minicv_transpose(e.a);
    if (e.alpha != 1) r.convertTo(r, -1, e.alpha);     /* MatOp_T::assign */
    return *this = r;
}
inline Mat& Mat::operator=(const MatMul& e) { return *this = minicv_gemm(e.a, e.b, e.alpha, e.c, e.beta, e.flags); }
inline Mat& Mat::operator=(const MatScaled& e) { Mat r; e.a.convertTo(r, -1, e.alpha); return *this = r; }   /* MatOp_AddEx::assign */

inline MatT operator-(const MatT& e) { return MatT{e.a, -e.alpha}; }
inline MatT operator*(double s, const MatT& e) { return MatT{e.a, e.alpha * s}; }
inline MatMul operator*(const MatT& e, const Mat& b) { return MatMul{e.a, b, Mat(), e.alpha, 0, 1}; }
inline MatMul operator*(const Mat& a, const Mat& b) { return MatMul{a, b, Mat(), 1, 0, 0}; }
inline MatMul operator*(const MatScaled& e, const Mat& b) { return MatMul{e.a, b, Mat(), e.alpha, 0, 0}; }
inline MatMul operator+(const MatMul& e, const Mat& c) { MatMul r = e; assert(r.c.empty()); r.c = c; r.beta = 1; return r; }
inline MatScaled operator-(const Mat& a) { return MatScaled{a, -1.0}; }
inline MatScaled operator*(double s, const Mat& a) { return MatScaled{a, s}; }
inline MatScaled operator*(const Mat& a, double s) { return MatScaled{a, s}; }
inline MatScaled operator/(const Mat& a, double s) { return MatScaled{a, 1.0 / s}; }
inline Mat operator-(const Mat& a, const Mat& b) { return minicv_addsub(a, b, -1); }
inline Mat operator+(const Mat& a, const Mat& b) { return minicv_addsub(a, b, +1); }
inline Mat operator-(const Mat& a, const MatScaled& b) { return minicv_addsub(a, Mat(b), -1); }
inline Mat operator-(const Mat& a, const MatInit& b) { return minicv_addsub(a, Mat(b), -1); }

double norm(const Mat& a, int normType = NORM_L2);
double norm(const Mat& a, const Mat& b, int normType = NORM_L2);

/* ---- array proxies ---- */
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size s, int type) const { m_->create(s, type); }
    void release() const { if (m_) m_->release(); }
    Mat& getMatRef() const { return *m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
typedef const _OutputArray& InputOutputArray;
inline InputArray noArray() { static _InputArray none; return none; }

/* ---- imgproc / features2d / calib3d primitives (forward to oracle/orc_prims.cpp) ---- */
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType, const Scalar& value = Scalar());
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
float fastAtan2(float y, float x);
void undistortPoints(InputArray src, OutputArray dst, InputArray cameraMatrix, InputArray distCoeffs, InputArray R = noArray(), InputArray P = noArray());
void cvtColor(InputArray src, OutputArray dst, int code, int dstCn = 0);
void vconcat(InputArray a, InputArray b, OutputArray dst);

struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

class BFMatcher {
public:
    BFMatcher(int normType = NORM_L2, bool crossCheck = false) : normType_(normType), crossCheck_(crossCheck) {}
    void match(InputArray queryDescriptors, InputArray trainDescriptors, std::vector<DMatch>& matches, InputArray mask = noArray()) const;
private:
    int normType_;
    bool crossCheck_;
};

/* debug side effects of the reference: not reproduced */
template <class R> inline void rectangle(Mat&, const R&, const Scalar&, int = 1, int = 8, int = 0) {}
template <class P> inline void putText(Mat&, const std::string&, P, int, double, Scalar, int = 1, int = 8, bool = false) {}
inline void drawMatches(InputArray, const std::vector<KeyPoint>&, InputArray, const std::vector<KeyPoint>&, const std::vector<DMatch>&, Mat&) {}
inline bool imwrite(const std::string&, InputArray) { return true; }

}  // namespace cv
