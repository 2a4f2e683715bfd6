/* Minimal stand-in for the OpenCV core types that cross the ORBextractor / ORBmatcher / Frame interfaces.
 * The build container has no OpenCV C++ headers; the adapter classes compile against the real
 * <opencv2/core/core.hpp> when it is present (define SDYN_HAVE_OPENCV) and against this stub otherwise.
 * Only what those interfaces touch is provided: Mat (8-bit / 32-bit float, ROI views, shared ownership),
 * KeyPoint (28-byte layout of cv::KeyPoint), Point_, Rect_, Size, _InputArray / _OutputArray. */
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 0
#define CV_32FC1 5

namespace cv {

typedef unsigned char uchar;

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    Point_& operator*=(T s) { x *= s; y *= s; return *this; }
};
typedef Point_<int> Point2i; typedef Point2i Point; typedef Point_<float> Point2f; typedef Point_<double> Point2d;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
    template <class U> bool contains(const Point_<U>& p) const
    { return x <= (T)p.x && (T)p.x < x + width && y <= (T)p.y && (T)p.y < y + height; }
    T area() const { return width * height; }
};
typedef Rect_<int> Rect; typedef Rect_<double> Rect2d; typedef Rect_<float> Rect2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };

struct KeyPoint {          /* same field order and size as cv::KeyPoint */
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class Mat {
public:
    int rows = 0, cols = 0, flags = 0;
    uchar* data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), flags(type), data((uchar*)ext),
        step(st ? st : (size_t)c * elem(type)) {}
    void create(int r, int c, int type)
    {
        if (r == rows && c == cols && type == flags && data && step == (size_t)c * elem(type)) return;
        rows = r; cols = c; flags = type; step = (size_t)c * elem(type);
        own.reset(new uchar[std::max<size_t>((size_t)r * step, 1)], std::default_delete<uchar[]>());
        data = own.get();
    }
    void release() { own.reset(); data = nullptr; rows = cols = 0; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return flags; }
    bool isContinuous() const { return step == (size_t)cols * elem(flags); }
    Mat operator()(const Rect& r) const
    { Mat m(*this); m.data = data + (size_t)r.y * step + (size_t)r.x * elem(flags); m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat clone() const
    { Mat m(rows, cols, flags); for (int r = 0; r < rows; ++r) std::memcpy(m.data + r * m.step, data + r * step, (size_t)cols * elem(flags)); return m; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
private:
    static size_t elem(int type) { return type == CV_32F ? 4 : 1; }
    std::shared_ptr<uchar> own;
};

class _InputArray {
public:
    _InputArray() : m(nullptr) {}
    _InputArray(const Mat& mat) : m(&mat) {}
    Mat getMat() const { return m ? *m : Mat(); }
    bool empty() const { return !m || m->empty(); }
protected:
    const Mat* m;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& mat) : _InputArray(mat), out(&mat) {}
    void create(int r, int c, int type) const { out->create(r, c, type); }
    void release() const { out->release(); }
    Mat getMat() const { return *out; }
private:
    Mat* out;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
