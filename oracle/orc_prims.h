/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the arithmetic the reference front end relies on.  Nothing in the
 * product path (slam-dynamic_b200/) may include, link or call this directory; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * This header: the OpenCV primitives the reference calls but does not vendor
 * (SURVEY.md Appendix A).  OpenCV is an un-vendored, un-pinned dependency of the
 * reference (CMakeLists.txt:33-39); the arithmetic below restates OpenCV's published
 * algorithms and is pinned bit-for-bit against cv2 4.13.0 by tests/test_oracle_prims.py
 * and the committed fixtures in tests/golden/.
 */
#pragma once
#include <cstdint>
#include <cmath>

namespace orc {

/* cvRound / cvFloor / cvCeil (A-6): cvRound is round-half-to-even. */
static inline int cv_round(double v) { return (int)std::lrint(v); }
static inline int cv_round(float v) { return (int)std::lrintf(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

/* BORDER_REFLECT_101 index map (A-1), valid for -n < i < 2n-1. */
static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for 8UC1 (A-2). */
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride);

/* cv::copyMakeBorder(..., b,b,b,b, BORDER_REFLECT_101): dst is (w+2b) x (h+2b); when
 * `in_place` the interior already lives at dst+b*dstride+b and only the frame is written
 * (the BORDER_ISOLATED in-place call of ORBextractor.cc:1122). */
void border_reflect101(const uint8_t* src, int w, int h, int sstride,
                       uint8_t* dst, int b, int dstride, bool in_place);

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for 8UC1 (A-3). */
void gaussian_blur7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);

/* FAST-9/16 corner score V(p) (A-4) of the pixel at `p` (row stride `stride`). */
int fast_score(const uint8_t* p, int stride);

/* cv::FAST(img, kps, th, true): appends (x, y, response) triples in raster order; returns count. */
int fast_nms(const uint8_t* img, int w, int h, int stride, int th, int* xyv, int cap);

/* cv::fastAtan2 (A-5), degrees in [0,360). */
float fast_atan2(float y, float x);

}  // namespace orc

namespace orc {
/* cv::undistortPoints(src, dst, K, distCoeffs, noArray(), K) as Frame::UndistortKeyPoints / ComputeImageBounds call
 * it (src/Frame.cc:812-872): 2-channel float points, float K = (fx, fy, cx, cy), float distortion (k1, k2, p1, p2[, k3]),
 * default termination = 5 fixed-point iterations, all arithmetic in double (OpenCV calib3d, un-vendored; pinned
 * against cv2 4.13 by tests/test_oracle_prims.py). */
void undistort_points(const float* srcXY, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist, float* dstXY);
}
