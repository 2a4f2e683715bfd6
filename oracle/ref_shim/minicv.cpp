/* ORACLE — TEST INFRASTRUCTURE ONLY.  Out-of-line part of minicv (see minicv.hpp). */
#include "minicv.hpp"
#include "../orc_prims.h"

namespace cv {

/* cv::gemm for CV_32F as matmul.dispatch.cpp evaluates it:
 *   flags == 0, 2 <= len <= 4 and len == D.cols or len == D.rows  ->  float products summed left to right in float,
 *       then d = (float)(t*alpha + c*beta) in double;
 *   otherwise (a transposed operand, longer products)             ->  products accumulated in double,
 *       d = (float)(s*alpha + c*beta). */
Mat minicv_gemm(const Mat& A0, const Mat& B, double alpha, const Mat& C, double beta, int flags)
{
    assert(A0.type() == CV_32F && B.type() == CV_32F);
    const bool aT = flags & 1;
    const int M = aT ? A0.cols : A0.rows, len = aT ? A0.rows : A0.cols, N = B.cols;
    assert(B.rows == len);
    assert(C.empty() || (C.rows == M && C.cols == N && C.type() == CV_32F));
    Mat D(M, N, CV_32F);
    const bool small = flags == 0 && len >= 2 && len <= 4 && (len == N || len == M);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            const double c = C.empty() ? 0.0 : (double)C.at<float>(i, j);
            if (small) {
                float t = A0.at<float>(i, 0) * B.at<float>(0, j);
                for (int k = 1; k < len; ++k) t = t + A0.at<float>(i, k) * B.at<float>(k, j);
                D.at<float>(i, j) = (float)(t * alpha + c * beta);
            } else {
                double s = 0;
                for (int k = 0; k < len; ++k) {
                    const float a = aT ? A0.at<float>(k, i) : A0.at<float>(i, k);
                    s += (double)a * (double)B.at<float>(k, j);
                }
                D.at<float>(i, j) = C.empty() ? (float)(s * alpha) : (float)(s * alpha + c * beta);
            }
        }
    return D;
}

Mat minicv_transpose(const Mat& a)
{
    Mat r(a.cols, a.rows, a.type());
    const size_t es = a.elemSize();
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) std::memcpy(r.ptr(j) + i * es, a.ptr(i) + j * es, es);
    return r;
}

Mat minicv_addsub(const Mat& a, const Mat& b, int sign)
{
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type() && a.type() == CV_32F);
    Mat r(a.rows, a.cols, a.type());
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j)
            r.at<float>(i, j) = sign > 0 ? a.at<float>(i, j) + b.at<float>(i, j) : a.at<float>(i, j) - b.at<float>(i, j);
    return r;
}

/* Mat::dot for CV_32F: products of the float elements accumulated in double (dotProd_<float>). */
double Mat::dot(const Mat& m) const
{
    assert(type() == CV_32F && m.type() == CV_32F && total() == m.total());
    double r = 0;
    const int n = (int)total();
    for (int i = 0; i < n; ++i) r += (double)at<float>(i) * (double)m.at<float>(i);
    return r;
}

/* cv::invert(DECOMP_LU) closed form for 3x3 CV_32F: cofactors and determinant in double, one rounding per element. */
Mat Mat::inv(int) const
{
    assert(rows == 3 && cols == 3 && type() == CV_32F);
    const Mat& S = *this;
#define Sf(y, x) ((double)S.at<float>(y, x))
    double d = Sf(0, 0) * (Sf(1, 1) * Sf(2, 2) - Sf(1, 2) * Sf(2, 1)) - Sf(0, 1) * (Sf(1, 0) * Sf(2, 2) - Sf(1, 2) * Sf(2, 0)) +
               Sf(0, 2) * (Sf(1, 0) * Sf(2, 1) - Sf(1, 1) * Sf(2, 0));
    Mat D(3, 3, CV_32F);
    D.setTo(0);
    if (d != 0.) {
        d = 1. / d;
        D.at<float>(0, 0) = (float)((Sf(1, 1) * Sf(2, 2) - Sf(1, 2) * Sf(2, 1)) * d);
        D.at<float>(0, 1) = (float)((Sf(0, 2) * Sf(2, 1) - Sf(0, 1) * Sf(2, 2)) * d);
        D.at<float>(0, 2) = (float)((Sf(0, 1) * Sf(1, 2) - Sf(0, 2) * Sf(1, 1)) * d);
        D.at<float>(1, 0) = (float)((Sf(1, 2) * Sf(2, 0) - Sf(1, 0) * Sf(2, 2)) * d);
        D.at<float>(1, 1) = (float)((Sf(0, 0) * Sf(2, 2) - Sf(0, 2) * Sf(2, 0)) * d);
        D.at<float>(1, 2) = (float)((Sf(0, 2) * Sf(1, 0) - Sf(0, 0) * Sf(1, 2)) * d);
        D.at<float>(2, 0) = (float)((Sf(1, 0) * Sf(2, 1) - Sf(1, 1) * Sf(2, 0)) * d);
        D.at<float>(2, 1) = (float)((Sf(0, 1) * Sf(2, 0) - Sf(0, 0) * Sf(2, 1)) * d);
        D.at<float>(2, 2) = (float)((Sf(0, 0) * Sf(1, 1) - Sf(0, 1) * Sf(1, 0)) * d);
    }
#undef Sf
    return D;
}

double norm(const Mat& a, int normType)
{
    assert(a.type() == CV_32F);
    double s = 0;
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) {
            const double v = a.at<float>(i, j);
            s += normType == NORM_L1 ? std::fabs(v) : v * v;
        }
    return normType == NORM_L1 ? s : std::sqrt(s);
}

double norm(const Mat& a, const Mat& b, int normType)
{
    assert(a.type() == CV_32F && b.type() == CV_32F && a.rows == b.rows && a.cols == b.cols);
    double s = 0;
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < a.cols; ++j) {
            const double v = a.at<float>(i, j) - b.at<float>(i, j);      /* float difference, double accumulation */
            s += normType == NORM_L1 ? std::fabs(v) : v * v;
        }
    return normType == NORM_L1 ? s : std::sqrt(s);
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double, double, int interpolation)
{
    Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && dsize.width > 0 && dsize.height > 0);
    (void)interpolation;
    dst_.create(dsize, src.type());
    Mat& dst = dst_.getMatRef();
    orc::resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar&)
{
    Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && (borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    assert(top == bottom && left == right && top == left);
    (void)bottom; (void)right; (void)borderType;
    dst_.create(src.rows + 2 * top, src.cols + 2 * left, src.type());
    Mat& dst = dst_.getMatRef();
    const bool inPlace = src.data == dst.data + (size_t)top * dst.step + left;
    orc::border_reflect101(src.data, src.cols, src.rows, (int)src.step, dst.data, top, (int)dst.step, inPlace);
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sx, double sy, int borderType)
{
    Mat src = src_.getMat();
    assert(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101);
    (void)ksize; (void)sx; (void)sy; (void)borderType;
    Mat tmp(src.rows, src.cols, src.type());
    orc::gaussian_blur7_s2(src.data, src.cols, src.rows, (int)src.step, tmp.data, (int)tmp.step);
    dst_.create(src.rows, src.cols, src.type());
    tmp.copyTo(dst_.getMatRef());
}

void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool nonmax)
{
    Mat img = image_.getMat();
    assert(img.type() == CV_8UC1 && nonmax);
    (void)nonmax;
    keypoints.clear();
    if (img.cols < 7 || img.rows < 7) return;
    std::vector<int> xyv((size_t)img.cols * img.rows * 3 / 4 + 16);
    const int n = orc::fast_nms(img.data, img.cols, img.rows, (int)img.step, threshold, xyv.data(), (int)(xyv.size() / 3));
    keypoints.reserve(n);
    for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint((float)xyv[3 * i], (float)xyv[3 * i + 1], 7.f, -1.f, (float)xyv[3 * i + 2]));
}

float fastAtan2(float y, float x) { return orc::fast_atan2(y, x); }

void undistortPoints(InputArray src_, OutputArray dst_, InputArray K_, InputArray dist_, InputArray, InputArray P_)
{
    Mat src = src_.getMat(), K = K_.getMat(), dist = dist_.getMat(), P = P_.getMat();
    assert(src.type() == CV_32FC2 && K.type() == CV_32F && dist.type() == CV_32F);
    assert(P.data == K.data || P.empty());    /* the reference passes P = K (Frame.cc:826,853) */
    const int n = (int)src.total();
    std::vector<float> in(2 * (size_t)n), out(2 * (size_t)n), dc(dist.total());
    for (int i = 0; i < n; ++i) { in[2 * i] = src.at<float>(i, 0); in[2 * i + 1] = src.at<float>(i, 1); }
    for (size_t i = 0; i < dc.size(); ++i) dc[i] = dist.at<float>((int)i);
    orc::undistort_points(in.data(), n, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), dc.data(), (int)dc.size(), out.data());
    dst_.create(src.rows, src.cols, src.type());
    Mat& dst = dst_.getMatRef();
    for (int i = 0; i < n; ++i) { dst.at<float>(i, 0) = out[2 * i]; dst.at<float>(i, 1) = out[2 * i + 1]; }
}

void cvtColor(InputArray src_, OutputArray dst_, int code, int)
{
    Mat src = src_.getMat();
    assert(code == CV_GRAY2BGR && src.type() == CV_8UC1);
    (void)code;
    Mat out(src.rows, src.cols, CV_8UC3);
    for (int i = 0; i < src.rows; ++i)
        for (int j = 0; j < src.cols; ++j) { const uchar v = src.at<uchar>(i, j); uchar* p = out.ptr(i) + 3 * j; p[0] = p[1] = p[2] = v; }
    dst_.getMatRef() = out;
}

void vconcat(InputArray a_, InputArray b_, OutputArray dst_)
{
    Mat a = a_.getMat(), b = b_.getMat();
    Mat out = a.clone();
    out.push_back(b);
    dst_.getMatRef() = out;
}

/* KeyPointsFilter::retainBest (only reachable from the dead ComputeKeyPointsOld): keep the npoints strongest, plus ties. */
void KeyPointsFilter::retainBest(std::vector<KeyPoint>& kps, int n)
{
    if (n < 0 || (int)kps.size() <= n) return;
    if (n == 0) { kps.clear(); return; }
    std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    const float edge = kps[n - 1].response;
    auto it = std::partition(kps.begin() + n, kps.end(), [edge](const KeyPoint& k) { return k.response >= edge; });
    kps.resize(it - kps.begin());
}

/* BFMatcher(NORM_HAMMING, crossCheck=true)::match = strict mutual nearest neighbour, lowest index on ties on both
 * sides, output ordered by query index (SURVEY A-7; pinned against cv2.BFMatcher in tests/test_oracle_matcher.py). */
void BFMatcher::match(InputArray q_, InputArray t_, std::vector<DMatch>& matches, InputArray) const
{
    Mat q = q_.getMat(), t = t_.getMat();
    assert(normType_ == NORM_HAMMING && crossCheck_);
    matches.clear();
    if (q.empty() || t.empty()) return;
    assert(q.type() == CV_8UC1 && t.type() == CV_8UC1 && q.cols == t.cols);
    const int nq = q.rows, nt = t.rows, len = q.cols;
    std::vector<int> bestT(nq, -1), bestDq(nq, INT32_MAX), bestQ(nt, -1), bestDt(nt, INT32_MAX);
    for (int i = 0; i < nq; ++i)
        for (int j = 0; j < nt; ++j) {
            int d = 0;
            const uchar *a = q.ptr(i), *b = t.ptr(j);
            for (int k = 0; k < len; ++k) d += __builtin_popcount((unsigned)(a[k] ^ b[k]));
            if (d < bestDq[i]) { bestDq[i] = d; bestT[i] = j; }
            if (d < bestDt[j]) { bestDt[j] = d; bestQ[j] = i; }
        }
    for (int i = 0; i < nq; ++i)
        if (bestT[i] >= 0 && bestQ[bestT[i]] == i) matches.push_back(DMatch(i, bestT[i], (float)bestDq[i]));
}

}  // namespace cv
