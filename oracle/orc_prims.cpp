/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * Restates the OpenCV primitives of SURVEY.md Appendix A.  Build with -ffp-contract=off. */
#include "orc_prims.h"
#include <vector>
#include <algorithm>
#include <cfloat>
#include <cstring>

namespace orc {

static inline short sat_short(int v) { return (short)std::min(std::max(v, -32768), 32767); }
static inline int clip_idx(int x, int n) { return x < 0 ? 0 : (x < n ? x : n - 1); }

/* A-2.  Follows OpenCV's 8-bit INTER_LINEAR path: 11-bit coefficients, horizontal pass in
 * int32, vertical pass ((b*(H>>4))>>16 summed, +2, >>2).  Called at ORBextractor.cc:1120. */
void resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                      uint8_t* dst, int dw, int dh, int dstride)
{
    const double scale_x = 1.0 / ((double)dw / (double)sw);
    const double scale_y = 1.0 / ((double)dh / (double)sh);
    std::vector<int> xo(dw), yo(dh);
    std::vector<short> ca(2 * (size_t)dw), cb(2 * (size_t)dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { sx = 0; fx = 0.f; }
        if (sx >= sw - 1) { sx = sw - 1; fx = 0.f; }
        xo[dx] = sx;
        ca[2 * dx] = sat_short(cv_round((1.f - fx) * 2048.f));
        ca[2 * dx + 1] = sat_short(cv_round(fx * 2048.f));
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yo[dy] = sy;
        cb[2 * dy] = sat_short(cv_round((1.f - fy) * 2048.f));
        cb[2 * dy + 1] = sat_short(cv_round(fy * 2048.f));
    }
    std::vector<int> h0(dw), h1(dw);
    auto hpass = [&](int sy, std::vector<int>& out) {
        const uint8_t* s = src + (size_t)sy * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xo[dx];
            int nx = std::min(sx + 1, sw - 1);
            out[dx] = s[sx] * ca[2 * dx] + s[nx] * ca[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        hpass(clip_idx(yo[dy], sh), h0);
        hpass(clip_idx(yo[dy] + 1, sh), h1);
        const int b0 = cb[2 * dy], b1 = cb[2 * dy + 1];
        uint8_t* d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (h0[dx] >> 4)) >> 16) + ((b1 * (h1[dx] >> 4)) >> 16) + 2) >> 2;
            d[dx] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

/* A-1.  ORBextractor.cc:1122,1127. */
void border_reflect101(const uint8_t* src, int w, int h, int sstride,
                       uint8_t* dst, int b, int dstride, bool in_place)
{
    for (int y = -b; y < h + b; ++y) {
        const int sy = reflect101(y, h);
        uint8_t* d = dst + (size_t)(y + b) * dstride;
        const uint8_t* s = in_place ? dst + (size_t)(sy + b) * dstride + b : src + (size_t)sy * sstride;
        const bool inner_row = (y >= 0 && y < h);
        for (int x = -b; x < w + b; ++x) {
            if (in_place && inner_row && x >= 0 && x < w) continue;
            d[x + b] = s[reflect101(x, w)];
        }
    }
}

/* A-3.  8.8 fixed-point taps (sum 256); exact integer separable passes; ORBextractor.cc:1086. */
void gaussian_blur7_s2(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<int> t((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstride;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int i = 0; i < 7; ++i) acc += k[i] * s[reflect101(x + i - 3, w)];
            t[(size_t)y * w + x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int j = 0; j < 7; ++j) acc += k[j] * t[(size_t)reflect101(y + j - 3, h) * w + x];
            d[x] = (uint8_t)((acc + 32768) >> 16);
        }
    }
}

/* A-4.  Bresenham ring of radius 3, clockwise from (0,3). */
static const int kRing[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                 {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

int fast_score(const uint8_t* p, int stride)
{
    int d[25];
    const int v = p[0];
    for (int i = 0; i < 16; ++i) d[i] = v - p[kRing[i][1] * stride + kRing[i][0]];
    for (int i = 16; i < 25; ++i) d[i] = d[i - 16];
    int best = -256;
    for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
        for (int i = 1; i < 9; ++i) { mn = std::min(mn, d[s + i]); mx = std::max(mx, d[s + i]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

int fast_nms(const uint8_t* img, int w, int h, int stride, int th, int* xyv, int cap)
{
    if (w < 7 || h < 7) return 0;
    th = std::min(std::max(th, 0), 255);
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int v = fast_score(img + (size_t)y * stride + x, stride);
            if (v >= th) sc[(size_t)y * w + x] = v;
        }
    int n = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = sc[(size_t)y * w + x];
            if (s <= 0) continue;  /* non-corner (or score 0, never a strict maximum) */
            bool keep = true;
            for (int dy = -1; dy <= 1 && keep; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    if (!(s > sc[(size_t)(y + dy) * w + x + dx])) { keep = false; break; }
                }
            if (!keep) continue;
            if (n < cap) { xyv[3 * n] = x; xyv[3 * n + 1] = y; xyv[3 * n + 2] = s; }
            ++n;
        }
    return n;
}

/* A-5.  Constants are float(c)*float(180/pi) multiplied in float; every op rounds to float. */
float fast_atan2(float y, float x)
{
    union { uint32_t u; float f; } p1 = {0x4265226fu}, p3 = {0xc19556eeu}, p5 = {0x410e9fbfu}, p7 = {0xc0228ad9u};
    const float eps = (float)DBL_EPSILON;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7.f * c2 + p5.f) * c2 + p3.f) * c2 + p1.f) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7.f * c2 + p5.f) * c2 + p3.f) * c2 + p1.f) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

}  // namespace orc

namespace orc {
void undistort_points(const float* srcXY, int n, float fxf, float fyf, float cxf, float cyf, const float* dist, int ndist, float* dstXY)
{
    double k[5] = {0, 0, 0, 0, 0};                           /* k1 k2 p1 p2 k3 */
    for (int i = 0; i < ndist && i < 5; ++i) k[i] = dist[i];
    const double fx = fxf, fy = fyf, cx = cxf, cy = cyf;
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        double x = srcXY[2 * i], y = srcXY[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx; y = (y - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            /* rational coefficients k4..k6 are zero: numerator 1 + ((0*r2 + 0)*r2 + 0)*r2 = 1 */
            const double icdist = (1 + ((0. * r2 + 0.) * r2 + 0.) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + 0. * r2 + 0. * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + 0. * r2 + 0. * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        /* R = I, P = K: RR = K; xx = fx*x + 0*y + cx, ww = 1/(0*x + 0*y + 1) */
        const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        dstXY[2 * i] = (float)(xx * ww); dstXY[2 * i + 1] = (float)(yy * ww);
    }
}
}  // namespace orc
