/* Host-side geometry: scale tables, per-level sizes, FAST cell grid, octree quotas and the fixed-point
 * bilinear tables.  Mirrors the arithmetic of the reference constructor and ComputePyramid exactly
 * (float/double mix included) so the device kernels reproduce the reference bit for bit.
 *   ORBextractor::ORBextractor        src/ORBextractor.cc:410-470
 *   ORBextractor::ComputePyramid      src/ORBextractor.cc:1107-1132
 *   ComputeKeyPointsOctTree (grid)    src/ORBextractor.cc:773-787
 *   DistributeOctTree (roots)         src/ORBextractor.cc:543-545
 *   cv::resize INTER_LINEAR tables    SURVEY.md Appendix A-2 (OpenCV, un-vendored dependency)
 */
#include "sdyn_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace sdyn {

static inline int round_half_even(float v) { return (int)std::lrintf(v); }      /* cvRound */
static inline int round_half_even(double v) { return (int)std::lrint(v); }
static inline int floor_int(double v) { int i = (int)v; return i - (i > v); }    /* cvFloor */
static inline int ceil_int(double v) { int i = (int)v; return i + (i < v); }     /* cvCeil */
static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void compute_scale_info(const sdyn_orb_params& p, sdyn_scale_info& s, int umax[16])
{
    std::memset(&s, 0, sizeof(s));
    const int n = p.nlevels;
    const double sf = (double)p.scale_factor;          /* the reference keeps scaleFactor in a double member */
    s.nlevels = n;
    s.scale[0] = 1.0f; s.sigma2[0] = 1.0f;
    for (int i = 1; i < n; ++i) {
        s.scale[i] = (float)(s.scale[i - 1] * sf);
        s.sigma2[i] = s.scale[i] * s.scale[i];
    }
    for (int i = 0; i < n; ++i) { s.inv_scale[i] = 1.0f / s.scale[i]; s.inv_sigma2[i] = 1.0f / s.sigma2[i]; }

    const float factor = (float)(1.0f / sf);
    float want = p.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; ++l) {
        s.features_per_level[l] = round_half_even(want);
        sum += s.features_per_level[l];
        want *= factor;
    }
    s.features_per_level[n - 1] = std::max(p.nfeatures - sum, 0);

    /* half-widths of the radius-15 orientation disc */
    const int R = 15;
    std::fill(umax, umax + 16, 0);
    const int vmax = floor_int(R * std::sqrt(2.f) / 2 + 1);
    const int vmin = ceil_int(R * std::sqrt(2.f) / 2);
    for (int v = 0; v <= vmax; ++v) umax[v] = round_half_even(std::sqrt((double)R * R - v * v));
    for (int v = R, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

static void level_size(const sdyn_scale_info& s, int W, int H, int l, int& w, int& h)
{
    const float inv = s.inv_scale[l];
    w = round_half_even((float)W * inv);
    h = round_half_even((float)H * inv);
}

static int node_cap(int quota, int nIni) { return std::max(quota + 3, 4 * nIni) + 4; }

int max_keypoints_per_frame(const sdyn_orb_params& p, const sdyn_scale_info& s, int maxW, int maxH)
{
    /* a level keeps at most max(quota+3, 4*nIni) keypoints; nIni depends on the aspect ratio, bound it
     * with the widest supported one */
    int total = 0;
    for (int l = 0; l < p.nlevels; ++l) {
        int w, h; level_size(s, maxW, maxH, l, w, h);
        int nIni = 1;
        if (h - 32 > 0) nIni = std::max(1, (int)std::round((float)(w - 32) / (float)(h - 32)));
        total += node_cap(s.features_per_level[l], std::max(nIni, 8));
    }
    return total;
}

int compute_geometry(const sdyn_orb_params& p, const sdyn_scale_info& s, int W, int H,
                     Geom& g, std::vector<uint8_t>& tables)
{
    std::memset(&g, 0, sizeof(g));
    g.nlevels = p.nlevels; g.W = W; g.H = H;
    tables.clear();
    size_t off = 0;
    int cells = 0, cand = 0, kp = 0, maxCap = 0;
    for (int l = 0; l < p.nlevels; ++l) {
        LevelGeom& L = g.L[l];
        level_size(s, W, H, l, L.w, L.h);
        if (L.w < 1 || L.h < 1) return SDYN_ERR_GEOMETRY;
        L.pitch = (int)align_up((size_t)kLeftPad + L.w + kEdge, kRowAlign);
        off = align_up(off, 256);
        L.off = (long long)(off + (size_t)kEdge * L.pitch + kLeftPad);
        off += (size_t)L.pitch * (L.h + 2 * kEdge);

        /* FAST window and 30-pixel cell grid */
        L.fw = L.w - 2 * kFastBorder;
        L.fh = L.h - 2 * kFastBorder;
        if (L.fw <= 0 || L.fh <= 0) return SDYN_ERR_GEOMETRY;
        const float width = (float)L.fw, height = (float)L.fh;
        L.nCols = (int)(width / 30.f);
        L.nRows = (int)(height / 30.f);
        if (L.nCols <= 0 || L.nRows <= 0) return SDYN_ERR_GEOMETRY;   /* reference: ceil(x/0) */
        L.wCell = (int)std::ceil(width / L.nCols);
        L.hCell = (int)std::ceil(height / L.nRows);
        if (L.wCell > 63 || L.hCell > 63 || L.fw >= 4096 || L.fh >= 4096) return SDYN_ERR_GEOMETRY;
        L.cellOff = cells;
        cells += L.nCols * L.nRows;
        L.candOff = cand;
        L.candCap = L.nCols * L.nRows * ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
        cand += L.candCap;

        /* octree roots */
        L.quota = s.features_per_level[l];
        L.nIni = (int)std::round(static_cast<float>(L.fw) / L.fh);
        if (L.nIni <= 0) return SDYN_ERR_GEOMETRY;                    /* reference indexes an empty vector */
        L.hX = static_cast<float>(L.fw) / L.nIni;
        L.nodeCap = node_cap(L.quota, L.nIni);
        maxCap = std::max(maxCap, L.nodeCap);
        L.kpOff = kp;
        kp += L.nodeCap;
        L.scale = s.scale[l];
        L.patchSize = (float)(int)(31 * s.scale[l]);

        /* bilinear tables, indexed by bordered destination coordinate */
        L.xtab = L.ytab = L.xTile = L.yTile = -1;
        if (l > 0) {
            const LevelGeom& P = g.L[l - 1];
            auto build = [&](int dn, int sn, bool clampCoef) {
                const double sc = 1.0 / ((double)dn / (double)sn);
                std::vector<ResizeTap> t((size_t)dn + 2 * kEdge);
                for (int b = 0; b < dn + 2 * kEdge; ++b) {
                    const int d = reflect101(b - kEdge, dn);
                    float f = (float)((d + 0.5) * sc - 0.5);
                    int si = floor_int(f);
                    f -= si;
                    if (clampCoef) {                     /* x direction: coefficient is zeroed at the edges */
                        if (si < 0) { si = 0; f = 0.f; }
                        if (si >= sn - 1) { si = sn - 1; f = 0.f; }
                    }
                    ResizeTap r;
                    r.s0 = (int16_t)std::min(std::max(si, 0), sn - 1);
                    r.s1 = (int16_t)std::min(std::max(si + 1, 0), sn - 1);
                    r.c0 = (int16_t)std::min(std::max(round_half_even((1.f - f) * 2048.f), -32768), 32767);
                    r.c1 = (int16_t)std::min(std::max(round_half_even(f * 2048.f), -32768), 32767);
                    t[b] = r;
                }
                size_t o = align_up(tables.size(), 16);
                tables.resize(o + t.size() * sizeof(ResizeTap));
                std::memcpy(tables.data() + o, t.data(), t.size() * sizeof(ResizeTap));
                return (int)o;
            };
            L.xtab = build(L.w, P.w, true);
            L.ytab = build(L.h, P.h, false);
            /* worst-case source rectangle over all 128 x 32 destination tiles (k_resize stages it in shared memory) */
            auto span = [&](int tabOff, int n, int tile, int lead) {
                const ResizeTap* t = reinterpret_cast<const ResizeTap*>(tables.data() + tabOff);
                int worst = 0;
                for (int b0 = -lead; b0 < n; b0 += tile) {
                    int lo = 1 << 30, hi = -1;
                    for (int b = b0; b < b0 + tile; ++b) {
                        const ResizeTap& r = t[std::min(std::max(b, 0), n - 1)];
                        lo = std::min(lo, (int)std::min(r.s0, r.s1)); hi = std::max(hi, (int)std::max(r.s0, r.s1));
                    }
                    worst = std::max(worst, hi - (lo & ~15) + 1);
                }
                return worst;
            };
            /* tiles start at padded column multiples of 128, i.e. bordered column -(kLeftPad-kEdge) + 128k */
            L.rsPitch = (int)align_up((size_t)span(L.xtab, L.w + 2 * kEdge, 128, kLeftPad - kEdge) + 16, 16);
            L.rsRows = span(L.ytab, L.h + 2 * kEdge, 32, 0) + 16;
            /* per tile column / row: first source column (16-byte aligned) / first source row and row count — the
             * origin of the TMA box k_resize stages */
            auto origins = [&](int tabOff, int n, int tile, int lead, int align, bool withCount) {
                const ResizeTap* t = reinterpret_cast<const ResizeTap*>(tables.data() + tabOff);
                std::vector<int16_t> o;
                for (int b0 = -lead; b0 < n; b0 += tile) {
                    int lo = 1 << 30, hi = -1;
                    for (int b = b0; b < b0 + tile; ++b) {
                        const ResizeTap& r = t[std::min(std::max(b, 0), n - 1)];
                        lo = std::min(lo, (int)std::min(r.s0, r.s1)); hi = std::max(hi, (int)std::max(r.s0, r.s1));
                    }
                    lo &= ~(align - 1);
                    o.push_back((int16_t)lo);
                    if (withCount) o.push_back((int16_t)(hi - lo + 1));
                }
                const size_t at = align_up(tables.size(), 16);
                tables.resize(at + o.size() * sizeof(int16_t));
                std::memcpy(tables.data() + at, o.data(), o.size() * sizeof(int16_t));
                return (int)at;
            };
            L.xTile = origins(L.xtab, L.w + 2 * kEdge, 128, kLeftPad - kEdge, 16, false);
            L.yTile = origins(L.ytab, L.h + 2 * kEdge, 32, 0, 1, true);
        }
    }
    g.frameBytes = (long long)align_up(off, 256);
    g.cellsPerFrame = cells;
    g.candPerFrame = cand;
    g.kpPerFrame = kp;
    g.maxNodeCap = maxCap;
    return SDYN_OK;
}

}  // namespace sdyn
