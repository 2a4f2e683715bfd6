/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * CPU restatement of the reference ORB extractor.  Every function cites the reference lines it
 * follows (paths relative to /root/reference).  Parity is pinned by: cv2 4.13.0 known-answer tests
 * for the primitives, the cv2.ORB cross-checks of SURVEY A-8, a Python/cv2 twin of the whole
 * extractor (tests/test_oracle_extractor.py) and the hashes in tests/golden/.  The reference
 * itself ships no tests or golden vectors, so those are the only pins that exist. */
#include "orc_extractor.h"
#include "orc_prims.h"
#include <algorithm>
#include <cmath>
#include <list>
#include <utility>

namespace orc {

static const int kEdge = 19;       /* EDGE_THRESHOLD  ORBextractor.cc:74 */
static const int kHalfPatch = 15;  /* HALF_PATCH_SIZE :73 */
static const int kPatch = 31;      /* PATCH_SIZE :72 */

static const signed char kPairs[256][4] = {
#include "../include/sdyn_brief_pattern.inc"
};

/* ORBextractor.cc:410-470 */
Extractor::Extractor(int nf, float sf, int nl, int ini, int mn)
    : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf)
{
    scale.assign(nl, 1.f); sigma2.assign(nl, 1.f); invScale.resize(nl); invSigma2.resize(nl);
    for (int i = 1; i < nl; ++i) {
        scale[i] = (float)(scale[i - 1] * scaleFactor);   /* float * double member -> float */
        sigma2[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < nl; ++i) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }
    pyr.resize(nl);

    quota.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float want = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nl - 1; ++l) {
        quota[l] = cv_round(want);
        sum += quota[l];
        want *= factor;
    }
    quota[nl - 1] = std::max(nfeatures - sum, 0);

    /* :454-469 disc half-widths */
    umax.assign(kHalfPatch + 1, 0);
    int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

/* ORBextractor.cc:1107-1132 */
void Extractor::computePyramid(const uint8_t* img, int w, int h, int stride)
{
    for (int l = 0; l < nlevels; ++l) {
        float s = invScale[l];
        Level& L = pyr[l];
        L.w = cv_round((float)w * s);
        L.h = cv_round((float)h * s);
        L.stride = L.w + 2 * kEdge;
        L.buf.assign((size_t)L.stride * (L.h + 2 * kEdge), 0);
        if (l == 0) {
            border_reflect101(img, w, h, stride, L.buf.data(), kEdge, L.stride, false);
        } else {
            const Level& P = pyr[l - 1];
            resize_linear_u8(P.roi(), P.w, P.h, P.stride, L.roi(), L.w, L.h, L.stride);
            border_reflect101(nullptr, L.w, L.h, 0, L.buf.data(), kEdge, L.stride, true);
        }
    }
}

namespace {
struct Node {
    std::vector<int> keys;             /* indices into the level's candidate array, stable order */
    int ulx = 0, uly = 0, urx = 0, bry = 0;   /* UL.x, UL.y, UR.x, BR.y — the only corners DivideNode reads */
    std::list<Node>::iterator self;
    bool leaf = false;                 /* bNoMore */
    long seq = 0;                      /* creation order: stands in for the heap address (B-1) */
};

/* ORBextractor.cc:481-537 */
void divide(const Node& p, const std::vector<KeyPoint>& K, Node c[4])
{
    const int hx = (int)std::ceil(static_cast<float>(p.urx - p.ulx) / 2);
    const int hy = (int)std::ceil(static_cast<float>(p.bry - p.uly) / 2);
    const int mx = p.ulx + hx, my = p.uly + hy;
    c[0].ulx = p.ulx; c[0].urx = mx;    c[0].uly = p.uly; c[0].bry = my;
    c[1].ulx = mx;    c[1].urx = p.urx; c[1].uly = p.uly; c[1].bry = my;
    c[2].ulx = p.ulx; c[2].urx = mx;    c[2].uly = my;    c[2].bry = p.bry;
    c[3].ulx = mx;    c[3].urx = p.urx; c[3].uly = my;    c[3].bry = p.bry;
    for (int id : p.keys) {
        const KeyPoint& k = K[id];
        if (k.x < mx) c[k.y < my ? 0 : 2].keys.push_back(id);
        else          c[k.y < my ? 1 : 3].keys.push_back(id);
    }
    for (int i = 0; i < 4; ++i) if (c[i].keys.size() == 1) c[i].leaf = true;
}
}  // namespace

/* ORBextractor.cc:539-763 */
std::vector<KeyPoint> Extractor::distributeOctTree(const std::vector<KeyPoint>& K, int minX, int maxX,
                                                   int minY, int maxY, int N, bool& ok)
{
    ok = true;
    std::vector<KeyPoint> out;
    if (maxY - minY <= 0) { ok = false; return out; }
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    if (nIni <= 0) { ok = false; return out; }   /* reference divides by zero / indexes an empty vector */
    const float hX = static_cast<float>(maxX - minX) / nIni;

    std::list<Node> L;
    long seq = 0;
    std::vector<Node*> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        Node n;
        n.ulx = (int)(hX * static_cast<float>(i));
        n.urx = (int)(hX * static_cast<float>(i + 1));
        n.uly = 0;
        n.bry = maxY - minY;
        n.seq = seq++;
        L.push_back(n);
        roots[i] = &L.back();
    }
    for (size_t i = 0; i < K.size(); ++i) {
        size_t r = (size_t)(K[i].x / hX);
        if (r >= roots.size()) { ok = false; return out; }
        roots[r]->keys.push_back((int)i);
    }
    for (auto it = L.begin(); it != L.end();) {
        if (it->keys.size() == 1) { it->leaf = true; ++it; }
        else if (it->keys.empty()) it = L.erase(it);
        else ++it;
    }

    bool done = false;
    std::vector<std::pair<int, Node*>> expandable;
    auto push_children = [&](Node c[4], int* nExpand) {
        for (int i = 0; i < 4; ++i) {
            if (c[i].keys.empty()) continue;
            c[i].seq = seq++;
            L.push_front(c[i]);
            if (c[i].keys.size() > 1) {
                if (nExpand) ++*nExpand;
                expandable.push_back(std::make_pair((int)c[i].keys.size(), &L.front()));
                L.front().self = L.begin();
            }
        }
    };

    while (!done) {
        int prevSize = (int)L.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = L.begin(); it != L.end();) {
            if (it->leaf) { ++it; continue; }
            Node c[4];
            divide(*it, K, c);
            push_children(c, &nToExpand);
            it = L.erase(it);
        }
        if ((int)L.size() >= N || (int)L.size() == prevSize) {
            done = true;
        } else if ((int)L.size() + nToExpand * 3 > N) {
            while (!done) {
                prevSize = (int)L.size();
                std::vector<std::pair<int, Node*>> prev = expandable;
                expandable.clear();
                /* reference: std::sort on (size, pointer); pointer order pinned to creation order */
                std::sort(prev.begin(), prev.end(), [](const std::pair<int, Node*>& a, const std::pair<int, Node*>& b) {
                    if (a.first != b.first) return a.first < b.first;
                    return a.second->seq < b.second->seq;
                });
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    Node c[4];
                    divide(*prev[j].second, K, c);
                    push_children(c, nullptr);
                    L.erase(prev[j].second->self);
                    if ((int)L.size() >= N) break;
                }
                if ((int)L.size() >= N || (int)L.size() == prevSize) done = true;
            }
        }
    }

    out.reserve(L.size());
    for (const Node& n : L) {
        int best = n.keys[0];
        float r = K[best].response;
        for (size_t k = 1; k < n.keys.size(); ++k)
            if (K[n.keys[k]].response > r) { best = n.keys[k]; r = K[best].response; }
        out.push_back(K[best]);
    }
    return out;
}

/* ORBextractor.cc:77-104 */
float ic_angle(const uint8_t* c, int step, const std::vector<int>& um)
{
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = um[v];
        for (int u = -d; u <= d; ++u) {
            int p = c[u + v * step], q = c[u - v * step];
            vsum += p - q;
            m10 += u * (p + q);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

/* ORBextractor.cc:107-147.  Separate float mul/add (no FMA contraction, Appendix B-3). */
void orb_descriptor(float angle_deg, const uint8_t* c, int step, uint8_t* out)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ang = angle_deg * factorPI;
    const float a = std::cos(ang), b = std::sin(ang);   /* float overloads = cosf/sinf */
    auto sample = [&](int px, int py) -> int {
        const float fy = px * b + py * a;
        const float fx = px * a - py * b;
        return c[cv_round(fy) * step + cv_round(fx)];
    };
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const signed char* q = kPairs[8 * i + k];
            val |= (sample(q[0], q[1]) < sample(q[2], q[3])) << k;
        }
        out[i] = (uint8_t)val;
    }
}

/* ORBextractor.cc:765-853 */
bool Extractor::computeKeyPoints(std::vector<std::vector<KeyPoint>>& all)
{
    all.assign(nlevels, {});
    cand.assign(nlevels, {});
    const float W = 30;
    std::vector<int> cell;
    for (int l = 0; l < nlevels; ++l) {
        const Level& P = pyr[l];
        const int minBX = kEdge - 3, minBY = minBX;
        const int maxBX = P.w - kEdge + 3, maxBY = P.h - kEdge + 3;
        std::vector<KeyPoint> keys;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols <= 0 || nRows <= 0) return false;   /* reference: ceil(x/0) -> undefined */
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                const uint8_t* view = P.roi() + (size_t)y0 * P.stride + x0;
                const int cw = x1 - x0, ch = y1 - y0;
                cell.resize((size_t)3 * std::max(cw * ch, 1));
                int n = fast_nms(view, cw, ch, P.stride, iniTh, cell.data(), cw * ch);
                if (n == 0) n = fast_nms(view, cw, ch, P.stride, minTh, cell.data(), cw * ch);
                for (int k = 0; k < n; ++k) {
                    KeyPoint kp;
                    kp.x = (float)cell[3 * k] + j * wCell;
                    kp.y = (float)cell[3 * k + 1] + i * hCell;
                    kp.size = 7.f; kp.angle = -1.f; kp.response = (float)cell[3 * k + 2];
                    kp.octave = 0; kp.class_id = -1;
                    keys.push_back(kp);
                    cand[l].push_back((int)kp.x); cand[l].push_back((int)kp.y); cand[l].push_back(cell[3 * k + 2]);
                }
            }
        }
        bool ok = true;
        all[l] = distributeOctTree(keys, minBX, maxBX, minBY, maxBY, quota[l], ok);
        if (!ok) return false;
        const int scaledPatch = (int)(kPatch * scale[l]);
        for (KeyPoint& kp : all[l]) {
            kp.x += minBX; kp.y += minBY; kp.octave = l; kp.size = (float)scaledPatch;
        }
    }
    /* :472-479 orientation on the unblurred levels */
    for (int l = 0; l < nlevels; ++l) {
        const Level& P = pyr[l];
        for (KeyPoint& kp : all[l])
            kp.angle = ic_angle(P.roi() + (size_t)cv_round(kp.y) * P.stride + cv_round(kp.x), P.stride, umax);
    }
    return true;
}

/* ORBextractor.cc:1043-1105 */
int Extractor::run(const uint8_t* img, int w, int h, int stride,
                   std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc)
{
    kps.clear(); desc.clear(); perLevel.assign(nlevels, 0);
    if (!img || w <= 0 || h <= 0) return 0;      /* empty image: silent return */
    computePyramid(img, w, h, stride);
    std::vector<std::vector<KeyPoint>> all;
    if (!computeKeyPoints(all)) return -1;
    size_t total = 0;
    for (auto& v : all) total += v.size();
    desc.assign(total * 32, 0);
    kps.reserve(total);
    size_t off = 0;
    std::vector<uint8_t> blur;
    for (int l = 0; l < nlevels; ++l) {
        std::vector<KeyPoint>& v = all[l];
        perLevel[l] = (int)v.size();
        if (v.empty()) continue;
        const Level& P = pyr[l];
        blur.assign((size_t)P.w * P.h, 0);
        gaussian_blur7_s2(P.roi(), P.w, P.h, P.stride, blur.data(), P.w);   /* clone + blur, :1085-1086 */
        for (size_t i = 0; i < v.size(); ++i)
            orb_descriptor(v[i].angle, blur.data() + (size_t)cv_round(v[i].y) * P.w + cv_round(v[i].x), P.w,
                           desc.data() + (off + i) * 32);
        off += v.size();
        if (l != 0) {
            const float s = scale[l];
            for (KeyPoint& kp : v) { kp.x *= s; kp.y *= s; }
        }
        kps.insert(kps.end(), v.begin(), v.end());
    }
    return (int)total;
}

}  // namespace orc
