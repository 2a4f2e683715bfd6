/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * Restatement of the fork's dynamic-keypoint rejection (reference lines cited per function).  Debug side
 * effects (cout, drawMatches, imwrite, putText) are not reproduced (SURVEY B-9).  Undefined behaviour of the
 * reference is pinned as documented inline. */
#include "orc_dynamic.h"
#include "orc_matcher.h"
#include <algorithm>
#include <cmath>
#include <unordered_set>

namespace orc {

static inline bool rect_empty(const Rect& r) { return r.w <= 0 || r.h <= 0; }
static inline bool rect_contains(const Rect& r, double px, double py)
{ return r.x <= px && px < r.x + r.w && r.y <= py && py < r.y + r.h; }
static Rect rect_and(Rect a, const Rect& b)
{
    const double x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    a.w = std::min(a.x + a.w, b.x + b.w) - x1;
    a.h = std::min(a.y + a.h, b.y + b.h) - y1;
    a.x = x1; a.y = y1;
    if (a.w <= 0 || a.h <= 0) a = Rect{0, 0, 0, 0};
    return a;
}
static Rect rect_or(Rect a, const Rect& b)
{
    if (rect_empty(a)) return b;
    if (!rect_empty(b)) {
        const double x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y);
        a.w = std::max(a.x + a.w, b.x + b.w) - x1;
        a.h = std::max(a.y + a.h, b.y + b.h) - y1;
        a.x = x1; a.y = y1;
    }
    return a;
}

/* Frame.cc:481-552 */
void box_track(std::vector<Rect>& boxes, const BoxState& last, int imgW, int imgH, BoxState& cur)
{
    const int n_box = (int)boxes.size();
    cur.box_idx.assign(n_box, -1);
    cur.omit.assign(n_box, 0);
    cur.vel.assign(2 * (size_t)n_box, 0.0);
    if (!last.objects.empty()) {
        for (size_t i = 0; i < last.objects.size(); ++i) {
            double minCost = 1; int minPos = -1;
            for (int j = 0; j < n_box; ++j) {
                const Rect inter = rect_and(last.objects[i], boxes[j]);
                const Rect uni = rect_or(last.objects[i], boxes[j]);
                const double cost = 1 - (inter.w * inter.h) / (uni.w * uni.h);
                if (cost < minCost) { minCost = cost; minPos = j; }
            }
            if (minPos != -1 && !last.omit[i]) {
                cur.box_idx[minPos] = last.box_idx[i];
                cur.vel[2 * minPos] = boxes[minPos].x + boxes[minPos].w / 2 - last.objects[i].x - last.objects[i].w / 2;
                cur.vel[2 * minPos + 1] = boxes[minPos].y + boxes[minPos].h / 2 - last.objects[i].y - last.objects[i].h / 2;
            }
        }
        for (size_t i = 0; i < last.objects.size(); ++i) {
            if (last.omit[i]) continue;
            if (std::count(cur.box_idx.begin(), cur.box_idx.end(), last.box_idx[i])) continue;
            const float ccx = (float)(last.objects[i].x + last.objects[i].w / 2 + last.vel[2 * i]);
            const float ccy = (float)(last.objects[i].y + last.objects[i].h / 2 + last.vel[2 * i + 1]);
            /* Rect2f(Point2f(0,0), Point2f(cols, rows)).contains(center_cur) */
            if (0.f <= ccx && ccx < (float)imgW && 0.f <= ccy && ccy < (float)imgH) {
                Rect b = last.objects[i];
                b.x += last.vel[2 * i]; b.y += last.vel[2 * i + 1];
                boxes.push_back(b);
                cur.box_idx.push_back(last.box_idx[i]);
                cur.omit.push_back(1);
                cur.vel.push_back(last.vel[2 * i]); cur.vel.push_back(last.vel[2 * i + 1]);
            }
        }
        for (int i = 0; i < n_box; ++i) {
            if (cur.box_idx[i] != -1) continue;
            const int mx = *std::max_element(cur.box_idx.begin(), cur.box_idx.end());
            cur.box_idx[i] = mx + 1;
        }
    } else {
        for (int i = 0; i < n_box; ++i) cur.box_idx[i] = i;
    }
}

/* Frame.cc:555-604 + :337-367 */
SplitResult first_separate(const KeyPoint* keys, int N, std::vector<Rect>& boxes, BoxState& cur)
{
    SplitResult R;
    R.hasKpts.assign(boxes.size(), 0);
    R.class_id.resize(N);
    std::vector<int> stat, dyn;
    for (int i = 0; i < N; ++i) {
        R.class_id[i] = keys[i].class_id;
        std::vector<int> idx;
        for (size_t j = 0; j < boxes.size(); ++j) {
            if (rect_contains(boxes[j], (double)keys[i].x, (double)keys[i].y)) {
                R.hasKpts[j] = 1;
                if (R.class_id[i] == -1) R.class_id[i] = i;
                idx.push_back((int)j);
            }
        }
        if (!idx.empty()) { R.index.push_back(idx); dyn.push_back(i); }
        else stat.push_back(i);
    }
    /* erase boxes without keypoints — `i` keeps advancing after an erase and hasKpts is NOT erased in step
     * (reference behaviour, Appendix B-4) */
    for (size_t i = 0; i < boxes.size(); ++i) {
        if (R.hasKpts[i]) continue;
        boxes.erase(boxes.begin() + i);
        cur.box_idx.erase(cur.box_idx.begin() + i);
        cur.omit.erase(cur.omit.begin() + i);
        cur.vel.erase(cur.vel.begin() + 2 * i, cur.vel.begin() + 2 * i + 2);
        R.empty_box = true;
    }
    R.N_d = (int)R.index.size();
    R.order = stat;
    R.order.insert(R.order.end(), dyn.begin(), dyn.end());
    if (R.N_d > 0) {
        R.dynKeys.resize(boxes.size());
        for (int i = 0; i < R.N_d; ++i) {
            for (size_t j = 0; j < R.index[i].size(); ++j) {
                if (R.empty_box)
                    R.index[i][j] -= (int)std::count(R.hasKpts.begin(), R.hasKpts.begin() + R.index[i][j] + 1, 0);
                R.dynKeys[R.index[i][j]].push_back(dyn[i]);
            }
        }
    }
    return R;
}

/* SURVEY A-7 */
std::vector<Match> bf_match_crosscheck(const uint8_t* q, int nq, const uint8_t* t, int nt)
{
    std::vector<Match> out;
    if (nq == 0 || nt == 0) return out;
    std::vector<int> nnQ(nq), dQ(nq), nnT(nt, -1), dT(nt, 1 << 30);
    for (int i = 0; i < nq; ++i) {
        int best = 1 << 30, bj = -1;
        for (int j = 0; j < nt; ++j) {
            const int d = descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < best) { best = d; bj = j; }
            if (d < dT[j]) { dT[j] = d; nnT[j] = i; }
        }
        nnQ[i] = bj; dQ[i] = best;
    }
    for (int i = 0; i < nq; ++i)
        if (nnT[nnQ[i]] == i) out.push_back({i, nnQ[i], dQ[i]});
    return out;
}

/* cv::invert (DECOMP_LU) special case for 3x3 CV_32F: closed form in double, pinned against cv2.invert */
void invert3x3(const float* m, float* o)
{
    const double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0) { for (int k = 0; k < 9; ++k) o[k] = 0; return; }
    det = 1. / det;
    o[0] = (float)((e * i - f * h) * det); o[1] = (float)((c * h - b * i) * det); o[2] = (float)((b * f - c * e) * det);
    o[3] = (float)((f * g - d * i) * det); o[4] = (float)((a * i - c * g) * det); o[5] = (float)((c * d - a * f) * det);
    o[6] = (float)((d * h - e * g) * det); o[7] = (float)((b * g - a * h) * det); o[8] = (float)((a * e - b * d) * det);
}

/* Tracking.cc:1311-1367 */
void classify_f(const float* F, const float* cur, const float* ref, const std::vector<Match>& m, std::vector<int>& out)
{
    const float f11 = F[0], f12 = F[1], f13 = F[2], f21 = F[3], f22 = F[4], f23 = F[5], f31 = F[6], f32 = F[7], f33 = F[8];
    const float th = 5.841;
    const float invSigmaSquare = 1.0 / (1.0f * 1.0f);
    for (size_t i = 0; i < m.size(); ++i) {
        const float u1 = ref[2 * m[i].trainIdx], v1 = ref[2 * m[i].trainIdx + 1];
        const float u2 = cur[2 * m[i].queryIdx], v2 = cur[2 * m[i].queryIdx + 1];
        const float a2 = f11 * u1 + f12 * v1 + f13;
        const float b2 = f21 * u1 + f22 * v1 + f23;
        const float c2 = f31 * u1 + f32 * v1 + f33;
        const float num2 = a2 * u2 + b2 * v2 + c2;
        const float squareDist1 = num2 * num2 / (a2 * a2 + b2 * b2);
        const float chiSquare1 = squareDist1 * invSigmaSquare;
        const float a1 = f11 * u2 + f21 * v2 + f31;
        const float b1 = f12 * u2 + f22 * v2 + f32;
        const float c1 = f13 * u2 + f23 * v2 + f33;
        const float num1 = a1 * u1 + b1 * v1 + c1;
        const float squareDist2 = num1 * num1 / (a1 * a1 + b1 * b1);
        const float chiSquare2 = squareDist2 * invSigmaSquare;
        if (chiSquare1 <= th && chiSquare2 <= th) out[i] = m[i].queryIdx;
    }
}

/* Tracking.cc:1241-1309 */
void classify_h(const float* H, const float* cur, const float* ref, const std::vector<Match>& m, std::vector<int>& out)
{
    float Hi[9];
    invert3x3(H, Hi);
    const float th = 5.991;
    const float invSigmaSquare = 1.0 / (1.0f * 1.0f);
    for (size_t i = 0; i < m.size(); ++i) {
        const float u1 = ref[2 * m[i].trainIdx], v1 = ref[2 * m[i].trainIdx + 1];
        const float u2 = cur[2 * m[i].queryIdx], v2 = cur[2 * m[i].queryIdx + 1];
        const float w2in1inv = (float)(1.0 / (Hi[6] * u2 + Hi[7] * v2 + Hi[8]));
        const float u2in1 = (Hi[0] * u2 + Hi[1] * v2 + Hi[2]) * w2in1inv;
        const float v2in1 = (Hi[3] * u2 + Hi[4] * v2 + Hi[5]) * w2in1inv;
        const float squareDist1 = (u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1);
        const float chiSquare1 = squareDist1 * invSigmaSquare;
        const float w1in2inv = (float)(1.0 / (H[6] * u1 + H[7] * v1 + H[8]));
        const float u1in2 = (H[0] * u1 + H[1] * v1 + H[2]) * w1in2inv;
        const float v1in2 = (H[3] * u1 + H[4] * v1 + H[5]) * w1in2inv;
        const float squareDist2 = (u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2);
        const float chiSquare2 = squareDist2 * invSigmaSquare;
        if (chiSquare2 <= th && chiSquare1 <= th) out[i] = m[i].queryIdx;
    }
}

/* Tracking.cc:1093-1239 */
int separate(const std::vector<BoxKeys>& cur, const std::vector<int>& curBoxIdx, std::vector<int>& curBoxStatus,
             const std::vector<BoxKeys>& ref, const std::vector<int>& refBoxIdx,
             const std::vector<int>& lastBoxIdx, const std::vector<int>& lastBoxStatus,
             const float* HorF, int flag, std::vector<std::vector<int>>& dynStatus,
             std::vector<std::vector<Match>>& matches)
{
    const int nb = (int)curBoxIdx.size();
    dynStatus.assign(nb, {});
    matches.assign(nb, {});
    bool static_exit = false;
    for (int b = 0; b < nb; ++b) {
        auto it = std::find(refBoxIdx.begin(), refBoxIdx.end(), curBoxIdx[b]);
        if (it == refBoxIdx.end()) continue;
        const int r = (int)(it - refBoxIdx.begin());
        /* the reference indexes mdynDescriptors[...] even when the vector was never resized (N_d == 0):
         * pinned to "no descriptors" */
        if (b >= (int)cur.size() || r >= (int)ref.size()) continue;
        const int nq = (int)cur[b].desc.size() / 32, nt = (int)ref[r].desc.size() / 32;
        if (nq == 0 || nt == 0) continue;
        std::vector<Match> good = bf_match_crosscheck(cur[b].desc.data(), nq, ref[r].desc.data(), nt);
        matches[b] = good;
        if (good.size() < 3 || good.size() < 0.2 * nq) continue;
        std::vector<int>& st = dynStatus[b];
        st.assign(good.size(), -1);
        if (flag == 1) classify_h(HorF, cur[b].xy.data(), ref[r].xy.data(), good, st);
        else classify_f(HorF, cur[b].xy.data(), ref[r].xy.data(), good, st);
        const int num0 = (int)st.size() - (int)std::count(st.begin(), st.end(), -1);
        if (num0 > std::max((double)1, 0.2 * good.size())) {
            static_exit = true;      /* `box_status[n_box] == 1;` is a comparison in the reference (B-5) */
        } else {
            auto lt = std::find(lastBoxIdx.begin(), lastBoxIdx.end(), curBoxIdx[b]);
            /* not found: the reference reads one past the end; pinned to "status -1" */
            const int ls = lt == lastBoxIdx.end() ? -1 : lastBoxStatus[lt - lastBoxIdx.begin()];
            curBoxStatus[b] = (ls == 0 || ls == 2) ? 2 : 0;
        }
    }
    return static_exit ? 1 : 0;
}

/* Frame.cc:607-641 */
std::vector<std::pair<int, int>> update_frame(const std::vector<std::vector<int>>& dynStatus,
                                              const std::vector<std::vector<int>>& dynClassId)
{
    std::vector<std::pair<int, int>> pushed;
    std::unordered_set<int> seen;
    for (size_t i = 0; i < dynStatus.size(); ++i) {
        if (dynStatus[i].empty()) continue;
        for (size_t j = 0; j < dynStatus[i].size(); ++j) {
            const int k = dynStatus[i][j];
            if (k == -1) continue;
            const int cid = dynClassId[i][k];
            if (seen.count(cid)) continue;
            seen.insert(cid);
            pushed.push_back({(int)i, k});
        }
    }
    return pushed;
}

}  // namespace orc
