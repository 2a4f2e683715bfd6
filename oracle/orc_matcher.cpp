/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * Restatement of the hot-path ORBmatcher searches and Frame grid (reference lines cited per function).
 * The reference has no tests or golden vectors for these: parity is pinned by the literal,
 * line-by-line structure below plus brute-force cross-checks in tests/test_oracle_matcher.py.
 * Float arithmetic is evaluated without FMA contraction (build flag -ffp-contract=off, Appendix B-3). */
#include "orc_matcher.h"
#include <algorithm>
#include <climits>
#include <cmath>

namespace orc {

/* ORBmatcher.cc:1804-1820 (SWAR popcount over 8 x 32 bit) */
int descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(a);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(b);
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

/* Frame.cc:383-385 */
void grid_bounds(FrameView& f)
{
    f.gridWInv = static_cast<float>(GRID_COLS) / static_cast<float>(f.maxX - f.minX);
    f.gridHInv = static_cast<float>(GRID_ROWS) / static_cast<float>(f.maxY - f.minY);
}

/* Frame.cc:790-800 */
static bool pos_in_grid(const FrameView& f, const KeyPoint& kp, int& px, int& py)
{
    px = (int)std::round((kp.x - f.minX) * f.gridWInv);
    py = (int)std::round((kp.y - f.minY) * f.gridHInv);
    return !(px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS);
}

/* Frame.cc:463-478 (from = 0) and UpdateFeaturesToGrid :643-653 (from = N_ori) */
void assign_features_to_grid(const FrameView& f, Grid& g, int from)
{
    for (int i = from; i < f.N; ++i) {
        int px, py;
        if (pos_in_grid(f, f.keysUn[i], px, py)) g.cell[px][py].push_back(i);
    }
}

/* Frame.cc:735-788 */
std::vector<int> features_in_area(const FrameView& f, const Grid& g, float x, float y, float r,
                                  int minLevel, int maxLevel)
{
    std::vector<int> out;
    const int nMinCellX = std::max(0, (int)std::floor((x - f.minX - r) * f.gridWInv));
    if (nMinCellX >= GRID_COLS) return out;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - f.minX + r) * f.gridWInv));
    if (nMaxCellX < 0) return out;
    const int nMinCellY = std::max(0, (int)std::floor((y - f.minY - r) * f.gridHInv));
    if (nMinCellY >= GRID_ROWS) return out;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - f.minY + r) * f.gridHInv));
    if (nMaxCellY < 0) return out;
    const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
        for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
            for (int id : g.cell[ix][iy]) {
                const KeyPoint& kp = f.keysUn[id];
                if (checkLevels) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(id);
            }
    return out;
}

/* ORBmatcher.cc:1758-1799 */
void compute_three_maxima(const int* h, int L, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = h[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

static inline int rot_bin(float rot)
{
    const float factor = 1.0f / HISTO_LENGTH;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

/* ORBmatcher.cc:45-129 */
int search_by_projection_map(const FrameView& F, const Grid& g, const MapPointQuery* mps, int nmp, float th,
                             float nnratio, int32_t* assign, uint8_t* locked, int assignBase)
{
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int iMP = 0; iMP < nmp; ++iMP) {
        const MapPointQuery& mp = mps[iMP];
        if (!mp.trackInView) continue;
        if (mp.bad) continue;
        const int level = mp.level;
        float r = ((double)mp.viewCos > 0.998) ? 2.5f : 4.0f;      /* RadiusByViewingCos :131-137 */
        if (bFactor) r *= th;
        const float win = r * F.scaleFactors[level];
        const std::vector<int> cand = features_in_area(F, g, mp.projX, mp.projY, win, level - 1, level);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (assign[idx] != -1 && locked[idx]) continue;       /* occupant with Observations() > 0 */
            if (F.uRight && F.uRight[idx] > 0) {
                const float er = std::fabs(mp.projXR - F.uRight[idx]);
                if (er > win) continue;
            }
            const int dist = descriptor_distance(mp.desc, F.desc + 32 * (size_t)idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F.keysUn[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F.keysUn[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            assign[bestIdx] = assignBase + iMP;   /* index encoding of `F.mvpMapPoints[bestIdx]=pMP` */
            locked[bestIdx] = mp.obsPositive;
            ++nmatches;
        }
    }
    return nmatches;
}

/* y = R*x + t the way cv::Mat evaluates `R*x+t` for 3x3 * 3x1 CV_32F: float products summed left to
 * right, then the addend (pinned against cv2.gemm in tests/test_oracle_matcher.py). */
static inline void rx_plus_t(const float* T, const float* x, float* y)
{
    for (int r = 0; r < 3; ++r) {
        const float s = T[4 * r] * x[0] + T[4 * r + 1] * x[1] + T[4 * r + 2] * x[2];
        y[r] = s + T[4 * r + 3];
    }
}

/* tlc of ORBmatcher.cc:1497-1503: twc = -Rcw^T tcw (generic gemm path: double accumulation, alpha = -1),
 * tlc = Rlw twc + tlw (3x3 small-matrix path). */
void motion_tlc(const float* Tcw, const float* Tlw, float* tlc)
{
    float twc[3];
    for (int r = 0; r < 3; ++r) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (double)Tcw[4 * k + r] * (double)Tcw[4 * k + 3];
        twc[r] = (float)(-1.0 * s);
    }
    rx_plus_t(Tlw, twc, tlc);
}

/* ORBmatcher.cc:1485-1627 / :407-559 */
int search_by_projection_frame(const FrameView& Cur, const Grid& gCur, const FrameView& Last,
                               const LastFramePoint* lp, float th, bool mono, bool checkOri,
                               int32_t* assign, uint8_t* locked, std::vector<float>* pairs)
{
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    float tlc[3];
    motion_tlc(Cur.Tcw, Last.Tcw, tlc);
    const bool bForward = tlc[2] > Cur.b && !mono;
    const bool bBackward = -tlc[2] > Cur.b && !mono;

    for (int i = 0; i < Last.N; ++i) {
        if (!lp[i].hasMP) continue;
        if (lp[i].outlier) continue;
        float pc[3];
        rx_plus_t(Cur.Tcw, lp[i].world, pc);
        const float xc = pc[0], yc = pc[1];
        const float invzc = (float)(1.0 / pc[2]);
        if (invzc < 0) continue;
        const float u = Cur.fx * xc * invzc + Cur.cx;
        const float v = Cur.fy * yc * invzc + Cur.cy;
        if (u < Cur.minX || u > Cur.maxX) continue;
        if (v < Cur.minY || v > Cur.maxY) continue;
        const int nLastOctave = Last.keys[i].octave;
        const float radius = th * Cur.scaleFactors[nLastOctave];
        std::vector<int> cand;
        if (bForward) cand = features_in_area(Cur, gCur, u, v, radius, nLastOctave, -1);
        else if (bBackward) cand = features_in_area(Cur, gCur, u, v, radius, 0, nLastOctave);
        else cand = features_in_area(Cur, gCur, u, v, radius, nLastOctave - 1, nLastOctave + 1);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (assign[i2] != -1 && locked[i2]) continue;
            if (Cur.uRight && Cur.uRight[i2] > 0) {
                const float ur = u - Cur.bf * invzc;
                const float er = std::fabs(ur - Cur.uRight[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(lp[i].desc, Cur.desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            assign[bestIdx2] = i;
            locked[bestIdx2] = lp[i].obsPositive;
            ++nmatches;
            if (pairs) {
                pairs->push_back(Last.keysUn[i].x); pairs->push_back(Last.keysUn[i].y);
                pairs->push_back(Cur.keysUn[bestIdx2].x); pairs->push_back(Cur.keysUn[bestIdx2].y);
            }
            if (checkOri) rotHist[rot_bin(Last.keysUn[i].angle - Cur.keysUn[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { assign[idx] = -1; locked[idx] = 0; --nmatches; }
        }
    }
    return nmatches;
}

/* ORBmatcher.cc:562-677 */
int search_for_initialization(const FrameView& F1, const FrameView& F2, const Grid& g2, float* prevMatched,
                              int32_t* m12, int windowSize, float nnratio, bool checkOri)
{
    int nmatches = 0;
    std::fill(m12, m12 + F1.N, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> matchedDistance(F2.N, INT_MAX), m21(F2.N, -1);
    for (int i1 = 0; i1 < F1.N; ++i1) {
        const int level1 = F1.keysUn[i1].octave;
        if (level1 > 0) continue;
        const std::vector<int> cand = features_in_area(F2, g2, prevMatched[2 * i1], prevMatched[2 * i1 + 1],
                                                       (float)windowSize, level1, level1);
        if (cand.empty()) continue;
        const uint8_t* d1 = F1.desc + 32 * (size_t)i1;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int i2 : cand) {
            const int dist = descriptor_distance(d1, F2.desc + 32 * (size_t)i2);
            if (matchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (m21[bestIdx2] >= 0) { m12[m21[bestIdx2]] = -1; --nmatches; }
                m12[i1] = bestIdx2;
                m21[bestIdx2] = i1;
                matchedDistance[bestIdx2] = bestDist;
                ++nmatches;
                if (checkOri) rotHist[rot_bin(F1.keysUn[i1].angle - F2.keysUn[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i])
                if (m12[idx1] >= 0) { m12[idx1] = -1; --nmatches; }
        }
    }
    for (int i1 = 0; i1 < F1.N; ++i1)
        if (m12[i1] >= 0) { prevMatched[2 * i1] = F2.keysUn[m12[i1]].x; prevMatched[2 * i1 + 1] = F2.keysUn[m12[i1]].y; }
    return nmatches;
}

/* ORBmatcher.cc:159-288 */
int search_by_bow(const FrameView& KF, const uint8_t* kfValid, const FeatureVec& a, const FrameView& F,
                  const FeatureVec& b, float nnratio, bool checkOri, int32_t* assign)
{
    std::fill(assign, assign + F.N, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    int ia = 0, ib = 0;
    while (ia < a.nnodes && ib < b.nnodes) {
        if (a.nodeId[ia] == b.nodeId[ib]) {
            for (int p = a.offset[ia]; p < a.offset[ia + 1]; ++p) {
                const unsigned kfIdx = a.index[p];
                if (!kfValid[kfIdx]) continue;
                const uint8_t* dKF = KF.desc + 32 * (size_t)kfIdx;
                int best1 = 256, bestIdxF = -1, best2 = 256;
                for (int q = b.offset[ib]; q < b.offset[ib + 1]; ++q) {
                    const unsigned fIdx = b.index[q];
                    if (assign[fIdx] != -1) continue;
                    const int dist = descriptor_distance(dKF, F.desc + 32 * (size_t)fIdx);
                    if (dist < best1) { best2 = best1; best1 = dist; bestIdxF = (int)fIdx; }
                    else if (dist < best2) best2 = dist;
                }
                if (best1 <= TH_LOW) {
                    if (static_cast<float>(best1) < nnratio * static_cast<float>(best2)) {
                        assign[bestIdxF] = (int)kfIdx;
                        if (checkOri) rotHist[rot_bin(KF.keysUn[kfIdx].angle - F.keys[bestIdxF].angle)].push_back(bestIdxF);
                        ++nmatches;
                    }
                }
            }
            ++ia; ++ib;
        } else if (a.nodeId[ia] < b.nodeId[ib]) {
            ia = (int)(std::lower_bound(a.nodeId, a.nodeId + a.nnodes, b.nodeId[ib]) - a.nodeId);
        } else {
            ib = (int)(std::lower_bound(b.nodeId, b.nodeId + b.nnodes, a.nodeId[ia]) - b.nodeId);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { assign[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) — src/ORBmatcher.cc:679-812.
 * valid1 / valid2 stand for "vpMapPoints[idx] != NULL && !isBad()"; matches12[i1] = index into KeyFrame 2 or -1. */
int search_by_bow_kf(const FrameView& KF1, const uint8_t* valid1, const FeatureVec& a, const FrameView& KF2,
                     const uint8_t* valid2, const FeatureVec& b, float nnratio, bool checkOri, int32_t* matches12)
{
    std::fill(matches12, matches12 + KF1.N, -1);                       /* :691 */
    std::vector<bool> vbMatched2(KF2.N, false);                         /* :692 */
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    int ia = 0, ib = 0;
    while (ia < a.nnodes && ib < b.nnodes) {                            /* :707 */
        if (a.nodeId[ia] == b.nodeId[ib]) {
            for (int p = a.offset[ia]; p < a.offset[ia + 1]; ++p) {
                const unsigned idx1 = a.index[p];
                if (!valid1[idx1]) continue;                            /* :715-719 */
                const uint8_t* d1 = KF1.desc + 32 * (size_t)idx1;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int q = b.offset[ib]; q < b.offset[ib + 1]; ++q) {
                    const unsigned idx2 = b.index[q];
                    if (vbMatched2[idx2] || !valid2[idx2]) continue;     /* :733-737 */
                    const int dist = descriptor_distance(d1, KF2.desc + 32 * (size_t)idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = (int)idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {                               /* :755 — strict, unlike the KeyFrame-Frame overload */
                    if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        matches12[idx1] = bestIdx2;
                        vbMatched2[bestIdx2] = true;
                        if (checkOri) rotHist[rot_bin(KF1.keysUn[idx1].angle - KF2.keysUn[bestIdx2].angle)].push_back((int)idx1);
                        ++nmatches;
                    }
                }
            }
            ++ia; ++ib;
        } else if (a.nodeId[ia] < b.nodeId[ib]) {
            ia = (int)(std::lower_bound(a.nodeId, a.nodeId + a.nnodes, b.nodeId[ib]) - a.nodeId);
        } else {
            ib = (int)(std::lower_bound(b.nodeId, b.nodeId + b.nnodes, a.nodeId[ia]) - b.nodeId);
        }
    }
    if (checkOri) {                                                     /* :790-809 */
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { matches12[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

/* cv::norm(3x1 CV_32F) and Mat::dot of two 3x1 CV_32F: products of the float elements accumulated in double */
static inline double norm3(const float* p) { double s = 0; for (int k = 0; k < 3; ++k) { const double v = p[k]; s += v * v; } return std::sqrt(s); }
static inline double dot3(const float* a, const float* b) { double r = 0; for (int k = 0; k < 3; ++k) r += (double)a[k] * b[k]; return r; }

/* MapPoint::PredictScale, src/MapPoint.cc:385-418 (`log` resolves to the float overload) */
static inline int predict_scale(float mfMaxDistance, float currentDist, float mfLogScaleFactor, int mnScaleLevels)
{
    const float ratio = mfMaxDistance / currentDist;
    int nScale = (int)std::ceil(std::log(ratio) / mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= mnScaleLevels) nScale = mnScaleLevels - 1;
    return nScale;
}

/* Frame::isInFrustum, src/Frame.cc:677-733 */
bool is_in_frustum(const FrameView& F, float mfLogScaleFactor, const float* world, const float* normal, float mfMinDistance,
                   float mfMaxDistance, float viewingCosLimit, float* projX, float* projY, float* projXR, int* level, float* viewCos)
{
    /* Pc = mRcw*P + mtcw: cv::Mat's fused 3x3 product (float products left to right, addend joined in double) */
    float Pc[3];
    for (int r = 0; r < 3; ++r) {
        const float t = F.Tcw[4 * r] * world[0] + F.Tcw[4 * r + 1] * world[1] + F.Tcw[4 * r + 2] * world[2];
        Pc[r] = (float)((double)t * 1.0 + (double)F.Tcw[4 * r + 3] * 1.0);
    }
    const float PcX = Pc[0], PcY = Pc[1], PcZ = Pc[2];
    if (PcZ < 0.0f) return false;
    const float invz = 1.0f / PcZ;
    const float u = F.fx * PcX * invz + F.cx;
    const float v = F.fy * PcY * invz + F.cy;
    if (u < F.minX || u > F.maxX) return false;
    if (v < F.minY || v > F.maxY) return false;
    const float maxDistance = 1.2f * mfMaxDistance;        /* GetMaxDistanceInvariance */
    const float minDistance = 0.8f * mfMinDistance;        /* GetMinDistanceInvariance */
    /* mOw = -mRcw.t()*mtcw (Frame::UpdatePoseMatrices :670-676; transposed operand: double accumulation) */
    float Ow[3];
    for (int r = 0; r < 3; ++r) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (double)F.Tcw[4 * k + r] * (double)F.Tcw[4 * k + 3];
        Ow[r] = (float)(-1.0 * s);
    }
    const float PO[3] = {world[0] - Ow[0], world[1] - Ow[1], world[2] - Ow[2]};
    const float dist = (float)norm3(PO);
    if (dist < minDistance || dist > maxDistance) return false;
    const float vc = (float)(dot3(PO, normal) / dist);
    if (vc < viewingCosLimit) return false;
    *level = predict_scale(mfMaxDistance, dist, mfLogScaleFactor, F.nlevels);
    *projX = u; *projXR = u - F.bf * invz; *projY = v; *viewCos = vc;
    return true;
}

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
 * src/ORBmatcher.cc:1629-1756.  Rcw/tcw/Ow: the three cv::Mat values of :1633-1635 as computed by the caller. */
int search_by_projection_reloc(const FrameView& Cur, const Grid& gCur, const ProjPoint* pts, int npts, const float* Rcw,
                               const float* tcw, const float* Ow, float th, int ORBdist, bool checkOri,
                               float mfLogScaleFactor, int mnScaleLevels, int32_t* assign)
{
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    float T[12];
    for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) T[4 * r + k] = Rcw[3 * r + k]; T[4 * r + 3] = tcw[r]; }
    for (int i = 0; i < npts; ++i) {
        if (!pts[i].valid) continue;                                    /* :1649-1652 */
        float pc[3];
        rx_plus_t(T, pts[i].world, pc);                                 /* x3Dc = Rcw*x3Dw+tcw */
        const float xc = pc[0], yc = pc[1];
        const float invzc = (float)(1.0 / pc[2]);
        const float u = Cur.fx * xc * invzc + Cur.cx;
        const float v = Cur.fy * yc * invzc + Cur.cy;
        if (u < Cur.minX || u > Cur.maxX) continue;
        if (v < Cur.minY || v > Cur.maxY) continue;
        float PO[3];
        for (int k = 0; k < 3; ++k) PO[k] = pts[i].world[k] - Ow[k];
        const float dist3D = (float)norm3(PO);                          /* :1672 */
        if (dist3D < pts[i].minDistance || dist3D > pts[i].maxDistance) continue;
        const int nPredictedLevel = predict_scale(pts[i].maxDistanceRaw, dist3D, mfLogScaleFactor, mnScaleLevels);
        const float radius = th * Cur.scaleFactors[nPredictedLevel];
        const std::vector<int> vIndices2 = features_in_area(Cur, gCur, u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1);
        if (vIndices2.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : vIndices2) {
            if (assign[i2] != -1) continue;                             /* CurrentFrame.mvpMapPoints[i2] */
            const int dist = descriptor_distance(pts[i].desc, Cur.desc + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= ORBdist) {                                      /* :1716 */
            assign[bestIdx2] = i;
            ++nmatches;
            if (checkOri) rotHist[rot_bin(pts[i].angle - Cur.keysUn[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {                                                     /* :1736-1753 */
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { assign[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
 * src/ORBmatcher.cc:290-403.  Rcw/tcw/Ow: the values of :300-304 (sRcw/scw, t/scw, -Rcw.t()*tcw) from the caller.
 * KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:569-608) is Frame's without a level check. */
int search_by_projection_sim3(const FrameView& KF, const Grid& gKF, const ProjPoint* pts, int npts, const float* Rcw,
                              const float* tcw, const float* Ow, int th, float mfLogScaleFactor, int mnScaleLevels, int32_t* assign)
{
    int nmatches = 0;
    float T[12];
    for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) T[4 * r + k] = Rcw[3 * r + k]; T[4 * r + 3] = tcw[r]; }
    for (int iMP = 0; iMP < npts; ++iMP) {
        if (!pts[iMP].valid) continue;                                  /* isBad() || spAlreadyFound.count(pMP), :316 */
        float p3Dc[3];
        rx_plus_t(T, pts[iMP].world, p3Dc);
        if (p3Dc[2] < 0.0) continue;                                    /* :325 */
        const float invz = 1 / p3Dc[2];
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = KF.fx * x + KF.cx, v = KF.fy * y + KF.cy;
        if (!(u >= KF.minX && u < KF.maxX && v >= KF.minY && v < KF.maxY)) continue;    /* KeyFrame::IsInImage */
        float PO[3];
        for (int k = 0; k < 3; ++k) PO[k] = pts[iMP].world[k] - Ow[k];
        const float dist = (float)norm3(PO);
        if (dist < pts[iMP].minDistance || dist > pts[iMP].maxDistance) continue;
        if (dot3(PO, pts[iMP].normal) < 0.5 * dist) continue;           /* :350 */
        const int nPredictedLevel = predict_scale(pts[iMP].maxDistanceRaw, dist, mfLogScaleFactor, mnScaleLevels);
        const float radius = th * KF.scaleFactors[nPredictedLevel];
        const std::vector<int> vIndices = features_in_area(KF, gKF, u, v, radius, -1, -1);
        if (vIndices.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            if (assign[idx] != -1) continue;                            /* vpMatched[idx] */
            const int kpLevel = KF.keysUn[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int d = descriptor_distance(pts[iMP].desc, KF.desc + 32 * (size_t)idx);
            if (d < bestDist) { bestDist = d; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) { assign[bestIdx] = iMP; ++nmatches; }  /* :393 */
    }
    return nmatches;
}

/* Search part of ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th),
 * src/ORBmatcher.cc:982-1100: per MapPoint the keypoint Fuse would pick (bestIdx, bestDist); the Replace /
 * AddObservation bookkeeping (:1103-1126) is pointer-graph code of the caller.  pts[i].valid stands for
 * pMP && !pMP->isBad() && !pMP->IsInKeyFrame(pKF). */
void fuse_search(const FrameView& KF, const Grid& gKF, const float* invLevelSigma2, const ProjPoint* pts, int npts,
                 const float* Rcw, const float* tcw, const float* Ow, float th, float mfLogScaleFactor, int mnScaleLevels,
                 int32_t* bestIdxOut, int32_t* bestDistOut)
{
    float T[12];
    for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) T[4 * r + k] = Rcw[3 * r + k]; T[4 * r + 3] = tcw[r]; }
    const float bf = KF.bf;
    for (int i = 0; i < npts; ++i) {
        bestIdxOut[i] = -1; bestDistOut[i] = 256;
        if (!pts[i].valid) continue;
        float p3Dc[3];
        rx_plus_t(T, pts[i].world, p3Dc);
        if (p3Dc[2] < 0.0f) continue;                                   /* :1015 */
        const float invz = 1 / p3Dc[2];
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = KF.fx * x + KF.cx, v = KF.fy * y + KF.cy;
        if (!(u >= KF.minX && u < KF.maxX && v >= KF.minY && v < KF.maxY)) continue;
        const float ur = u - bf * invz;
        float PO[3];
        for (int k = 0; k < 3; ++k) PO[k] = pts[i].world[k] - Ow[k];
        const float dist3D = (float)norm3(PO);
        if (dist3D < pts[i].minDistance || dist3D > pts[i].maxDistance) continue;
        if (dot3(PO, pts[i].normal) < 0.5 * dist3D) continue;           /* :1042 */
        const int nPredictedLevel = predict_scale(pts[i].maxDistanceRaw, dist3D, mfLogScaleFactor, mnScaleLevels);
        const float radius = th * KF.scaleFactors[nPredictedLevel];
        const std::vector<int> vIndices = features_in_area(KF, gKF, u, v, radius, -1, -1);
        if (vIndices.empty()) continue;
        int bestDist = 256, bestIdx = -1;
        for (int idx : vIndices) {
            const KeyPoint& kp = KF.keysUn[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            if (KF.uRight && KF.uRight[idx] >= 0) {                     /* :1067-1080 */
                const float kpx = kp.x, kpy = kp.y, kpr = KF.uRight[idx];
                const float ex = u - kpx, ey = v - kpy, er = ur - kpr;
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * invLevelSigma2[kpLevel] > 7.8) continue;
            } else {
                const float kpx = kp.x, kpy = kp.y;
                const float ex = u - kpx, ey = v - kpy;
                const float e2 = ex * ex + ey * ey;
                if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
            }
            const int dist = descriptor_distance(pts[i].desc, KF.desc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        bestIdxOut[i] = bestIdx; bestDistOut[i] = bestDist;
    }
}

/* Search part of ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, float th,
 * vector<MapPoint *> &vpReplacePoint), src/ORBmatcher.cc:1132-1237.  pts[i].valid: !isBad() && !spAlreadyFound.count(pMP). */
void fuse_sim3_search(const FrameView& KF, const Grid& gKF, const ProjPoint* pts, int npts, const float* Rcw,
                      const float* tcw, const float* Ow, float th, float mfLogScaleFactor, int mnScaleLevels,
                      int32_t* bestIdxOut, int32_t* bestDistOut)
{
    float T[12];
    for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) T[4 * r + k] = Rcw[3 * r + k]; T[4 * r + 3] = tcw[r]; }
    for (int iMP = 0; iMP < npts; ++iMP) {
        bestIdxOut[iMP] = -1; bestDistOut[iMP] = 256;
        if (!pts[iMP].valid) continue;
        float p3Dc[3];
        rx_plus_t(T, pts[iMP].world, p3Dc);
        if (p3Dc[2] < 0.0f) continue;
        const float invz = (float)(1.0 / p3Dc[2]);                      /* :1176 */
        const float x = p3Dc[0] * invz, y = p3Dc[1] * invz;
        const float u = KF.fx * x + KF.cx, v = KF.fy * y + KF.cy;
        if (!(u >= KF.minX && u < KF.maxX && v >= KF.minY && v < KF.maxY)) continue;
        float PO[3];
        for (int k = 0; k < 3; ++k) PO[k] = pts[iMP].world[k] - Ow[k];
        const float dist3D = (float)norm3(PO);
        if (dist3D < pts[iMP].minDistance || dist3D > pts[iMP].maxDistance) continue;
        if (dot3(PO, pts[iMP].normal) < 0.5 * dist3D) continue;
        const int nPredictedLevel = predict_scale(pts[iMP].maxDistanceRaw, dist3D, mfLogScaleFactor, mnScaleLevels);
        const float radius = th * KF.scaleFactors[nPredictedLevel];
        const std::vector<int> vIndices = features_in_area(KF, gKF, u, v, radius, -1, -1);
        if (vIndices.empty()) continue;
        int bestDist = 0x7fffffff, bestIdx = -1;
        for (int idx : vIndices) {
            const int kpLevel = KF.keysUn[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int dist = descriptor_distance(pts[iMP].desc, KF.desc + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        bestIdxOut[iMP] = bestIdx; bestDistOut[iMP] = bestIdx >= 0 ? bestDist : 256;
    }
}

/* ORBmatcher::SearchBySim3, src/ORBmatcher.cc:1259-1483.  pts1 / pts2: the MapPoints of KeyFrame 1 / 2 per keypoint
 * (valid: pMP && !vbAlreadyMatched && !isBad()); T1w/T2w = [R|t] of the two keyframes; sR12|t12 and sR21|t21 the
 * similarity in both directions as the reference derives them (:1276-1278).  matches12[i1] = i2 or -1 (only the NEW
 * matches; the reference leaves earlier entries of vpMatches12 untouched). */
int search_by_sim3(const FrameView& KF1, const Grid& g1, const FrameView& KF2, const Grid& g2, const ProjPoint* pts1,
                   const ProjPoint* pts2, const float* T1w, const float* T2w, const float* S12, const float* S21, float th,
                   float mfLogScaleFactor, int mnScaleLevels, int32_t* matches12)
{
    const int N1 = KF1.N, N2 = KF2.N;
    std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
    auto pass = [&](const ProjPoint* pts, int n, const float* Tsrc, const float* S, const FrameView& dst, const Grid& g, std::vector<int>& out) {
        for (int i = 0; i < n; ++i) {
            if (!pts[i].valid) continue;
            float pa[3], pb[3];
            rx_plus_t(Tsrc, pts[i].world, pa);                          /* p3Dc1 = R1w*p3Dw + t1w */
            rx_plus_t(S, pa, pb);                                       /* p3Dc2 = sR21*p3Dc1 + t21 */
            if (pb[2] < 0.0) continue;
            const float invz = (float)(1.0 / pb[2]);
            const float x = pb[0] * invz, y = pb[1] * invz;
            const float u = dst.fx * x + dst.cx, v = dst.fy * y + dst.cy;
            if (!(u >= dst.minX && u < dst.maxX && v >= dst.minY && v < dst.maxY)) continue;
            const float dist3D = (float)norm3(pb);
            if (dist3D < pts[i].minDistance || dist3D > pts[i].maxDistance) continue;
            const int nPredictedLevel = predict_scale(pts[i].maxDistanceRaw, dist3D, mfLogScaleFactor, mnScaleLevels);
            const float radius = th * dst.scaleFactors[nPredictedLevel];
            const std::vector<int> vIndices = features_in_area(dst, g, u, v, radius, -1, -1);
            if (vIndices.empty()) continue;
            int bestDist = 0x7fffffff, bestIdx = -1;
            for (int idx : vIndices) {
                const int oct = dst.keysUn[idx].octave;
                if (oct < nPredictedLevel - 1 || oct > nPredictedLevel) continue;
                const int dist = descriptor_distance(pts[i].desc, dst.desc + 32 * (size_t)idx);
                if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
            }
            if (bestDist <= TH_HIGH) out[i] = bestIdx;
        }
    };
    pass(pts1, N1, T1w, S21, KF2, g2, vnMatch1);
    pass(pts2, N2, T2w, S12, KF1, g1, vnMatch2);
    int nFound = 0;
    for (int i1 = 0; i1 < N1; ++i1) {                                   /* :1462-1478 */
        matches12[i1] = -1;
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1) { matches12[i1] = idx2; ++nFound; }
    }
    return nFound;
}

/* ORBmatcher::CheckDistEpipolarLine, src/ORBmatcher.cc:140-157 */
static bool check_dist_epipolar_line(const KeyPoint& kp1, const KeyPoint& kp2, const float* F12, const float* mvLevelSigma2)
{
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * mvLevelSigma2[kp2.octave];
}

/* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:814-980.  hasMp1/2: pKF->GetMapPoint(idx) != NULL; ex, ey: the
 * epipole of :823-829.  matches12 = vMatches12 after the rotation cull (vMatchedPairs lists its non-negative entries). */
int search_for_triangulation(const FrameView& KF1, const uint8_t* hasMp1, const FeatureVec& a, const FrameView& KF2,
                             const uint8_t* hasMp2, const FeatureVec& b, const float* F12, float ex, float ey,
                             const float* mvLevelSigma2, bool bOnlyStereo, bool checkOri, int32_t* matches12)
{
    int nmatches = 0;
    std::vector<bool> vbMatched2(KF2.N, false);                         /* never set to true by the reference */
    std::fill(matches12, matches12 + KF1.N, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    int ia = 0, ib = 0;
    while (ia < a.nnodes && ib < b.nnodes) {
        if (a.nodeId[ia] == b.nodeId[ib]) {
            for (int p = a.offset[ia]; p < a.offset[ia + 1]; ++p) {
                const unsigned idx1 = a.index[p];
                if (hasMp1[idx1]) continue;
                const bool bStereo1 = KF1.uRight && KF1.uRight[idx1] >= 0;
                if (bOnlyStereo) if (!bStereo1) continue;
                const KeyPoint& kp1 = KF1.keysUn[idx1];
                const uint8_t* d1 = KF1.desc + 32 * (size_t)idx1;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int q = b.offset[ib]; q < b.offset[ib + 1]; ++q) {
                    const unsigned idx2 = b.index[q];
                    if (vbMatched2[idx2] || hasMp2[idx2]) continue;
                    const bool bStereo2 = KF2.uRight && KF2.uRight[idx2] >= 0;
                    if (bOnlyStereo) if (!bStereo2) continue;
                    const int dist = descriptor_distance(d1, KF2.desc + 32 * (size_t)idx2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const KeyPoint& kp2 = KF2.keysUn[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x, distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * KF2.scaleFactors[kp2.octave]) continue;
                    }
                    if (check_dist_epipolar_line(kp1, kp2, F12, mvLevelSigma2)) { bestIdx2 = (int)idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    matches12[idx1] = bestIdx2;
                    ++nmatches;
                    if (checkOri) rotHist[rot_bin(kp1.angle - KF2.keysUn[bestIdx2].angle)].push_back((int)idx1);
                }
            }
            ++ia; ++ib;
        } else if (a.nodeId[ia] < b.nodeId[ib]) {
            ia = (int)(std::lower_bound(a.nodeId, a.nodeId + a.nnodes, b.nodeId[ib]) - a.nodeId);
        } else {
            ib = (int)(std::lower_bound(b.nodeId, b.nodeId + b.nnodes, a.nodeId[ia]) - b.nodeId);
        }
    }
    if (checkOri) {
        int sizes[HISTO_LENGTH], ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx : rotHist[i]) { matches12[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

}  // namespace orc
