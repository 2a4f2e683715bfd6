/* ORACLE — TEST INFRASTRUCTURE ONLY.  Frame::ComputeStereoMatches, src/Frame.cc:874-1048, statement by statement.
 * cv::Mat arithmetic is restated as the float / double operations OpenCV performs: convertTo(CV_32F) is an exact
 * u8 -> float conversion, `IL - IL.at(w,w) * ones` a float subtraction per element, and cv::norm(IL, IR, NORM_L1)
 * sums |a - b| (float difference) into a double accumulator (normDiffL1_32f); tests/test_oracle_stereo.py checks
 * this against cv2.norm on the same windows. */
#include "orc_stereo.h"
#include "orc_matcher.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <utility>
#include <vector>

namespace orc {

int compute_stereo_matches(const KeyPoint* keysL, int N, const uint8_t* descL,
                           const KeyPoint* keysR, int Nr, const uint8_t* descR,
                           const LevelRef* pyrL, const LevelRef* pyrR, int nlevels,
                           const float* mvScaleFactors, const float* mvInvScaleFactors,
                           float mb, float mbf, float* mvuRight, float* mvDepth)
{
    (void)nlevels;
    for (int i = 0; i < N; ++i) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }          /* :876-877 */
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                                       /* :879 */
    const int nRows = pyrL[0].h;                                                        /* :881 */

    /* :883-902  row table of the right keypoints */
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < Nr; ++iR) {
        const KeyPoint& kp = keysR[iR];
        const float kpY = kp.y;
        const float r = 2.0f * mvScaleFactors[kp.octave];
        const int maxr = (int)std::ceil(kpY + r);
        const int minr = (int)std::floor(kpY - r);
        if (minr < 0 || maxr >= nRows) return -1;       /* the reference writes out of bounds here */
        for (int yi = minr; yi <= maxr; ++yi) vRowIndices[yi].push_back(iR);
    }

    const float minZ = mb, minD = 0, maxD = mbf / minZ;                                 /* :905-907 */
    std::vector<std::pair<int, int>> vDistIdx;

    for (int iL = 0; iL < N; ++iL) {                                                    /* :913 */
        const KeyPoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        if (!(vL >= 0) || (size_t)vL >= (size_t)nRows) return -1;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];               /* :920 */
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;                                                         /* :931 */
        size_t bestIdxR = 0;
        const uint8_t* dL = descL + 32 * (size_t)iL;
        for (size_t iC = 0; iC < vCandidates.size(); ++iC) {                            /* :937-958 */
            const size_t iR = vCandidates[iC];
            const KeyPoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, descR + 32 * iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {                                                     /* :961 */
            const float uR0 = keysR[bestIdxR].x;
            const float scaleFactor = mvInvScaleFactors[kpL.octave];
            const float scaleduL = std::round(kpL.x * scaleFactor);
            const float scaledvL = std::round(kpL.y * scaleFactor);
            const float scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5;
            const LevelRef& PL = pyrL[kpL.octave];
            const LevelRef& PR = pyrR[kpL.octave];
            /* IL = rows [scaledvL-w, scaledvL+w+1) x cols [scaleduL-w, scaleduL+w+1), float, minus its centre */
            const int r0 = (int)(scaledvL - w), c0 = (int)(scaleduL - w);
            float IL[11][11];
            {
                const float centre = (float)PL.roi[(size_t)(r0 + w) * PL.stride + c0 + w];
                for (int y = 0; y < 11; ++y)
                    for (int x = 0; x < 11; ++x) IL[y][x] = (float)PL.roi[(size_t)(r0 + y) * PL.stride + c0 + x] - centre * 1.0f;
            }
            int bestDistW = INT_MAX;                                                    /* :976 (shadows bestDist) */
            int bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= PR.w) continue;                                     /* :982-984 */
            for (int incR = -L; incR <= +L; ++incR) {                                   /* :986-1000 */
                const int cr = (int)(scaleduR0 + incR - w);
                const float centre = (float)PR.roi[(size_t)(r0 + w) * PR.stride + cr + w];
                double acc = 0.0;
                for (int y = 0; y < 11; ++y)
                    for (int x = 0; x < 11; ++x) {
                        const float ir = (float)PR.roi[(size_t)(r0 + y) * PR.stride + cr + x] - centre * 1.0f;
                        acc += (double)std::fabs(IL[y][x] - ir);
                    }
                const float dist = (float)acc;
                if (dist < bestDistW) { bestDistW = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;                              /* :1002 */
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;                                    /* :1012 */
            float bestuR = mvScaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {                                /* :1020 */
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistW, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;      /* the reference reads vDistIdx[0] of an empty vector here (:1035) */
    std::sort(vDistIdx.begin(), vDistIdx.end());                                        /* :1034 */
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {                               /* :1038-1047 */
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
        --kept;
    }
    return kept;
}

}  // namespace orc
