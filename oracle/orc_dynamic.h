/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * CPU restatement of the fork's dynamic-keypoint rejection:
 *   Frame::boxTrack / firstSeparate / tail split / UpdateFrame   src/Frame.cc:481-653, :337-367
 *   Tracking::Separate / classifyH / classifyF                   src/Tracking.cc:1093-1367
 *   cv::BFMatcher(NORM_HAMMING, crossCheck=true)::match          SURVEY A-7 (OpenCV, un-vendored)
 *
 * Parity pin: BFMatcher cross-check and the 3x3 inverse are checked against cv2 (tests/test_oracle_matcher.py);
 * firstSeparate's erase-while-iterating behaviour and boxTrack have hand-derived known answers there; classifyF/H and
 * UpdateFrame are restatement-only ("parity unpinned": no reference binary, test or vector exists for them).
 */
#pragma once
#include "orc_extractor.h"
#include <cstdint>
#include <vector>

namespace orc {

struct Rect { double x, y, w, h; };                 /* cv::Rect2d */
struct Match { int queryIdx, trainIdx, dist; };     /* cv::DMatch (distance is an exact small integer) */

/* The per-frame box bookkeeping members of Frame (include/Frame.h) */
struct BoxState {
    std::vector<Rect> objects;
    std::vector<int> box_idx;
    std::vector<uint8_t> omit;
    std::vector<double> vel;            /* box_velocity, 2 per box */
    std::vector<int> box_status;
};

/* Frame.cc:481-552.  `boxes` may grow (boxes carried over from the last frame). */
void box_track(std::vector<Rect>& boxes, const BoxState& last, int imgW, int imgH, BoxState& cur);

/* Per-frame dynamic split, index based.  Input keypoints are the extractor output (order i = 0..N-1). */
struct SplitResult {
    std::vector<int> order;             /* new mvKeys order: order[k] = original index (static first) */
    int N_d = 0;                        /* #keypoints inside at least one box */
    std::vector<std::vector<int>> index;    /* per dynamic keypoint (in order of appearance): box list */
    std::vector<uint8_t> hasKpts;       /* per ORIGINAL box */
    bool empty_box = false;
    std::vector<int> class_id;          /* mvKeys[i].class_id after the call, per ORIGINAL keypoint index */
    /* after the tail split (Frame.cc:337-367): per remaining box, the original keypoint indices moved to it */
    std::vector<std::vector<int>> dynKeys;
};

/* Frame.cc:555-604 followed by the tail split :337-367.  boxes / cur are edited exactly as the reference
 * does (including the erase-while-iterating skip, Appendix B-4). */
SplitResult first_separate(const KeyPoint* keys, int N, std::vector<Rect>& boxes, BoxState& cur);

/* SURVEY A-7: strict mutual nearest neighbour, lowest index on ties, sorted by query index. */
std::vector<Match> bf_match_crosscheck(const uint8_t* q, int nq, const uint8_t* t, int nt);

/* Tracking.cc:1311-1367 / :1241-1309.  M = F21 or H21, row-major 3x3 float.  falseDyn[i] = queryIdx or -1. */
void classify_f(const float* F21, const float* curXY, const float* refXY, const std::vector<Match>& m, std::vector<int>& falseDyn);
void classify_h(const float* H21, const float* curXY, const float* refXY, const std::vector<Match>& m, std::vector<int>& falseDyn);
void invert3x3(const float* m, float* out);           /* cv::Mat::inv() for 3x3 CV_32F (LU path, closed form) */

struct BoxKeys {                                       /* one box's dynamic keypoints */
    std::vector<float> xy;                             /* mvdynKeysUn[box][k].pt */
    std::vector<uint8_t> desc;                         /* mdynDescriptors[box] */
};

/* Tracking.cc:1093-1239 without drawing / imwrite.  Returns 1 if any box was classified static.
 * dynStatus[box][m] = queryIdx | -1; matches[box] = the BF matches (for stage-level parity). */
int separate(const std::vector<BoxKeys>& cur, const std::vector<int>& curBoxIdx, std::vector<int>& curBoxStatus,
             const std::vector<BoxKeys>& ref, const std::vector<int>& refBoxIdx,
             const std::vector<int>& lastBoxIdx, const std::vector<int>& lastBoxStatus,
             const float* HorF, int flag, std::vector<std::vector<int>>& dynStatus,
             std::vector<std::vector<Match>>& matches);

/* Frame.cc:607-641: returns, in push order, (box, k) pairs re-admitted to the frame; dedupe on class_id. */
std::vector<std::pair<int, int>> update_frame(const std::vector<std::vector<int>>& dynStatus,
                                              const std::vector<std::vector<int>>& dynClassId);

}  // namespace orc
