/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * CPU restatement of the ORBmatcher searches on the hot path and of the Frame grid they use
 * (reference: src/ORBmatcher.cc, src/Frame.cc).  Pointers of the reference become indices:
 *   MapPoint* stored in Frame::mvpMapPoints[i]   ->  assign[i]  (-1 = NULL, >= 0 = index of the query)
 *   pMP->Observations() > 0                      ->  a per-query flag; locked[i] mirrors it for the occupant
 *
 * Parity pin: the reference ships no tests or vectors for these functions and cannot be built here, so they are
 * pinned by independent Python re-statements that run the OpenCV-dependent steps through cv2 itself
 * (tests/test_oracle_matcher.py: GetFeaturesInArea, SearchByProjection(F,MPs), SearchByProjection(Cur,Last),
 * SearchForInitialization, SearchByBoW(KF,F) and (KF,KF), both pose searches, the Fuse search, SearchForTriangulation;
 * cv2.gemm / cv2.norm / BFMatcher / invert known answers) and by the regression hashes of
 * tests/golden/match_oracle.json.  Fuse(Scw) and SearchBySim3 share every step with a pinned function (projection,
 * gates, PredictScale, window walk) but have no twin of their own.
 */
#pragma once
#include "orc_extractor.h"
#include <cstdint>
#include <vector>

namespace orc {

constexpr int GRID_COLS = 64, GRID_ROWS = 48;   /* include/Frame.h:39-40 */
constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;   /* src/ORBmatcher.cc:37-39 */

int descriptor_distance(const uint8_t* a, const uint8_t* b);   /* src/ORBmatcher.cc:1804-1820 */

/* What the searches read from a Frame. */
struct FrameView {
    int N = 0;
    const KeyPoint* keys = nullptr;      /* mvKeys */
    const KeyPoint* keysUn = nullptr;    /* mvKeysUn */
    const uint8_t* desc = nullptr;       /* mDescriptors, N x 32 */
    const float* uRight = nullptr;       /* mvuRight (nullptr = all -1, monocular) */
    float minX = 0, minY = 0, maxX = 0, maxY = 0;   /* mnMinX .. mnMaxY */
    float gridWInv = 0, gridHInv = 0;    /* mfGridElementWidthInv / HeightInv */
    int nlevels = 0;
    const float* scaleFactors = nullptr; /* mvScaleFactors */
    float fx = 0, fy = 0, cx = 0, cy = 0, bf = 0, b = 0;   /* fx, fy, cx, cy, mbf, mb */
    float Tcw[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};  /* rows 0..2 of mTcw, row-major */
};

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];   /* mGrid[x][y] */
};

void grid_bounds(FrameView& f);                                       /* Frame.cc:383-385: inverse cell sizes */
void assign_features_to_grid(const FrameView& f, Grid& g, int from = 0);   /* Frame.cc:463-478, 643-653 */
std::vector<int> features_in_area(const FrameView& f, const Grid& g, float x, float y, float r,
                                  int minLevel, int maxLevel);      /* Frame.cc:735-788 */

struct MapPointQuery {          /* the MapPoint fields SearchByProjection(F, vpMapPoints) reads */
    float projX, projY, projXR; /* mTrackProjX / Y / XR */
    float viewCos;              /* mTrackViewCos */
    int32_t level;              /* mnTrackScaleLevel */
    uint8_t trackInView, bad, obsPositive, pad;
    uint8_t desc[32];           /* GetDescriptor() */
};

/* ORBmatcher.cc:45-129 */
int search_by_projection_map(const FrameView& F, const Grid& g, const MapPointQuery* mps, int nmp, float th,
                             float nnratio, int32_t* assign, uint8_t* locked, int assignBase = 0);

struct LastFramePoint {         /* per keypoint of LastFrame */
    uint8_t hasMP, outlier, obsPositive, pad;   /* mvpMapPoints[i] != NULL, mvbOutlier[i], Observations()>0 */
    float world[3];             /* pMP->GetWorldPos() */
    uint8_t desc[32];           /* pMP->GetDescriptor() */
};

/* ORBmatcher.cc:1485-1627 and the fork's overload :407-559 (pairs != nullptr).  pairs receives
 * (last.x, last.y, cur.x, cur.y) per accepted match, appended BEFORE the rotation cull (Appendix B-8). */
int search_by_projection_frame(const FrameView& Cur, const Grid& gCur, const FrameView& Last,
                               const LastFramePoint* lp, float th, bool mono, bool checkOri,
                               int32_t* assign, uint8_t* locked, std::vector<float>* pairs);

/* ORBmatcher.cc:562-677.  prevMatched: N1 x 2 floats, updated in place. */
int search_for_initialization(const FrameView& F1, const FrameView& F2, const Grid& g2, float* prevMatched,
                              int32_t* matches12, int windowSize, float nnratio, bool checkOri);

struct ProjPoint {              /* MapPoint fields of the pose-projection searches; layout of sdyn_proj_point */
    uint8_t valid, pad[3];
    float world[3], normal[3];
    float minDistance, maxDistance; /* GetMinDistanceInvariance(), GetMaxDistanceInvariance() */
    float maxDistanceRaw;           /* mfMaxDistance */
    float angle;                    /* pKF->mvKeysUn[i].angle */
    uint8_t desc[32];
};
/* ORBmatcher.cc:1629-1756 */
int search_by_projection_reloc(const FrameView& Cur, const Grid& gCur, const ProjPoint* pts, int npts, const float* Rcw,
                               const float* tcw, const float* Ow, float th, int ORBdist, bool checkOri,
                               float mfLogScaleFactor, int mnScaleLevels, int32_t* assign);
/* ORBmatcher.cc:290-403 */
int search_by_projection_sim3(const FrameView& KF, const Grid& gKF, const ProjPoint* pts, int npts, const float* Rcw,
                              const float* tcw, const float* Ow, int th, float mfLogScaleFactor, int mnScaleLevels, int32_t* assign);

/* ORBmatcher.cc:982-1100 / 1132-1237 (search parts of the two Fuse overloads) and :1259-1483 */
void fuse_search(const FrameView& KF, const Grid& gKF, const float* invLevelSigma2, const ProjPoint* pts, int npts,
                 const float* Rcw, const float* tcw, const float* Ow, float th, float mfLogScaleFactor, int mnScaleLevels,
                 int32_t* bestIdx, int32_t* bestDist);
void fuse_sim3_search(const FrameView& KF, const Grid& gKF, const ProjPoint* pts, int npts, const float* Rcw,
                      const float* tcw, const float* Ow, float th, float mfLogScaleFactor, int mnScaleLevels,
                      int32_t* bestIdx, int32_t* bestDist);
int search_by_sim3(const FrameView& KF1, const Grid& g1, const FrameView& KF2, const Grid& g2, const ProjPoint* pts1,
                   const ProjPoint* pts2, const float* T1w, const float* T2w, const float* S12, const float* S21, float th,
                   float mfLogScaleFactor, int mnScaleLevels, int32_t* matches12);

struct FeatureVec {             /* DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>> as CSR */
    int nnodes = 0;
    const uint32_t* nodeId = nullptr;   /* ascending */
    const int32_t* offset = nullptr;    /* nnodes + 1 */
    const uint32_t* index = nullptr;
};

/* ORBmatcher.cc:159-288.  kfValid[i] = (pMP != NULL && !pMP->isBad()) for keyframe keypoint i.
 * assign[fIdx] = kfIdx | -1. */
int search_by_bow(const FrameView& KF, const uint8_t* kfValid, const FeatureVec& fvKF, const FrameView& F,
                  const FeatureVec& fvF, float nnratio, bool checkOri, int32_t* assign);
/* ORBmatcher.cc:679-812 */
int search_by_bow_kf(const FrameView& KF1, const uint8_t* valid1, const FeatureVec& fv1, const FrameView& KF2,
                     const uint8_t* valid2, const FeatureVec& fv2, float nnratio, bool checkOri, int32_t* matches12);

/* Frame::isInFrustum, src/Frame.cc:677-733 (+ MapPoint::PredictScale, GetMin/MaxDistanceInvariance, MapPoint.cc:373-417) for one
 * MapPoint: world / normal, mfMinDistance / mfMaxDistance; the frame's pose (rows 0..2 of mTcw), camera and pyramid constants.
 * Returns mbTrackInView and fills the five tracking fields when true. */
bool is_in_frustum(const FrameView& F, float mfLogScaleFactor, const float* world, const float* normal, float mfMinDistance,
                   float mfMaxDistance, float viewingCosLimit, float* projX, float* projY, float* projXR, int* level, float* viewCos);

/* ORBmatcher.cc:1758-1799 */
void compute_three_maxima(const int* histoSizes, int L, int& ind1, int& ind2, int& ind3);

/* ORBmatcher.cc:814-980 (+ CheckDistEpipolarLine :140-157) */
int search_for_triangulation(const FrameView& KF1, const uint8_t* hasMp1, const FeatureVec& fv1, const FrameView& KF2,
                             const uint8_t* hasMp2, const FeatureVec& fv2, const float* F12, float ex, float ey,
                             const float* mvLevelSigma2, bool bOnlyStereo, bool checkOri, int32_t* matches12);

}  // namespace orc
