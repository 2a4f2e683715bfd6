/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Force-included (-include) in front of the reference's src/Frame.cc and src/ORBmatcher.cc when they are compiled
 * unchanged into oracle/_ref/libref.so.  The reference's include/Frame.h, ORBmatcher.h and ORBextractor.h are used as
 * they are; the entity headers they pull in (MapPoint.h, KeyFrame.h, ORBVocabulary.h, Converter.h — which drag in the
 * map graph, g2o, Eigen and DBoW2's vocabulary) are replaced by the minimal stand-ins below by pre-defining their
 * include guards.  The stand-ins carry exactly the members the hot path reads.  Member functions with arithmetic in
 * them are NOT restated here: their bodies are cut from the reference's own sources at build time
 * (oracle/Makefile: MapPoint.cc:373-417, KeyFrame.cc:70-121 and :569-613 -> _ref/gen/*.inc, compiled by ref_entities.cpp).
 */
#pragma once
#define MAPPOINT_H
#define KEYFRAME_H
#define ORBVOCABULARY_H
#define CONVERTER_H

#include <opencv2/core/core.hpp>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

/* include/Frame.h writes `vector<...>` unqualified: in the reference build the directive arrives through
 * ORBVocabulary.h -> Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:36. */
using namespace std;

namespace ORB_SLAM2 {

class Frame;
class KeyFrame;
class Map;
class KeyFrameDatabase;

/* include/ORBVocabulary.h: typedef of DBoW2::TemplatedVocabulary; Frame::ComputeBoW only calls transform() */
class ORBVocabulary {
public:
    virtual ~ORBVocabulary() {}
    virtual void transform(const std::vector<cv::Mat>& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const;
};

/* include/Converter.h (only the descriptor splitter is on this path; src/Converter.cc:27-35) */
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat& Descriptors);
};

/* include/MapPoint.h:36-150, reduced to what ORBmatcher.cc / Frame.cc touch */
class MapPoint {
public:
    MapPoint(const cv::Mat& Pos, const cv::Mat& normal, const cv::Mat& desc, float minDist, float maxDist, int nobs, bool bad);
    cv::Mat GetWorldPos();
    cv::Mat GetNormal();
    cv::Mat GetDescriptor();
    int Observations();
    bool isBad();
    void AddObservation(KeyFrame* pKF, size_t idx);
    int GetIndexInKeyFrame(KeyFrame* pKF);
    bool IsInKeyFrame(KeyFrame* pKF);
    void Replace(MapPoint* pMP);
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    int PredictScale(const float& currentDist, Frame* pF);

    long unsigned int mnId;
    static long unsigned int nNextId;
    int nObs;
    float mTrackProjX, mTrackProjY, mTrackProjXR;
    bool mbTrackInView;
    int mnTrackScaleLevel;
    float mTrackViewCos;
    long unsigned int mnTrackReferenceForFrame, mnLastFrameSeen, mnFuseCandidateForKF;

    MapPoint* mpReplaced;            /* recorded by Replace() for the tests */
    std::map<KeyFrame*, size_t> mObservations;

protected:
    cv::Mat mWorldPos, mNormalVector, mDescriptor;
    bool mbBad;
    float mfMinDistance, mfMaxDistance;
    std::mutex mMutexPos, mMutexFeatures;
};

/* include/KeyFrame.h:49-240, reduced to what ORBmatcher.cc touches; constructed from a Frame as KeyFrame.cc:30-58 does */
class KeyFrame {
public:
    KeyFrame(Frame& F, Map* pMap, KeyFrameDatabase* pKFDB);
    void SetPose(const cv::Mat& Tcw);
    cv::Mat GetPose();
    cv::Mat GetPoseInverse();
    cv::Mat GetCameraCenter();
    cv::Mat GetStereoCenter();
    cv::Mat GetRotation();
    cv::Mat GetTranslation();
    void AddMapPoint(MapPoint* pMP, const size_t& idx);
    std::set<MapPoint*> GetMapPoints();
    std::vector<MapPoint*> GetMapPointMatches();
    MapPoint* GetMapPoint(const size_t& idx);
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    bool IsInImage(const float& x, const float& y) const;
    bool isBad();

    static long unsigned int nNextId;
    long unsigned int mnId;
    const long unsigned int mnFrameId;
    const int mnGridCols, mnGridRows;
    const float mfGridElementWidthInv, mfGridElementHeightInv;
    const float fx, fy, cx, cy, invfx, invfy, mbf, mb, mThDepth;
    const int N;
    const std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    const std::vector<float> mvuRight, mvDepth;
    const cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    const int mnScaleLevels;
    const float mfScaleFactor, mfLogScaleFactor;
    const std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    const int mnMinX, mnMinY, mnMaxX, mnMaxY;

    std::vector<MapPoint*> mvpMapPoints;
protected:
    cv::Mat Tcw, Twc, Ow, Cw;
    std::vector<std::vector<std::vector<size_t>>> mGrid;
    bool mbBad;
    float mHalfBaseline;
    std::mutex mMutexPose, mMutexConnections, mMutexFeatures;
};

}  // namespace ORB_SLAM2
