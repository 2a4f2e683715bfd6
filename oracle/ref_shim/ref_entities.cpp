/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Stand-in MapPoint / KeyFrame / ORBVocabulary / Converter for the reference TUs compiled into oracle/_ref/libref.so
 * (see ref_entities.h).  Everything that computes is NOT written here: it is the reference's own text, cut at build time
 * from /root/reference/src into _ref/gen/ (oracle/Makefile) and included below. */
#include "ref_entities.h"
#include "Frame.h"
#include <cmath>

using namespace std;

namespace ORB_SLAM2 {

void ORBVocabulary::transform(const std::vector<cv::Mat>&, DBoW2::BowVector&, DBoW2::FeatureVector&, int) const {}

/* src/Converter.cc:27-35 */
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat& Descriptors)
{
    std::vector<cv::Mat> v;
    v.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
}

long unsigned int MapPoint::nNextId = 0;

MapPoint::MapPoint(const cv::Mat& Pos, const cv::Mat& normal, const cv::Mat& desc, float minDist, float maxDist, int nobs, bool bad)
    : mnId(nNextId++), nObs(nobs), mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), mbTrackInView(false), mnTrackScaleLevel(0),
      mTrackViewCos(0), mnTrackReferenceForFrame(0), mnLastFrameSeen(0), mnFuseCandidateForKF(0), mpReplaced(nullptr),
      mWorldPos(Pos.clone()), mNormalVector(normal.clone()), mDescriptor(desc.clone()), mbBad(bad), mfMinDistance(minDist), mfMaxDistance(maxDist)
{}

/* accessors: src/MapPoint.cc returns clones under a lock; nothing is computed */
cv::Mat MapPoint::GetWorldPos() { return mWorldPos.clone(); }
cv::Mat MapPoint::GetNormal() { return mNormalVector.clone(); }
cv::Mat MapPoint::GetDescriptor() { return mDescriptor.clone(); }
int MapPoint::Observations() { return nObs; }
bool MapPoint::isBad() { return mbBad; }
void MapPoint::AddObservation(KeyFrame* pKF, size_t idx) { if (!mObservations.count(pKF)) { mObservations[pKF] = idx; ++nObs; } }
int MapPoint::GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
bool MapPoint::IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
void MapPoint::Replace(MapPoint* pMP) { mpReplaced = pMP; }

/* MapPoint::GetMinDistanceInvariance, GetMaxDistanceInvariance, PredictScale x2 — reference text, src/MapPoint.cc:373-417 */
#include "gen/MapPoint_373_417.inc"

long unsigned int KeyFrame::nNextId = 0;

/* src/KeyFrame.cc:30-58, member for member (the map / database / covisibility members do not exist here) */
KeyFrame::KeyFrame(Frame& F, Map*, KeyFrameDatabase*)
    : mnFrameId(F.mnId), mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS), mfGridElementWidthInv(F.mfGridElementWidthInv),
      mfGridElementHeightInv(F.mfGridElementHeightInv), fx(F.fx), fy(F.fy), cx(F.cx), cy(F.cy), invfx(F.invfx), invfy(F.invfy),
      mbf(F.mbf), mb(F.mb), mThDepth(F.mThDepth), N(F.N), mvKeys(F.mvKeys), mvKeysUn(F.mvKeysUn), mvuRight(F.mvuRight),
      mvDepth(F.mvDepth), mDescriptors(F.mDescriptors.clone()), mBowVec(F.mBowVec), mFeatVec(F.mFeatVec),
      mnScaleLevels(F.mnScaleLevels), mfScaleFactor(F.mfScaleFactor), mfLogScaleFactor(F.mfLogScaleFactor),
      mvScaleFactors(F.mvScaleFactors), mvLevelSigma2(F.mvLevelSigma2), mvInvLevelSigma2(F.mvInvLevelSigma2),
      mnMinX(F.mnMinX), mnMinY(F.mnMinY), mnMaxX(F.mnMaxX), mnMaxY(F.mnMaxY), mvpMapPoints(F.mvpMapPoints), mbBad(false),
      mHalfBaseline(F.mb / 2)
{
    mnId = nNextId++;
    mGrid.resize(mnGridCols);
    for (int i = 0; i < mnGridCols; i++) {
        mGrid[i].resize(mnGridRows);
        for (int j = 0; j < mnGridRows; j++) mGrid[i][j] = F.mGrid[i][j];
    }
    SetPose(F.mTcw);
}

void KeyFrame::AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
std::vector<MapPoint*> KeyFrame::GetMapPointMatches() { return mvpMapPoints; }
MapPoint* KeyFrame::GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
bool KeyFrame::isBad() { return mbBad; }
/* KeyFrame::GetMapPoints — reference text, src/KeyFrame.cc:235-248 */
#include "gen/KeyFrame_235_248.inc"

/* KeyFrame::SetPose ... GetTranslation — reference text, src/KeyFrame.cc:70-121 */
#include "gen/KeyFrame_70_121.inc"
/* KeyFrame::GetFeaturesInArea, IsInImage — reference text, src/KeyFrame.cc:569-613 */
#include "gen/KeyFrame_569_613.inc"

}  // namespace ORB_SLAM2
