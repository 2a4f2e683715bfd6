/* Drop-in replacement of the reference's src/ORBmatcher.cc: every member of ORB_SLAM2::ORBmatcher as declared in the reference's
 * OWN include/ORBmatcher.h (:41-95, unchanged), implemented on the sm_100a kernels behind the sdyn C ABI.  Compile this file in
 * place of src/ORBmatcher.cc; Tracking, LocalMapping, LoopClosing and Frame keep calling the same signatures.
 *
 * Each member gathers the fields the search reads out of the reference's Frame / KeyFrame / MapPoint objects, runs the search
 * on the device and writes the matches back as pointers (host/sdyn_adapters.hpp holds the gathering code as templates so that
 * the same text is type-checked against the repository's stand-ins).  Without a GPU context there is no fallback: the
 * searches report the failure on stderr and return 0 matches. */
#include "ORBmatcher.h"
#include "sdyn_adapters.hpp"
#include "sdyn_context.h"

#include <climits>

using namespace std;

namespace ORB_SLAM2
{

const int ORBmatcher::TH_HIGH = SDYN_TH_HIGH;        /* src/ORBmatcher.cc:37-39 */
const int ORBmatcher::TH_LOW = SDYN_TH_LOW;
const int ORBmatcher::HISTO_LENGTH = SDYN_HISTO_LENGTH;

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

/* static, pair-at-a-time (MapPoint.cc:281, Frame.cc:949): host popcount */
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return sdyn_hamming(a.data, b.data); }

int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)
{
    return sdyn_host::SearchByProjection(sdyn_host::ThreadContext(), F, vpMapPoints, th, mfNNratio);
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
{
    return sdyn_host::SearchByProjection<Frame, cv::Point2f>(sdyn_host::ThreadContext(), CurrentFrame, LastFrame, th, bMono,
                                                             mbCheckOrientation, nullptr, nullptr);
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono,
                                   vector<cv::Point2f>& points_last, vector<cv::Point2f>& points_current)
{
    return sdyn_host::SearchByProjection<Frame, cv::Point2f>(sdyn_host::ThreadContext(), CurrentFrame, LastFrame, th, bMono,
                                                             mbCheckOrientation, &points_last, &points_current);
}

int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist)
{
    return sdyn_host::SearchByProjection(sdyn_host::ThreadContext(), CurrentFrame, pKF, sAlreadyFound, th, ORBdist, mbCheckOrientation);
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)
{
    return sdyn_host::SearchByProjection(sdyn_host::ThreadContext(), pKF, Scw, vpPoints, vpMatched, th);
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
{
    return sdyn_host::SearchByBoW(sdyn_host::ThreadContext(), pKF, F, vpMapPointMatches, mfNNratio, mbCheckOrientation);
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
{
    return sdyn_host::SearchByBoW(sdyn_host::ThreadContext(), pKF1, pKF2, vpMatches12, mfNNratio, mbCheckOrientation);
}

int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize)
{
    return sdyn_host::SearchForInitialization(sdyn_host::ThreadContext(), F1, F2, vbPrevMatched, vnMatches12, windowSize, mfNNratio,
                                              mbCheckOrientation);
}

int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo)
{
    return sdyn_host::SearchForTriangulation(sdyn_host::ThreadContext(), pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo,
                                             mbCheckOrientation);
}

int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th)
{
    return sdyn_host::SearchBySim3(sdyn_host::ThreadContext(), pKF1, pKF2, vpMatches12, s12, R12, t12, th);
}

int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th)
{
    return sdyn_host::Fuse(sdyn_host::ThreadContext(), pKF, vpMapPoints, th);
}

int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
{
    return sdyn_host::Fuse(sdyn_host::ThreadContext(), pKF, Scw, vpPoints, th, vpReplacePoint);
}

/* The three protected helpers of the class (declared in the header; the searches above do not go through them). */
float ORBmatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5 : 4.0; }      /* src/ORBmatcher.cc:131-137 */

bool ORBmatcher::CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF2)
{
    /* src/ORBmatcher.cc:140-157: distance of kp2 to the epipolar line l = x1' F12, against 3.84 sigma^2 of kp2's level */
    const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);
    const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
    const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
    const float num = a * kp2.pt.x + b * kp2.pt.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * pKF2->mvLevelSigma2[kp2.octave];
}

void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3)
{
    /* src/ORBmatcher.cc:1758-1799 */
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

}  // namespace ORB_SLAM2
