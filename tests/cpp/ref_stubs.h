/* Minimal stand-ins for the reference's Frame / MapPoint / KeyFrame (include/Frame.h, MapPoint.h, KeyFrame.h):
 * only the members the hot-path adapters touch, with the reference's names.  TEST INFRASTRUCTURE. */
#pragma once
#include "opencv2/core/core.hpp"
#include <map>
namespace ORB_SLAM2 { class ORBextractor; }
#include <vector>

namespace ref_stub {

class MapPoint {
public:
    bool mbTrackInView = false; int mnTrackScaleLevel = 0; float mTrackViewCos = 0, mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool bad = false; int nObs = 1; cv::Mat desc, pos;
    bool isBad() const { return bad; }
    int Observations() const { return nObs; }
    cv::Mat GetDescriptor() const { return desc.clone(); }
    cv::Mat GetWorldPos() const { return pos.clone(); }
};

class Frame {
public:
    int N = 0, mnScaleLevels = 8;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    cv::Mat mDescriptors, mTcw;
    std::vector<float> mvuRight, mvDepth, mvScaleFactors;
    std::vector<cv::KeyPoint> mvKeysRight;
    cv::Mat mDescriptorsRight;
    ORB_SLAM2::ORBextractor* mpORBextractorLeft = nullptr; ORB_SLAM2::ORBextractor* mpORBextractorRight = nullptr;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::map<unsigned, std::vector<unsigned>> mFeatVec;
    float mbf = 0, mb = 0;
    static float mnMinX, mnMinY, mnMaxX, mnMaxY, fx, fy, cx, cy;
};
float Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY, Frame::fx, Frame::fy, Frame::cx, Frame::cy;

}  // namespace ref_stub
