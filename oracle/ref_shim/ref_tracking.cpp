/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Tracking::Separate / classifyH / classifyF compiled from the reference's own text (src/Tracking.cc:1093-1367, cut at
 * build time into _ref/gen/) over the stand-in class of ref_tracking_mock.h.  With -DREF_DROPIN the members come from the
 * product's host/Tracking_Separate.cc instead (oracle/_ref/libdropin.so) and only the glue below is compiled. */
#include "ref_tracking_mock.h"

#ifndef REF_DROPIN
namespace ORB_SLAM2 {
#include "gen/Tracking_1093_1367.inc"
}  // namespace ORB_SLAM2
#endif
/* entry points used by ref_api_frame.cpp */
namespace refapi {
int tracking_separate(ORB_SLAM2::Frame& cur, ORB_SLAM2::Frame& ref, ORB_SLAM2::Frame& last, const cv::Mat& HorF, int flag,
                      std::vector<std::vector<int>>& dynStatus, std::vector<int>& curStatusOut)
{
    ORB_SLAM2::Tracking T;
    T.mCurrentFrame = cur;          /* Frame copy constructor (src/Frame.cc:38-58) — what Tracking itself holds */
    T.mRefFrame = &ref;
    T.mLastFrame = last;
    const int r = T.Separate(HorF, flag, dynStatus);
    curStatusOut = T.mCurrentFrame.box_status;
    return r;
}
void classify(int flag, const cv::Mat& M, const std::vector<cv::KeyPoint>& cur, const std::vector<cv::KeyPoint>& ref,
              std::vector<cv::DMatch>& matches, std::vector<int>& falseDyn)
{
#ifndef REF_DROPIN
    ORB_SLAM2::Tracking T;
    if (flag == 1) T.classifyH(M, cur, ref, matches, falseDyn);
    else T.classifyF(M, cur, ref, matches, falseDyn);
#else
    (void)flag; (void)M; (void)cur; (void)ref; (void)matches; (void)falseDyn;      /* subsumed by Separate in the drop-in */
#endif
}
}  // namespace refapi
