/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Tracking::Separate / classifyH / classifyF compiled from the reference's own text (src/Tracking.cc:1093-1367, cut at
 * build time into _ref/gen/).  The rest of Tracking.cc (viewer, PCL, local mapping, the state machine) cannot be compiled
 * here, so the class below declares only the three member functions and the three frames they read, with the
 * declarations of include/Tracking.h:103,149-151,225. */
#define TRACKING_H
#include "ref_entities.h"
#include "Frame.h"
#include "ORBmatcher.h"
#include <iostream>
#include <string>

using namespace std;

namespace ORB_SLAM2 {

class Tracking {
public:
    Frame mCurrentFrame, *mRefFrame;
    int Separate(cv::Mat HorF, int flag, vector<vector<int>>& dynStatus);
    void classifyH(const cv::Mat& H21, const vector<cv::KeyPoint>& cur_kpts, const vector<cv::KeyPoint>& ref_kpts, vector<cv::DMatch>& matches, vector<int>& falseDyn);
    void classifyF(const cv::Mat& F21, const vector<cv::KeyPoint>& cur_kpts, const vector<cv::KeyPoint>& ref_kpts, vector<cv::DMatch>& matches, vector<int>& falseDyn);
    Frame mLastFrame;
};

#include "gen/Tracking_1093_1367.inc"

}  // namespace ORB_SLAM2

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
    ORB_SLAM2::Tracking T;
    if (flag == 1) T.classifyH(M, cur, ref, matches, falseDyn);
    else T.classifyF(M, cur, ref, matches, falseDyn);
}
}  // namespace refapi
