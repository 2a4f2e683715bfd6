/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Stand-in for include/Tracking.h (viewer, PCL, local mapping, the state machine cannot be compiled here): the three member
 * functions of the dynamic-keypoint path and the three frames they read, with the declarations of include/Tracking.h:103,
 * 149-151, 225.  Force-included in front of the TU that defines the members (the reference's own text in _ref/libref.so, the
 * product's host/Tracking_Separate.cc in _ref/libdropin.so). */
#pragma once
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

}  // namespace ORB_SLAM2
