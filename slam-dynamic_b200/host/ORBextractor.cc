/* See ORBextractor.h.  Maps ORBextractor::operator() (reference: src/ORBextractor.cc:1043-1105) onto
 * sdyn_extract; errors of the C ABI become the reference's behaviour (silent return / assert). */
#include "ORBextractor.h"
#include "../../include/sdyn.h"
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace ORB_SLAM2
{

static int g_device = 0;
void ORBextractor::SetDevice(int device) { g_device = device; }

ORBextractor::ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
    : mFeatures(nfeatures), mLevels(nlevels), mIniTh(iniThFAST), mMinTh(minThFAST), mScaleFactor(scaleFactor),
      mCtx(nullptr), mCtxW(0), mCtxH(0)
{
    sdyn_orb_params p = {nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST};
    sdyn_scale_info si;
    int32_t umax[16];
    const int rc = sdyn_orb_tables(&p, &si, umax);
    assert(rc == SDYN_OK && "ORBextractor: unsupported parameters");
    (void)rc;
    mvScaleFactor.assign(si.scale, si.scale + nlevels);
    mvInvScaleFactor.assign(si.inv_scale, si.inv_scale + nlevels);
    mvLevelSigma2.assign(si.sigma2, si.sigma2 + nlevels);
    mvInvLevelSigma2.assign(si.inv_sigma2, si.inv_sigma2 + nlevels);
    mnFeaturesPerLevel.assign(si.features_per_level, si.features_per_level + nlevels);
    mvImagePyramid.resize(nlevels);
    mBordered.resize(nlevels);
}

ORBextractor::~ORBextractor() { sdyn_destroy(mCtx); }

void ORBextractor::EnsureContext(int width, int height)
{
    if (mCtx && width <= mCtxW && height <= mCtxH) return;
    /* grow to the envelope of everything seen so far: inputs alternating between a wide and a tall image settle after one
     * re-creation instead of thrashing */
    if (mCtx) { width = std::max(width, mCtxW); height = std::max(height, mCtxH); }
    sdyn_destroy(mCtx);
    mCtx = nullptr;
    sdyn_orb_params p = {mFeatures, (float)mScaleFactor, mLevels, mIniTh, mMinTh};
    if (sdyn_create(&p, width, height, 1, g_device, &mCtx) != SDYN_OK) {
        /* no CPU fallback exists: report and leave mCtx null; operator() then returns without output */
        std::fprintf(stderr, "ORBextractor: %s\n", sdyn_last_error(nullptr));
        mCtx = nullptr;
        return;
    }
    mCtxW = width; mCtxH = height;
    if (mHaveCamera &&
        sdyn_set_camera(mCtx, mCamera[0], mCamera[1], mCamera[2], mCamera[3], mDistCoef.data(), (int)mDistCoef.size()) != SDYN_OK)
        std::fprintf(stderr, "ORBextractor: camera model lost on context re-creation: %s\n", sdyn_last_error(mCtx));
    const int cap = sdyn_max_keypoints(mCtx);
    mStageKp.resize(cap);
    mStageDesc.resize((size_t)cap * 32);
}

int ORBextractor::SetCamera(float fx, float fy, float cx, float cy, const float* distCoef, int nCoef)
{
    mHaveCamera = true;
    mCamera[0] = fx; mCamera[1] = fy; mCamera[2] = cx; mCamera[3] = cy;
    mDistCoef.assign(distCoef, distCoef + (distCoef ? nCoef : 0));
    return mCtx ? sdyn_set_camera(mCtx, fx, fy, cx, cy, mDistCoef.data(), (int)mDistCoef.size()) : SDYN_OK;   /* else applied on creation */
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors)
{
    if (_image.empty()) return;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);

    EnsureContext(image.cols, image.rows);
    if (!mCtx) return;

    /* level sizes follow cvRound(size * invScale) like ComputePyramid; allocate the bordered owners and hand
     * their interiors out as mvImagePyramid, exactly the ROI-of-a-bordered-Mat shape of the reference */
    uint8_t* pyr[SDYN_MAX_LEVELS] = {nullptr};
    for (int l = 0; l < mLevels && mPyramidOnHost; ++l) {
        const float s = mvInvScaleFactor[l];
        const int w = (int)std::lrintf((float)image.cols * s), h = (int)std::lrintf((float)image.rows * s);
        mBordered[l].create(h + 2 * SDYN_EDGE, w + 2 * SDYN_EDGE, CV_8UC1);
        mvImagePyramid[l] = mBordered[l](cv::Rect(SDYN_EDGE, SDYN_EDGE, w, h));
        pyr[l] = mBordered[l].data;
    }

    int n = 0;
    const int cap = (int)mStageKp.size();
    static_assert(sizeof(cv::KeyPoint) == sizeof(sdyn_keypoint), "cv::KeyPoint and sdyn_keypoint must share a layout");
    const int rc = sdyn_extract(mCtx, image.data, image.cols, image.rows, (int)image.step,
                                reinterpret_cast<sdyn_keypoint*>(mStageKp.data()), mStageDesc.data(), cap, &n,
                                mPyramidOnHost ? pyr : nullptr);
    if (rc != SDYN_OK) {
        std::fprintf(stderr, "ORBextractor: %s\n", sdyn_last_error(mCtx));
        assert(rc == SDYN_OK);
        return;
    }
    if (n == 0)
        _descriptors.release();
    else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), mStageDesc.data() + (size_t)32 * i, 32);
    }
    _keypoints.clear();
    _keypoints.reserve(n);
    _keypoints.insert(_keypoints.end(), mStageKp.begin(), mStageKp.begin() + n);
}

}  // namespace ORB_SLAM2
