/* Drop-in replacement of the reference's ORBextractor (include/ORBextractor.h:45-111, src/ORBextractor.cc):
 * identical public surface — constructor, operator(), the six getters returning by value, the public
 * mvImagePyramid member — implemented on the sm_100a kernels behind the sdyn C ABI (include/sdyn.h).
 * Tracking / Frame / LocalMapping keep compiling and linking against it unchanged.
 *
 * Build: -DSDYN_HAVE_OPENCV with the real OpenCV headers, or add host/cv_stub to the include path. */
#ifndef SDYN_HOST_ORBEXTRACTOR_H
#define SDYN_HOST_ORBEXTRACTOR_H

#include <vector>
#ifdef SDYN_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#else
#include "opencv2/core/core.hpp"
#endif

struct sdyn_ctx;

namespace ORB_SLAM2
{

class ORBextractor
{
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    /* Same contract as the reference: mask is ignored; an empty image returns with the outputs untouched;
     * keypoints come out level-major in octree order with pt scaled to level 0; descriptors is N x 32 CV_8U. */
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                    cv::OutputArray descriptors);

    int GetLevels() { return mLevels; }
    float GetScaleFactor() { return (float)mScaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    /* Bordered pyramid of the last image, each entry a ROI of a (w+38) x (h+38) buffer, as in the reference
     * (Frame::ComputeStereoMatches reads it on the host). */
    std::vector<cv::Mat> mvImagePyramid;

    /* GPU context of this extractor (one per instance: left and right extractors run concurrently).  The pointer is valid
     * until the next operator() call: an image larger than every earlier one re-creates the context (EnsureContext), so
     * callers fetch it per frame and do not cache it. */
    sdyn_ctx* Context() { return mCtx; }
    /* Camera model for the device-side undistortion (Frame::UndistortKeyPoints).  Set it here rather than on Context()
     * directly: the extractor re-applies it whenever it has to re-create its context. */
    int SetCamera(float fx, float fy, float cx, float cy, const float* distCoef, int nCoef);
    /* mvImagePyramid is the one output that costs a 1.4 MB device-to-host copy per image.  Its only host reader in the
     * reference is Frame::ComputeStereoMatches (src/Frame.cc:881-988); once that runs on the device through
     * sdyn_host::ComputeStereoMatches the copy can be switched off (the levels stay available via sdyn_fetch_level). */
    void SetPyramidOnHost(bool on) { mPyramidOnHost = on; }
    /* Device the next context is created on (default 0); sharded deployments set it before the first frame. */
    static void SetDevice(int device);

protected:
    void EnsureContext(int width, int height);

    int mFeatures, mLevels, mIniTh, mMinTh;
    double mScaleFactor;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel;

    sdyn_ctx* mCtx;
    int mCtxW, mCtxH;
    bool mPyramidOnHost = true;
    bool mHaveCamera = false;
    float mCamera[4] = {0, 0, 0, 0};
    std::vector<float> mDistCoef;
    std::vector<cv::Mat> mBordered;          /* owners of the bordered level buffers */
    std::vector<cv::KeyPoint> mStageKp;      /* reused staging, sized once */
    std::vector<unsigned char> mStageDesc;
};

}  // namespace ORB_SLAM2
#endif
