/* Keypoint undistortion (K10): Frame::UndistortKeyPoints, src/Frame.cc:812-842.
 *   arithmetic: cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) — OpenCV calib3d, un-vendored: 5 fixed-point
 *   iterations of the Brown model in double, R = I, P = K (pinned against cv2 4.13 through the oracle).
 * One thread per keypoint; every double operation is a separate IEEE operation (the library is built with
 * -fmad=false), in the order OpenCV evaluates it, so the float results are bit-identical. */
#include "sdyn_internal.h"

namespace sdyn {

__device__ __forceinline__ void undistort_point(const CameraModel& cam, float xin, float yin, float& xout, float& yout)
{
    const double fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
    const double k1 = cam.k[0], k2 = cam.k[1], p1 = cam.k[2], p2 = cam.k[3], k3 = cam.k[4];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double u = xin, v = yin;
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        /* the rational terms k4..k6 and the thin-prism terms are zero for the 4/5-coefficient models of the reference:
         * 0*r2 + 0 etc. are exact, so only the terms below remain */
        const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
        if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
        const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
        const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    xout = (float)(fx * x + cx);
    yout = (float)(fy * y + cy);
}

__global__ void __launch_bounds__(256)
k_undistort(CameraModel cam, const sdyn_keypoint* __restrict__ kp, const int32_t* __restrict__ count, int cap,
            sdyn_keypoint* __restrict__ kpUn)
{
    const int f = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (i >= min(count[f], cap)) return;
    sdyn_keypoint k = kp[(size_t)f * cap + i];
    undistort_point(cam, k.x, k.y, k.x, k.y);
    kpUn[(size_t)f * cap + i] = k;
}

__global__ void k_undistort_xy(CameraModel cam, const float2* __restrict__ src, int n, float2* __restrict__ dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 p = src[i];
    float2 o;
    undistort_point(cam, p.x, p.y, o.x, o.y);
    dst[i] = o;
}

cudaError_t launch_undistort(const CameraModel& cam, const sdyn_keypoint* dKp, const int32_t* dCount, int cap,
                             sdyn_keypoint* dKpUn, int nframes, cudaStream_t st)
{
    dim3 grid((cap + 255) / 256, nframes);
    k_undistort<<<grid, 256, 0, st>>>(cam, dKp, dCount, cap, dKpUn);
    return cudaGetLastError();
}

cudaError_t launch_undistort_xy(const CameraModel& cam, const float* dSrc, int n, float* dDst, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    k_undistort_xy<<<(n + 255) / 256, 256, 0, st>>>(cam, reinterpret_cast<const float2*>(dSrc), n, reinterpret_cast<float2*>(dDst));
    return cudaGetLastError();
}

}  // namespace sdyn
