/* Dynamic-keypoint rejection kernels (K9-K11).
 *   reference: Frame::firstSeparate (keypoint-in-box test)              src/Frame.cc:555-572
 *              Tracking::Separate: cv::BFMatcher(NORM_HAMMING, crossCheck) per box   src/Tracking.cc:1096,1122
 *              Tracking::classifyF / classifyH                           src/Tracking.cc:1311-1367, 1241-1309
 *   arithmetic: cv::Rect2d::contains (doubles), BFMatcher cross-check = strict mutual nearest neighbour with
 *              lowest-index ties (SURVEY A-7), float epipolar / transfer errors without FMA contraction.
 */
#include "match_internal.h"

namespace sdyn {

__global__ void __launch_bounds__(256)
k_box_mask(const sdyn_keypoint* __restrict__ keys, const int32_t* __restrict__ nPtr, int n, int keyStride,
           const double* __restrict__ boxes, const int32_t* __restrict__ nBoxesPtr, int nboxes, int boxStride,
           uint64_t* __restrict__ mask, size_t boxPitch, size_t nbPitch)
{
    __shared__ double sb[64 * 4];
    const int job = blockIdx.y;
    const int nb = min(nBoxesPtr ? *frame_part(nBoxesPtr, job, 4, (long long)nbPitch) : nboxes, 64);
    const int nk = nPtr ? min(nPtr[job], n) : n;
    const double* jb = frame_part(boxes, job, (size_t)boxStride * 32, (long long)boxPitch);
    for (int i = threadIdx.x; i < nb * 4; i += blockDim.x) sb[i] = jb[i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nk) return;
    const sdyn_keypoint kp = keys[(size_t)job * keyStride + i];
    const double px = (double)kp.x, py = (double)kp.y;
    uint64_t m = 0;
    for (int b = 0; b < nb; ++b) {
        const double x = sb[4 * b], y = sb[4 * b + 1], w = sb[4 * b + 2], h = sb[4 * b + 3];
        if (x <= px && px < x + w && y <= py && py < y + h) m |= 1ull << b;
    }
    mask[(size_t)job * keyStride + i] = m;
}

__device__ __forceinline__ int hamming_bytes(const uint8_t* a, const uint8_t* b)
{
    const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 16);
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 16);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

constexpr int BP = 128;

__global__ void __launch_bounds__(BP)
k_box_pairs(const BoxPairJob* __restrict__ jobs, const float* __restrict__ M, const float* __restrict__ Minv, int mode)
{
    const BoxPairJob& J = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    __shared__ int warpCnt[BP / 32];
    __shared__ int sBase;
    /* nearest train of every query, nearest query of every train; strict '<' keeps the lowest index */
    for (int i = tid; i < J.nq; i += BP) {
        int best = 1 << 30, bj = -1;
        for (int j = 0; j < J.nt; ++j) {
            const int d = hamming_bytes(J.qDesc + 32 * (size_t)i, J.tDesc + 32 * (size_t)j);
            if (d < best) { best = d; bj = j; }
        }
        J.nnQ[i] = bj; J.dQ[i] = best;
    }
    for (int j = tid; j < J.nt; j += BP) {
        int best = 1 << 30, bi = -1;
        for (int i = 0; i < J.nq; ++i) {
            const int d = hamming_bytes(J.qDesc + 32 * (size_t)i, J.tDesc + 32 * (size_t)j);
            if (d < best) { best = d; bi = i; }
        }
        J.nnT[j] = bi;
    }
    if (tid == 0) sBase = 0;
    __syncthreads();

    float m[9], mi[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { m[k] = M[k]; mi[k] = Minv[k]; }

    /* emit mutual matches in query order, classify each */
    for (int i0 = 0; i0 < J.nq; i0 += BP) {
        const int i = i0 + tid;
        bool ok = false; int tr = -1;
        if (i < J.nq && J.nt > 0) { tr = J.nnQ[i]; ok = tr >= 0 && J.nnT[tr] == i; }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0) warpCnt[tid >> 5] = __popc(bal);
        __syncthreads();
        int pos = sBase;
        for (int w = 0; w < (tid >> 5); ++w) pos += warpCnt[w];
        pos += __popc(bal & ((1u << (tid & 31)) - 1));
        if (ok) {
            const float u1 = J.tXY[2 * tr], v1 = J.tXY[2 * tr + 1];      /* reference keypoint */
            const float u2 = J.qXY[2 * i], v2 = J.qXY[2 * i + 1];        /* current keypoint */
            bool isStatic;
            if (mode == 0) {
                const float th = 5.841f;
                const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(m[0], u1), __fmul_rn(m[1], v1)), m[2]);
                const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(m[3], u1), __fmul_rn(m[4], v1)), m[5]);
                const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(m[6], u1), __fmul_rn(m[7], v1)), m[8]);
                const float num2 = __fadd_rn(__fadd_rn(__fmul_rn(a2, u2), __fmul_rn(b2, v2)), c2);
                const float d1 = __fdiv_rn(__fmul_rn(num2, num2), __fadd_rn(__fmul_rn(a2, a2), __fmul_rn(b2, b2)));
                const float a1 = __fadd_rn(__fadd_rn(__fmul_rn(m[0], u2), __fmul_rn(m[3], v2)), m[6]);
                const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(m[1], u2), __fmul_rn(m[4], v2)), m[7]);
                const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(m[2], u2), __fmul_rn(m[5], v2)), m[8]);
                const float num1 = __fadd_rn(__fadd_rn(__fmul_rn(a1, u1), __fmul_rn(b1, v1)), c1);
                const float d2 = __fdiv_rn(__fmul_rn(num1, num1), __fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1)));
                isStatic = d1 <= th && d2 <= th;      /* invSigmaSquare == 1 */
            } else {
                const float th = 5.991f;
                const float w2 = (float)(1.0 / (double)__fadd_rn(__fadd_rn(__fmul_rn(mi[6], u2), __fmul_rn(mi[7], v2)), mi[8]));
                const float u2in1 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(mi[0], u2), __fmul_rn(mi[1], v2)), mi[2]), w2);
                const float v2in1 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(mi[3], u2), __fmul_rn(mi[4], v2)), mi[5]), w2);
                const float e1x = __fsub_rn(u1, u2in1), e1y = __fsub_rn(v1, v2in1);
                const float s1 = __fadd_rn(__fmul_rn(e1x, e1x), __fmul_rn(e1y, e1y));
                const float w1 = (float)(1.0 / (double)__fadd_rn(__fadd_rn(__fmul_rn(m[6], u1), __fmul_rn(m[7], v1)), m[8]));
                const float u1in2 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], u1), __fmul_rn(m[1], v1)), m[2]), w1);
                const float v1in2 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[3], u1), __fmul_rn(m[4], v1)), m[5]), w1);
                const float e2x = __fsub_rn(u2, u1in2), e2y = __fsub_rn(v2, v1in2);
                const float s2 = __fadd_rn(__fmul_rn(e2x, e2x), __fmul_rn(e2y, e2y));
                isStatic = s2 <= th && s1 <= th;
            }
            J.outQuery[pos] = i; J.outTrain[pos] = tr; J.outDist[pos] = J.dQ[i];
            J.outFalseDyn[pos] = isStatic ? i : -1;
        }
        __syncthreads();
        if (tid == 0) { int s = 0; for (int w = 0; w < BP / 32; ++w) s += warpCnt[w]; sBase += s; }
        __syncthreads();
    }
    if (tid == 0) *J.outCount = sBase;
}

cudaError_t launch_box_mask(const sdyn_keypoint* dKeys, const int32_t* nPtr, int n, int keyStride,
                            const double* dBoxes, const int32_t* nBoxesPtr, int nboxes, int boxStride,
                            uint64_t* dMask, int njobs, cudaStream_t st, size_t boxPitch, size_t nbPitch)
{
    if (n <= 0 || njobs <= 0) return cudaSuccess;
    dim3 grid((n + 255) / 256, njobs);
    k_box_mask<<<grid, 256, 0, st>>>(dKeys, nPtr, n, keyStride, dBoxes, nBoxesPtr, nboxes, boxStride, dMask, boxPitch, nbPitch);
    return cudaGetLastError();
}

cudaError_t launch_box_pairs(const BoxPairJob* dJobs, int njobs, const float* dM, const float* dMinv, int mode, cudaStream_t st)
{
    if (njobs <= 0) return cudaSuccess;
    k_box_pairs<<<njobs, BP, 0, st>>>(dJobs, dM, dMinv, mode);
    return cudaGetLastError();
}

}  // namespace sdyn
