/* Orientation (K4), rotated-BRIEF descriptors (K6) and final keypoint assembly: one warp per keypoint.
 *   reference: IC_Angle                 src/ORBextractor.cc:77-104   (on the UNBLURRED level)
 *              computeOrbDescriptor     src/ORBextractor.cc:107-147  (on the blurred level)
 *              operator() output loop   src/ORBextractor.cc:1072-1103 (level-major order, pt *= scale)
 *   arithmetic: cv::fastAtan2 (SURVEY A-5), cvRound = round-half-even (A-6); float rotate without FMA
 *              contraction (Appendix B-3)
 *
 * Orientation: lane u+15 owns disc column u (31 lanes); rows are walked with coalesced 31-byte reads and
 * the two int32 moments are combined with warp shuffles.  Descriptor: lane i owns byte i (8 point pairs,
 * 16 gathers from the blurred level).  Output slot = (keypoints of lower levels) + list position, which
 * reproduces the reference's level-major concatenation.
 */
#include "sdyn_internal.h"

namespace sdyn {

/* the 256 point pairs as floats (x0, y0, x1, y1): the rotation is float arithmetic, so no per-sample I2F */
__device__ const float4 kBriefPairs[256] = {
#include "../../include/sdyn_brief_pattern.inc"
};

__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float p1 = __uint_as_float(0x4265226fu), p3 = __uint_as_float(0xc19556eeu);
    const float p5 = __uint_as_float(0x410e9fbfu), p7 = __uint_as_float(0xc0228ad9u);
    const float eps = 2.220446049250313e-16f;   /* (float)DBL_EPSILON */
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int DW = 8;   /* warps per CTA */

__global__ void __launch_bounds__(DW * 32)
k_orient_describe(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                  const LevelKp* __restrict__ levelKp, const int32_t* __restrict__ levelCount,
                  sdyn_keypoint* __restrict__ kpOut, uint8_t* __restrict__ descOut, int32_t* __restrict__ countOut,
                  int maxKp)
{
    const int f = blockIdx.z, level = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * DW + (threadIdx.x >> 5);        /* list position inside the level */
    const int32_t* lc = levelCount + f * SDYN_MAX_LEVELS;
    /* output slot = keypoints of the lower levels + list position (level-major concatenation of operator()) */
    const int mine = lane < g.nlevels ? lc[lane] : 0;
    int before = lane < level ? mine : 0, total = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (k == 0 && level == 0 && lane == 0) countOut[f] = min(total, maxKp);
    const LevelGeom& L = g.L[level];
    if (k >= __shfl_sync(0xffffffffu, mine, level)) return;
    const int outIdx = before + k;
    if (outIdx >= maxKp) return;
    const int slot = L.kpOff + k;

    const LevelKp kp = levelKp[(size_t)f * g.kpPerFrame + slot];
    const int pitch = L.pitch;
    const uint8_t* img = pyr + (size_t)f * g.frameBytes + L.off + (long long)kp.y * pitch + kp.x;
    const uint8_t* bl = blur + (size_t)f * g.frameBytes + L.off + (long long)kp.y * pitch + kp.x;

    /* ---- intensity centroid --------------------------------------------------------------------------- */
    const int u = lane - 15;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int au = abs(u);
        /* half-widths as compile-time constants: the unrolled loop folds them into immediates, and the 31 row
         * loads are independent, so they are all in flight together */
        /* umax[] of ORBextractor.cc:454-469 for HALF_PATCH_SIZE = 15 (geometry.cpp recomputes it; tests/test_abi.py
         * checks the two agree) */
        constexpr int HW[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        int vals[31];
#pragma unroll
        for (int v = -15; v <= 15; ++v)
            vals[v + 15] = au <= HW[v < 0 ? -v : v] ? (int)img[v * pitch + u] : 0;
        int rowsum = 0;
#pragma unroll
        for (int v = -15; v <= 15; ++v) { rowsum += vals[v + 15]; m01 += v * vals[v + 15]; }
        m10 = u * rowsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    /* ---- rotated BRIEF ------------------------------------------------------------------------------------
     * The reference evaluates cosf/sinf of the float angle; evaluating in double and rounding once gives
     * the correctly rounded float, which is what glibc returns except for a handful of inputs. */
    const float rad = __fmul_rn(angle, __uint_as_float(0x3c8efa35u));   /* factorPI = (float)(CV_PI/180.f) */
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);
    unsigned val = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 q = kBriefPairs[lane * 8 + j];
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(q.x, b), __fmul_rn(q.y, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(q.x, a), __fmul_rn(q.y, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(q.z, b), __fmul_rn(q.w, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(q.z, a), __fmul_rn(q.w, b)));
        const int t0 = bl[r0 * pitch + c0], t1 = bl[r1 * pitch + c1];
        val |= (unsigned)(t0 < t1) << j;
    }
    descOut[((size_t)f * maxKp + outIdx) * 32 + lane] = (uint8_t)val;

    if (lane == 0) {
        sdyn_keypoint o;
        o.x = level ? __fmul_rn((float)kp.x, L.scale) : (float)kp.x;
        o.y = level ? __fmul_rn((float)kp.y, L.scale) : (float)kp.y;
        o.size = L.patchSize;
        o.angle = angle;
        o.response = (float)kp.score;
        o.octave = level;
        o.class_id = -1;
        kpOut[(size_t)f * maxKp + outIdx] = o;
    }
}

cudaError_t launch_orient_describe(const Geom& g, const uint8_t* dPyr, const uint8_t* dBlur,
                                   const LevelKp* dLevelKp, const int32_t* dLevelCount,
                                   sdyn_keypoint* dKp, uint8_t* dDesc, int32_t* dCount, int maxKp,
                                   int nframes, cudaStream_t st)
{
    dim3 grid((g.maxNodeCap + DW - 1) / DW, g.nlevels, nframes);
    k_orient_describe<<<grid, DW * 32, 0, st>>>(g, dPyr, dBlur, dLevelKp, dLevelCount, dKp, dDesc, dCount, maxKp);
    return cudaGetLastError();
}

}  // namespace sdyn
