/* Orientation (K4), rotated-BRIEF descriptors (K6) and final keypoint assembly: a warp per group of keypoints.
 *   reference: IC_Angle                 src/ORBextractor.cc:77-104   (on the UNBLURRED level)
 *              computeOrbDescriptor     src/ORBextractor.cc:107-147  (on the blurred level)
 *              operator() output loop   src/ORBextractor.cc:1072-1103 (level-major order, pt *= scale)
 *   arithmetic: cv::fastAtan2 (SURVEY A-5), cvRound = round-half-even (A-6); float rotate without FMA
 *              contraction (Appendix B-3)
 *
 * Patches arrive in shared memory by TMA box loads (tma.h), one per keypoint and stage.  Orientation: lane v+15 owns
 * disc row v (31 lanes), its pixels are three 16-byte loads, the two int32 moments are byte dot products combined with
 * warp shuffles.  Descriptor: lane i owns byte i (8 point pairs, 16 gathers from the blurred patch).  Output slot =
 * (keypoints of lower levels) + list position, which reproduces the reference's level-major concatenation.
 */
#include "sdyn_internal.h"
#include "tma.h"

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

/* umax[v], v = 0..15, one nibble each: 15 15 15 15 14 14 14 13 13 12 11 10 9 8 6 3 */
constexpr unsigned long long kDiscHalfWidth = 0x3689ABCDDEEEFFFFull;
constexpr int DW = 8;    /* warps per CTA */
constexpr int KW = 4;    /* keypoints per warp */
constexpr int SP = kPatchPitch;   /* shared patch pitch = box width: 64 needed bytes, rows 20 banks apart */
constexpr int PR_ = kDescRows;    /* patch rows / columns: rotated pattern coordinates reach +-18 (|p| <= 18.39) */
constexpr int PBUF = (PR_ * SP + 127) / 128 * 128;   /* per-warp patch buffer: 128-byte aligned (TMA destination) */
static_assert(SP % 16 == 0, "TMA box rows are 16-byte multiples");

/* A patch = rows [y0, y0+nrows) x SP bytes starting at the 16-byte aligned interior column ax of one level, brought into
 * a warp's buffer by ONE TMA box load issued by lane 0 (tma.h; ax is a multiple of 16 as the TMA unit requires).
 * ax >= -16 and the patch never leaves the bordered rows; bytes past the padded row are zero-filled and never sampled. */
__device__ __forceinline__ void patch_load(const CUtensorMap* map, int ax, int y0, int f, uint8_t* buf, uint64_t* bar,
                                           int bytes, int lane)
{
    if (lane == 0) {
        mbar_expect_tx(bar, bytes);
        tma_load_3d(buf, map, kLeftPad + ax, kEdge + y0, f, bar);
    }
}

/* A warp owns KW consecutive keypoints of one (frame, level).
 *   A. per keypoint: the unblurred 31-row patch is loaded (TMA) and reduced to (m01, m10), one disc row per lane;
 *   B. lanes 0..KW-1 evaluate fastAtan2 / cos / sin for one keypoint each — the scalar tail costs one pass per
 *      warp instead of one per keypoint;
 *   C. per keypoint: the blurred 37-row patch is loaded (TMA) and lane i builds descriptor byte i from 16 LDS gathers. */
__global__ void __launch_bounds__(DW * 32)
k_orient_describe(const __grid_constant__ Geom g, const __grid_constant__ LevelMaps orientMaps,
                  const __grid_constant__ LevelMaps descMaps,
                  const LevelKp* __restrict__ levelKp, const int32_t* __restrict__ levelCount,
                  sdyn_keypoint* __restrict__ kpOut, uint8_t* __restrict__ descOut, int32_t* __restrict__ countOut,
                  int maxKp)
{
    __shared__ __align__(16) float4 sPat[256];                 /* pair j of descriptor byte i at [j*32 + i] */
    __shared__ __align__(128) uint8_t sPatch[DW][PBUF];
    __shared__ __align__(8) uint64_t sBar[DW];
    const int f = blockIdx.z, level = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const int t = threadIdx.x;                               /* pair t = byte (t >> 3), bit (t & 7) */
        sPat[(t & 7) * 32 + (t >> 3)] = kBriefPairs[t];
    }
    __syncthreads();
    const int k0 = (blockIdx.x * DW + warp) * KW;               /* first list position of this warp */
    const int32_t* lc = levelCount + f * SDYN_MAX_LEVELS;
    /* output slot = keypoints of the lower levels + list position (level-major concatenation of operator()) */
    const int mine = lane < g.nlevels ? lc[lane] : 0;
    int before = lane < level ? mine : 0, total = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        before += __shfl_xor_sync(0xffffffffu, before, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (k0 == 0 && level == 0 && lane == 0) countOut[f] = min(total, maxKp);
    const LevelGeom& L = g.L[level];
    int nk = min(KW, __shfl_sync(0xffffffffu, mine, level) - k0);
    nk = min(nk, maxKp - (before + k0));                         /* output capacity */
    if (nk <= 0) return;

    uint8_t* buf = sPatch[warp];
    uint64_t* bar = &sBar[warp];              /* one barrier per warp; load j completes its phase j & 1 */
    uint32_t phase = 0;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    LevelKp kp = {0, 0, 0};
    if (lane < nk) kp = levelKp[(size_t)f * g.kpPerFrame + L.kpOff + k0 + lane];

    /* ---- A. intensity centroid ---------------------------------------------------------------------------- */
    int my10 = 0, my01 = 0;
    /* byte masks of this lane's disc row over the 32-byte window u = -15 .. 16: |u| <= umax[|v|] (umax[] of
     * ORBextractor.cc:454-469 for HALF_PATCH_SIZE = 15; geometry.cpp recomputes it and tests/test_abi.py checks the two
     * agree).  Lane 31 owns no row. */
    uint32_t discMask[8];
    {
        const int av = abs(lane - 15);
        const int hw = lane < 31 ? (int)((kDiscHalfWidth >> (4 * av)) & 15) : -1;
        const uint32_t bits = hw < 0 ? 0u : ((2u << (2 * hw)) - 1u) << (15 - hw);      /* bit b = pixel u = b - 15 */
#pragma unroll
        for (int k = 0; k < 8; ++k) discMask[k] = ((((bits >> (4 * k)) & 15u) * 0x00204081u) & 0x01010101u) * 0xffu;
    }
    for (int i = 0; i < nk; ++i) {
        const int kx = __shfl_sync(0xffffffffu, (int)kp.x, i), ky = __shfl_sync(0xffffffffu, (int)kp.y, i);
        const int ax = (kx - 18) & ~15, dx = kx - ax;
        __syncwarp();                         /* every lane is done with the previous patch */
        patch_load(&orientMaps.m[level], ax, ky - 15, f, buf, bar, kOrientRows * SP, lane);
        mbar_wait(bar, phase); phase ^= 1;
        /* lane r owns disc row v = r - 15: three aligned 16-byte loads cover its 31 pixels (rows are 20 banks apart, so a
         * quarter-warp's loads are conflict-free), a funnel shift brings pixel u = -15 to byte 0, and the two moments
         * are byte dot products: m10 += sum u * I (IDP.4A against constant signed weights), m01 += v * sum I */
        int m10 = 0, m01 = 0;
        {
            const int o = dx - 15;                                 /* byte of u = -15 in the staged row: 3 .. 18 */
            const uint4* rowp = reinterpret_cast<const uint4*>(buf + min(lane, kOrientRows - 1) * SP + (o & ~15));
            const uint4 q0 = rowp[0], q1 = rowp[1], q2 = rowp[2];
            const uint32_t W[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
            const int sh = 8 * (o & 3);
            uint32_t A[8];
            switch ((o & 15) >> 2) {                               /* warp-uniform */
            case 0:
#pragma unroll
                for (int k = 0; k < 8; ++k) A[k] = __funnelshift_r(W[k], W[k + 1], sh);
                break;
            case 1:
#pragma unroll
                for (int k = 0; k < 8; ++k) A[k] = __funnelshift_r(W[k + 1], W[k + 2], sh);
                break;
            case 2:
#pragma unroll
                for (int k = 0; k < 8; ++k) A[k] = __funnelshift_r(W[k + 2], W[k + 3], sh);
                break;
            default:
#pragma unroll
                for (int k = 0; k < 8; ++k) A[k] = __funnelshift_r(W[k + 3], k + 4 < 12 ? W[k + 4] : 0u, sh);
                break;
            }
            uint32_t rowsum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t a = A[k] & discMask[k];
                /* weights u = 4k-15 .. 4k-12 as signed bytes */
                const uint32_t wt = (uint32_t)((4 * k - 15) & 0xff) | ((uint32_t)((4 * k - 14) & 0xff) << 8) |
                                        ((uint32_t)((4 * k - 13) & 0xff) << 16) | ((uint32_t)((4 * k - 12) & 0xff) << 24);
                asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(m10) : "r"(a), "r"(wt));
                rowsum = __dp4a(a, 0x01010101u, rowsum);
            }
            m01 = (lane - 15) * (int)rowsum;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
        if (lane == i) { my10 = m10; my01 = m01; }
    }

    /* ---- B. angle, rotation (one keypoint per lane) ------------------------------------------------------------
     * The reference evaluates cosf/sinf of the float angle; evaluating in double and rounding once gives
     * the correctly rounded float, which is what glibc returns except for a handful of inputs. */
    float angle = 0.f, ca = 1.f, sb = 0.f;
    if (lane < nk) {
        angle = fast_atan2_deg((float)my01, (float)my10);
        const float rad = __fmul_rn(angle, __uint_as_float(0x3c8efa35u));   /* factorPI = (float)(CV_PI/180.f) */
        ca = (float)cos((double)rad); sb = (float)sin((double)rad);
    }

    /* ---- C. rotated BRIEF ------------------------------------------------------------------------------------ */
    const int outBase = before + k0;
    for (int i = 0; i < nk; ++i) {
        const int kx = __shfl_sync(0xffffffffu, (int)kp.x, i), ky = __shfl_sync(0xffffffffu, (int)kp.y, i);
        const float a = __shfl_sync(0xffffffffu, ca, i), b = __shfl_sync(0xffffffffu, sb, i);
        const int ax = (kx - 18) & ~15, dx = kx - ax;
        __syncwarp();
        patch_load(&descMaps.m[level], ax, ky - 18, f, buf, bar, PR_ * SP, lane);
        mbar_wait(bar, phase); phase ^= 1;
        const uint8_t* c = buf + 18 * SP + dx;
        unsigned val = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 q = sPat[j * 32 + lane];
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(q.x, b), __fmul_rn(q.y, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(q.x, a), __fmul_rn(q.y, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(q.z, b), __fmul_rn(q.w, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(q.z, a), __fmul_rn(q.w, b)));
            const int t0 = c[r0 * SP + c0], t1 = c[r1 * SP + c1];
            val |= (unsigned)(t0 < t1) << j;
        }
        descOut[((size_t)f * maxKp + outBase + i) * 32 + lane] = (uint8_t)val;
    }

    if (lane < nk) {
        sdyn_keypoint o;
        o.x = level ? __fmul_rn((float)kp.x, L.scale) : (float)kp.x;
        o.y = level ? __fmul_rn((float)kp.y, L.scale) : (float)kp.y;
        o.size = L.patchSize;
        o.angle = angle;
        o.response = (float)kp.score;
        o.octave = level;
        o.class_id = -1;
        kpOut[(size_t)f * maxKp + outBase + lane] = o;
    }
}

cudaError_t launch_orient_describe(const Geom& g, const void* tmaMaps,
                                   const LevelKp* dLevelKp, const int32_t* dLevelCount,
                                   sdyn_keypoint* dKp, uint8_t* dDesc, int32_t* dCount, int maxKp,
                                   int nframes, cudaStream_t st)
{
    dim3 grid((g.maxNodeCap + DW * KW - 1) / (DW * KW), g.nlevels, nframes);
    k_orient_describe<<<grid, DW * 32, 0, st>>>(g, static_cast<const TmaMaps*>(tmaMaps)->orientPatch,
                                               static_cast<const TmaMaps*>(tmaMaps)->descPatch, dLevelKp, dLevelCount, dKp, dDesc, dCount, maxKp);
    return cudaGetLastError();
}

}  // namespace sdyn
