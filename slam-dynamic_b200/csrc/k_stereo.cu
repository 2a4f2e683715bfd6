/* Stereo association of a rectified pair (K9): Frame::ComputeStereoMatches, src/Frame.cc:874-1048.
 *
 *  k_stereo_rows    row table of the right keypoints (:883-902) as a CSR over image rows; one CTA per frame.
 *                   In-row order is irrelevant: the search keeps min (distance, right index), which is the
 *                   reference's "first strictly smaller" over its ascending-index row lists.
 *  k_stereo_match   one warp per left keypoint: level / disparity gates and __popc distances over the row list
 *                   (:913-958), then the 11x11 L1 sliding window over +-5 px on the keypoint's pyramid level
 *                   (:961-1000), parabola sub-pixel fit and disparity gates (:1002-1031).  The reference converts
 *                   the windows to float and sums |a - b| in double; every term is an integer below 2^24, so the
 *                   integer sum here is the same number.
 *  k_stereo_cull    median of the window distances by a two-level radix select and the 1.5*1.4*median cut
 *                   (:1034-1047); one CTA per frame.
 * Both pyramids stay in HBM where the two extractions left them — no level is copied to the host.
 */
#include "sdyn_internal.h"
#include "stereo_internal.h"

namespace sdyn {

constexpr int kThOrbDist = (SDYN_TH_HIGH + SDYN_TH_LOW) / 2;

__global__ void __launch_bounds__(256)
k_stereo_rows(StereoArgs a)
{
    extern __shared__ int sCnt[];                     /* nRows + 1 */
    __shared__ int warpSum[8];
    const int f = blockIdx.x, tid = threadIdx.x, nRows = a.nRows;
    const int n = min(a.countR[f], a.capR);
    const sdyn_keypoint* keys = a.keysR + (size_t)f * a.capR;
    int2* rk = a.rightKey + (size_t)f * a.capR;
    int32_t* rowOff = a.rowOff + (size_t)f * (a.maxRows + 1);
    int32_t* rowList = a.rowList + (size_t)f * a.listCap;
    for (int r = tid; r <= nRows; r += 256) sCnt[r] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const sdyn_keypoint kp = keys[i];
        const float r = __fmul_rn(2.0f, a.scale[kp.octave]);
        int maxr = (int)ceilf(__fadd_rn(kp.y, r)), minr = (int)floorf(__fsub_rn(kp.y, r));
        if (minr < 0 || maxr >= nRows) { a.status[f] = 1; minr = max(minr, 0); maxr = min(maxr, nRows - 1); }
        for (int y = minr; y <= maxr; ++y) atomicAdd(&sCnt[y], 1);
        rk[i] = make_int2(__float_as_int(kp.x), kp.octave | (minr << 8) | (maxr << 20));
    }
    __syncthreads();
    /* exclusive scan over the rows: a contiguous chunk per thread */
    const int per = (nRows + 255) / 256, r0 = tid * per, r1 = min(r0 + per, nRows);
    int sum = 0;
    for (int r = r0; r < r1; ++r) sum += sCnt[r];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) warpSum[tid >> 5] = incl;
    __syncthreads();
    int run = incl - sum;
    for (int w = 0; w < (tid >> 5); ++w) run += warpSum[w];
    for (int r = r0; r < r1; ++r) { const int c = sCnt[r]; rowOff[r] = run; sCnt[r] = run; run += c; }
    if (tid == 255) { int tot = 0; for (int w = 0; w < 8; ++w) tot += warpSum[w]; rowOff[nRows] = tot; }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int pk = rk[i].y, minr = (pk >> 8) & 0xfff, maxr = (pk >> 20) & 0xfff;
        for (int y = minr; y <= maxr; ++y) {
            const int pos = atomicAdd(&sCnt[y], 1);
            if (pos < a.listCap) rowList[pos] = i;
        }
    }
}

__device__ __forceinline__ int hamming256s(const uint32_t (&q)[8], const uint8_t* d)
{
    const uint4 a = *reinterpret_cast<const uint4*>(d), b = *reinterpret_cast<const uint4*>(d + 16);
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
           __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

constexpr int SMW = 8;   /* warps (left keypoints) per CTA */

__global__ void __launch_bounds__(SMW * 32)
k_stereo_match(const __grid_constant__ Geom g, StereoArgs a)
{
    __shared__ int sSad[SMW][12];
    const int f = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * SMW + warp;
    const int n = min(a.countL[f], a.capL);
    if (iL >= n) return;
    const sdyn_keypoint kp = a.keysL[(size_t)f * a.capL + iL];
    const int levelL = kp.octave;
    const float uL = kp.x, vL = kp.y;
    float outU = -1.0f, outZ = -1.0f;
    int outSad = -1;

    const int row = (int)vL;
    const float maxD = __fdiv_rn(a.mbf, a.mb);                          /* minZ = mb, minD = 0 (:905-907) */
    const float minU = __fsub_rn(uL, maxD), maxU = uL;
    constexpr uint32_t NONE = 0xffffffffu;
    uint32_t key = NONE;
    const int2* rk = a.rightKey + (size_t)f * a.capR;
    if (row >= 0 && row < a.nRows && !(maxU < 0)) {
        const int32_t* rowOff = a.rowOff + (size_t)f * (a.maxRows + 1);
        const int32_t* rowList = a.rowList + (size_t)f * a.listCap;
        const int b = rowOff[row], e = min(rowOff[row + 1], a.listCap);
        uint32_t qd[8];
        {
            const uint32_t* w = reinterpret_cast<const uint32_t*>(a.descL + ((size_t)f * a.capL + iL) * 32);
#pragma unroll
            for (int k = 0; k < 8; ++k) qd[k] = w[k];
        }
        const uint8_t* descR = a.descR + (size_t)f * a.capR * 32;
        for (int p = b + lane; p < e; p += 32) {
            const int iR = rowList[p];
            const int2 r = rk[iR];
            const int oct = r.y & 0xff;
            if (oct < levelL - 1 || oct > levelL + 1) continue;
            const float uR = __int_as_float(r.x);
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming256s(qd, descR + 32 * (size_t)iR);
                if (dist < SDYN_TH_HIGH) key = min(key, ((uint32_t)dist << 16) | (uint32_t)iR);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
    }

    if (key != NONE && (int)(key >> 16) < kThOrbDist) {
        /* sub-pixel match by correlation, coordinates in the pyramid level of the left keypoint (:961-1000) */
        const float uR0 = __int_as_float(rk[key & 0xffff].x);
        const float sf = a.invScale[levelL];
        const float suL = roundf(__fmul_rn(uL, sf)), svL = roundf(__fmul_rn(vL, sf)), suR0 = roundf(__fmul_rn(uR0, sf));
        const LevelGeom& L = g.L[levelL];
        const float iniu = suR0, endu = __fadd_rn(suR0, 11.0f);          /* scaleduR0 + L - w, scaleduR0 + L + w + 1 */
        if (!(iniu < 0 || endu >= (float)L.w)) {
            const int r0 = (int)svL - 5, c0 = (int)suL - 5, cr0 = (int)suR0 - 5;
            const int pitch = L.pitch;
            const uint8_t* pl = a.pyrL + (size_t)f * a.frameBytesL + L.off + (long long)r0 * pitch + c0;
            const uint8_t* pr = a.pyrR + (size_t)f * a.frameBytesR + L.off + (long long)r0 * pitch + cr0;
            if (lane < 11) sSad[warp][lane] = 0;
            __syncwarp();
            const int cL = pl[5 * pitch + 5];
            for (int p = lane; p < 121; p += 32) {
                const int ii = p / 11, dy = p - 11 * ii;                  /* incR = ii - 5, window row dy */
                const int cR = pr[5 * pitch + ii];                        /* centre of the window at incR: column cr0 + incR + 5 */
                const uint8_t* l = pl + dy * pitch;
                const uint8_t* r = pr + dy * pitch + ii - 5;
                int s = 0;
#pragma unroll
                for (int x = 0; x < 11; ++x) s += abs(((int)l[x] - cL) - ((int)r[x] - cR));
                atomicAdd(&sSad[warp][ii], s);
            }
            __syncwarp();
            int bestD = 0x7fffffff, binc = 0;
#pragma unroll
            for (int ii = 0; ii < 11; ++ii) { const int d = sSad[warp][ii]; if (d < bestD) { bestD = d; binc = ii - 5; } }
            if (binc != -5 && binc != 5) {
                const float d1 = (float)sSad[warp][binc + 4], d2 = (float)sSad[warp][binc + 5], d3 = (float)sSad[warp][binc + 6];
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(a.scale[levelL], __fadd_rn(__fadd_rn(suR0, (float)binc), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= 0 && disparity < maxD) {
                        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                        outZ = __fdiv_rn(a.mbf, disparity); outU = bestuR; outSad = bestD;
                    }
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        const size_t o = (size_t)f * a.capL + iL;
        a.uRight[o] = outU; a.depth[o] = outZ; a.sad[o] = outSad;
    }
}

__global__ void __launch_bounds__(256)
k_stereo_cull(StereoArgs a)
{
    __shared__ int hist[256];
    __shared__ int sM, sBin, sRank, sMedian;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(a.countL[f], a.capL);
    const int32_t* sad = a.sad + (size_t)f * a.capL;
    float* uR = a.uRight + (size_t)f * a.capL;
    float* dz = a.depth + (size_t)f * a.capL;
    hist[tid] = 0;
    if (tid == 0) sM = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < n; i += 256) { const int d = sad[i]; if (d >= 0) { ++mine; atomicAdd(&hist[d >> 8], 1); } }
    atomicAdd(&sM, mine);
    __syncthreads();
    const int M = sM;
    if (M == 0) { if (tid == 0) a.kept[f] = 0; return; }
    if (tid == 0) {              /* element M/2 of the ascending distances (:1035) */
        int rank = M / 2, b = 0;
        while (rank >= hist[b]) { rank -= hist[b]; ++b; }
        sBin = b; sRank = rank;
    }
    __syncthreads();
    const int bin = sBin;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) { const int d = sad[i]; if (d >= 0 && (d >> 8) == bin) atomicAdd(&hist[d & 255], 1); }
    __syncthreads();
    if (tid == 0) {
        int rank = sRank, b = 0;
        while (rank >= hist[b]) { rank -= hist[b]; ++b; }
        sMedian = (bin << 8) | b;
        sM = 0;
    }
    __syncthreads();
    constexpr float k = 1.5f * 1.4f;
    const float thDist = __fmul_rn(k, (float)sMedian);
    int kept = 0;
    for (int i = tid; i < n; i += 256) {
        const int d = sad[i];
        if (d < 0) continue;
        if ((float)d < thDist) ++kept;
        else { uR[i] = -1.0f; dz[i] = -1.0f; }
    }
    atomicAdd(&sM, kept);
    __syncthreads();
    if (tid == 0) a.kept[f] = sM;
}

cudaError_t launch_stereo(const Geom& g, const StereoArgs& a, int nframes, int maxKpL, cudaStream_t st)
{
    const size_t smem = (size_t)(a.nRows + 1) * sizeof(int);
    if (smem + 2048 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_stereo_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_stereo_rows<<<nframes, 256, smem, st>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid((maxKpL + SMW - 1) / SMW, nframes);
    k_stereo_match<<<grid, SMW * 32, 0, st>>>(g, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_stereo_cull<<<nframes, 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace sdyn
