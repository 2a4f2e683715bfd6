/* Pyramid kernels (K1): level 0 = input + REFLECT_101 frame; level l>0 = fixed-point bilinear resize of
 * level l-1, frame produced in the same pass.
 *   reference: ORBextractor::ComputePyramid, src/ORBextractor.cc:1107-1132
 *   arithmetic: cv::resize INTER_LINEAR 8-bit path + cv::copyMakeBorder REFLECT_101 (SURVEY A-1, A-2)
 *
 * Bound: HBM/L2 bandwidth.  Algorithmic bytes per frame: read W*H once, write every bordered level once.
 * Every thread produces 4 horizontally adjacent bytes of a bordered row and stores them with one aligned
 * 32-bit store; the rows of a level are 32-byte aligned at interior column 0.
 */
#include "sdyn_internal.h"

namespace sdyn {

__device__ __forceinline__ int refl(int i, int n)
{
    /* valid for -n < i < 2n-1, which 19-pixel borders satisfy for n >= 20; loop keeps tiny levels correct */
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

__global__ void __launch_bounds__(256)
k_level0(const __grid_constant__ Geom g, const uint8_t* __restrict__ in, size_t inFrameStride, int inRowStride,
         uint8_t* __restrict__ pyr)
{
    const LevelGeom& L = g.L[0];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;     /* 4-byte group inside the padded row */
    const int row = blockIdx.y * blockDim.y + threadIdx.y;   /* bordered row */
    if (q * 4 >= L.pitch || row >= L.h + 2 * kEdge) return;
    const uint8_t* src = in + (size_t)blockIdx.z * inFrameStride + (size_t)refl(row - kEdge, L.h) * inRowStride;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int col = q * 4 + i - kLeftPad;                       /* interior coordinate */
        col = max(-kEdge, min(col, L.w + kEdge - 1));          /* padding bytes replicate the frame edge */
        out |= (uint32_t)__ldg(src + refl(col, L.w)) << (8 * i);
    }
    uint8_t* dst = pyr + (size_t)blockIdx.z * g.frameBytes + L.off + (long long)(row - kEdge) * L.pitch - kLeftPad;
    reinterpret_cast<uint32_t*>(dst)[q] = out;
}

/* Bilinear resize of one level, tile-staged.  A CTA produces RT_W x RT_H bytes of the bordered (and left-padded)
 * destination; the source rectangle its taps touch — found with a block min/max over the tile's column and row
 * taps, so mirrored border tiles need no special case — is staged in shared memory with aligned 16-byte loads,
 * and every thread then produces 4 adjacent bytes of two rows from shared memory. */
constexpr int RT_W = 128, RT_H = 16;

__global__ void __launch_bounds__(256)
k_resize(const __grid_constant__ Geom g, int level, const uint8_t* __restrict__ tables, uint8_t* __restrict__ pyr)
{
    extern __shared__ __align__(16) uint8_t src[];
    __shared__ ResizeTap sx[RT_W], sy[RT_H];
    __shared__ int sMinC, sMaxC, sMinR, sMaxR;
    const LevelGeom& L = g.L[level];
    const LevelGeom& P = g.L[level - 1];
    const int tid = threadIdx.x;
    const int pc0 = blockIdx.x * RT_W, row0 = blockIdx.y * RT_H;     /* padded column / bordered row of the tile */
    const int bw = L.w + 2 * kEdge, bh = L.h + 2 * kEdge;
    const int srcPitch = L.rsPitch;
    if (tid == 0) { sMinC = 1 << 30; sMaxC = -1; sMinR = 1 << 30; sMaxR = -1; }
    __syncthreads();
    if (tid < RT_W) {
        const int bc = max(0, min(pc0 + tid - (kLeftPad - kEdge), bw - 1));   /* padding bytes replicate the frame edge */
        const ResizeTap t = reinterpret_cast<const ResizeTap*>(tables + L.xtab)[bc];
        sx[tid] = t;
        int lo = min((int)t.s0, (int)t.s1), hi = max((int)t.s0, (int)t.s1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if ((tid & 31) == 0) { atomicMin(&sMinC, lo); atomicMax(&sMaxC, hi); }
    } else if (tid < RT_W + RT_H) {
        const int r = min(row0 + tid - RT_W, bh - 1);
        const ResizeTap t = reinterpret_cast<const ResizeTap*>(tables + L.ytab)[r];
        sy[tid - RT_W] = t;
        atomicMin(&sMinR, min((int)t.s0, (int)t.s1)); atomicMax(&sMaxR, max((int)t.s0, (int)t.s1));
    }
    __syncthreads();
    uint8_t* frame = pyr + (size_t)blockIdx.z * g.frameBytes;
    const int ax0 = sMinC & ~15, minR = sMinR;
    const int n16 = (sMaxC - ax0) / 16 + 1, nrows = sMaxR - minR + 1;
    const uint8_t* sbase = frame + P.off + (long long)minR * P.pitch + ax0;
    for (int i = tid; i < nrows * n16; i += 256) {
        const int r = i / n16, q = i - r * n16;
        reinterpret_cast<uint4*>(src + r * srcPitch)[q] = __ldg(reinterpret_cast<const uint4*>(sbase + (long long)r * P.pitch) + q);
    }
    __syncthreads();
    const int gx = tid & 31;
    if (pc0 + gx * 4 >= L.pitch) return;
    ResizeTap tx[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { tx[i] = sx[gx * 4 + i]; tx[i].s0 -= ax0; tx[i].s1 -= ax0; }
#pragma unroll
    for (int rr = tid >> 5; rr < RT_H; rr += 8) {
        const int row = row0 + rr;
        if (row >= bh) break;
        const ResizeTap ty = sy[rr];
        const uint8_t* s0 = src + (ty.s0 - minR) * srcPitch;
        const uint8_t* s1 = src + (ty.s1 - minR) * srcPitch;
        const int b0 = ty.c0, b1 = ty.c1;
        uint32_t out = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int h0 = s0[tx[i].s0] * tx[i].c0 + s0[tx[i].s1] * tx[i].c1;
            const int h1 = s1[tx[i].s0] * tx[i].c0 + s1[tx[i].s1] * tx[i].c1;
            int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            v = max(0, min(v, 255));
            out |= (uint32_t)v << (8 * i);
        }
        uint8_t* dst = frame + L.off + (long long)(row - kEdge) * L.pitch - kLeftPad + pc0;
        reinterpret_cast<uint32_t*>(dst)[gx] = out;
    }
}

cudaError_t launch_level0(const Geom& g, const uint8_t* dIn, size_t inFrameStride, int inRowStride,
                          uint8_t* dPyr, int nframes, cudaStream_t st)
{
    const LevelGeom& L = g.L[0];
    dim3 block(64, 4);
    dim3 grid((L.pitch / 4 + block.x - 1) / block.x, (L.h + 2 * kEdge + block.y - 1) / block.y, nframes);
    k_level0<<<grid, block, 0, st>>>(g, dIn, inFrameStride, inRowStride, dPyr);
    return cudaGetLastError();
}

cudaError_t launch_resize(const Geom& g, int level, const uint8_t* dTables, uint8_t* dPyr, int nframes, cudaStream_t st)
{
    const LevelGeom& L = g.L[level];
    const size_t smem = (size_t)L.rsPitch * L.rsRows;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_resize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid((L.pitch + RT_W - 1) / RT_W, (L.h + 2 * kEdge + RT_H - 1) / RT_H, nframes);
    k_resize<<<grid, 256, smem, st>>>(g, level, dTables, dPyr);
    return cudaGetLastError();
}

}  // namespace sdyn
