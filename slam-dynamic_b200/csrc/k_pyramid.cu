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

__global__ void __launch_bounds__(256)
k_resize(const __grid_constant__ Geom g, int level, const uint8_t* __restrict__ tables, uint8_t* __restrict__ pyr)
{
    const LevelGeom& L = g.L[level];
    const LevelGeom& P = g.L[level - 1];
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y * blockDim.y + threadIdx.y;   /* bordered row */
    if (q * 4 >= L.pitch || row >= L.h + 2 * kEdge) return;
    const ResizeTap* xt = reinterpret_cast<const ResizeTap*>(tables + L.xtab);
    const ResizeTap ty = reinterpret_cast<const ResizeTap*>(tables + L.ytab)[row];
    uint8_t* frame = pyr + (size_t)blockIdx.z * g.frameBytes;
    const uint8_t* s0 = frame + P.off + (long long)ty.s0 * P.pitch;
    const uint8_t* s1 = frame + P.off + (long long)ty.s1 * P.pitch;
    const int b0 = ty.c0, b1 = ty.c1;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int bc = q * 4 + i - (kLeftPad - kEdge);               /* bordered column */
        bc = max(0, min(bc, L.w + 2 * kEdge - 1));
        const ResizeTap tx = xt[bc];
        const int h0 = s0[tx.s0] * tx.c0 + s0[tx.s1] * tx.c1;
        const int h1 = s1[tx.s0] * tx.c0 + s1[tx.s1] * tx.c1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v = max(0, min(v, 255));
        out |= (uint32_t)v << (8 * i);
    }
    uint8_t* dst = frame + L.off + (long long)(row - kEdge) * L.pitch - kLeftPad;
    reinterpret_cast<uint32_t*>(dst)[q] = out;
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
    dim3 block(64, 4);
    dim3 grid((L.pitch / 4 + block.x - 1) / block.x, (L.h + 2 * kEdge + block.y - 1) / block.y, nframes);
    k_resize<<<grid, block, 0, st>>>(g, level, dTables, dPyr);
    return cudaGetLastError();
}

}  // namespace sdyn
