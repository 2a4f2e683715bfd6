/* 7x7 Gaussian blur, sigma 2, of every pyramid level (K5).
 *   reference: GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on a clone of
 *              each level, src/ORBextractor.cc:1085-1086
 *   arithmetic: OpenCV's 8-bit fixed-point path, taps {18,34,48,56,48,34,18}/256, exact integer separable
 *              passes, (V + 32768) >> 16 (SURVEY A-3, pinned against cv2 4.13.0)
 * The level's own 19-px REFLECT_101 frame already holds the pixels the blur's border mode would
 * synthesise, so the kernel just reads the bordered buffer.  Only the interior is written.
 * Bound: HBM/L2 bandwidth (1 byte in, 1 byte out per pixel, 14 MACs).
 */
#include "sdyn_internal.h"

namespace sdyn {

constexpr int BW = kBlurTileW, BH = kBlurTileH;
constexpr int PXB = BW + 32;                 /* staged bytes per row: columns x0-16 .. x0+BW+15, 16-byte aligned */
static_assert(BW % 16 == 0 && BH % 2 == 0, "tile shape");

__global__ void __launch_bounds__(256)
k_blur(const __grid_constant__ Geom g, const TileRef* __restrict__ tiles, const uint8_t* __restrict__ pyr,
       uint8_t* __restrict__ blur)
{
    __shared__ __align__(16) uint8_t px[(BH + 6) * PXB];
    __shared__ __align__(8) uint16_t hz[(BH + 6) * BW];
    const TileRef t = tiles[blockIdx.x];
    const LevelGeom& L = g.L[t.level];
    const int x0 = t.tx * BW, y0 = t.ty * BH, tid = threadIdx.x;
    const uint8_t* img = pyr + (size_t)blockIdx.y * g.frameBytes + L.off;
    /* rows y0-3 .. y0+BH+2 (clamped to the bordered extent), 16-byte vector loads; the frame supplies the halo */
    for (int i = tid; i < (BH + 6) * (PXB / 16); i += 256) {
        const int yy = i / (PXB / 16), q = i - yy * (PXB / 16);
        const int gy = min(y0 - 3 + yy, L.h + kEdge - 1);
        reinterpret_cast<uint4*>(px + yy * PXB)[q] =
            __ldg(reinterpret_cast<const uint4*>(img + (long long)gy * L.pitch + x0 - 16) + q);
    }
    __syncthreads();
    /* horizontal pass: 4 adjacent outputs per thread from 10 staged bytes */
    for (int i = tid; i < (BH + 6) * (BW / 4); i += 256) {
        const int yy = i / (BW / 4), xq = i - yy * (BW / 4);
        const uint8_t* p = &px[yy * PXB + 13 + 4 * xq];           /* column x0 + 4*xq - 3 */
        int v[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) v[k] = p[k];
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = 18 * (v[k] + v[k + 6]) + 34 * (v[k + 1] + v[k + 5]) + 48 * (v[k + 2] + v[k + 4]) + 56 * v[k + 3];
        reinterpret_cast<uint2*>(hz + yy * BW)[xq] = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
    }
    __syncthreads();
    /* vertical pass: 4 adjacent columns per thread, one aligned 32-bit store */
    uint8_t* out = blur + (size_t)blockIdx.y * g.frameBytes + L.off;
    for (int i = tid; i < BH * (BW / 4); i += 256) {
        const int yy = i / (BW / 4), xq = i - yy * (BW / 4);
        const int gx = x0 + 4 * xq, gy = y0 + yy;
        if (gx >= L.w || gy >= L.h) continue;
        int acc[4] = {0, 0, 0, 0};
        const int kk[7] = {18, 34, 48, 56, 48, 34, 18};
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const uint2 w = reinterpret_cast<const uint2*>(hz + (yy + r) * BW)[xq];
            acc[0] += kk[r] * (int)(w.x & 0xffff); acc[1] += kk[r] * (int)(w.x >> 16);
            acc[2] += kk[r] * (int)(w.y & 0xffff); acc[3] += kk[r] * (int)(w.y >> 16);
        }
        const uint32_t o = (uint32_t)((acc[0] + 32768) >> 16) | ((uint32_t)((acc[1] + 32768) >> 16) << 8) |
                           ((uint32_t)((acc[2] + 32768) >> 16) << 16) | ((uint32_t)((acc[3] + 32768) >> 16) << 24);
        /* columns past the level's width land in the right frame / pad of the blurred buffer, which nobody reads */
        *reinterpret_cast<uint32_t*>(out + (long long)gy * L.pitch + gx) = o;
    }
}

cudaError_t launch_blur(const Geom& g, const TileRef* tiles, int ntiles, const uint8_t* dPyr,
                        uint8_t* dBlur, int nframes, cudaStream_t st)
{
    dim3 grid(ntiles, nframes);
    k_blur<<<grid, 256, 0, st>>>(g, tiles, dPyr, dBlur);
    return cudaGetLastError();
}

}  // namespace sdyn
