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

__global__ void __launch_bounds__(256)
k_blur(const __grid_constant__ Geom g, const TileRef* __restrict__ tiles, const uint8_t* __restrict__ pyr,
       uint8_t* __restrict__ blur)
{
    __shared__ uint8_t px[(BH + 6) * (BW + 8)];
    __shared__ uint16_t hz[(BH + 6) * BW];
    const TileRef t = tiles[blockIdx.x];
    const LevelGeom& L = g.L[t.level];
    const int x0 = t.tx * BW, y0 = t.ty * BH, tid = threadIdx.x;
    const uint8_t* img = pyr + (size_t)blockIdx.y * g.frameBytes + L.off;
    const int hiX = L.w + kEdge - 1, hiY = L.h + kEdge - 1;
    for (int i = tid; i < (BH + 6) * (BW + 6); i += 256) {
        const int yy = i / (BW + 6), xx = i - yy * (BW + 6);
        const int gx = min(x0 - 3 + xx, hiX), gy = min(y0 - 3 + yy, hiY);
        px[yy * (BW + 8) + xx] = img[(long long)gy * L.pitch + gx];
    }
    __syncthreads();
    for (int i = tid; i < (BH + 6) * BW; i += 256) {
        const int yy = i / BW, xx = i - yy * BW;
        const uint8_t* p = &px[yy * (BW + 8) + xx];
        hz[i] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    uint8_t* out = blur + (size_t)blockIdx.y * g.frameBytes + L.off;
    for (int i = tid; i < BH * BW; i += 256) {
        const int yy = i / BW, xx = i - yy * BW;
        const int gx = x0 + xx, gy = y0 + yy;
        if (gx >= L.w || gy >= L.h) continue;
        const uint16_t* p = &hz[yy * BW + xx];
        const int v = 18 * (p[0] + p[6 * BW]) + 34 * (p[BW] + p[5 * BW]) + 48 * (p[2 * BW] + p[4 * BW]) + 56 * p[3 * BW];
        out[(long long)gy * L.pitch + gx] = (uint8_t)((v + 32768) >> 16);
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
