/* 7x7 Gaussian blur, sigma 2, of every pyramid level (K5).
 *   reference: GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on a clone of
 *              each level, src/ORBextractor.cc:1085-1086
 *   arithmetic: OpenCV's 8-bit fixed-point path, taps {18,34,48,56,48,34,18}/256, exact integer separable
 *              passes, (V + 32768) >> 16 (SURVEY A-3, pinned against cv2 4.13.0)
 * The level's own 19-px REFLECT_101 frame already holds the pixels the blur's border mode would
 * synthesise, so the kernel just reads the bordered buffer.  Only the interior is written.
 *
 * Bound: instruction issue first, HBM second (1 byte in, 1 byte out per pixel, 14 MACs).  Scalar MACs cost ~48
 * thread instructions per pixel; here both passes run on the integer dot-product unit instead:
 *   horizontal  the 7 taps of 4 adjacent outputs are 10 IDP.4A against CONSTANT coefficient words — the
 *               staged pixel words stay aligned, the tap window is shifted inside the immediates;
 *   vertical    the 16-bit row sums of two consecutive rows share a 32-bit word, so a tap pair is one
 *               IDP.2A (u16 x u8); the accumulator starts at the rounding constant and the result byte is
 *               bits 16-23 (max 65280*256 + 32768 < 2^24), extracted with PRMT.
 */
#include "sdyn_internal.h"
#include "tma.h"

namespace sdyn {

constexpr int BW = kBlurTileW, BH = kBlurTileH;
constexpr int PXB = kBlurStageW;             /* staged bytes per row: columns x0-16 .. x0+BW+15, 16-byte aligned */
constexpr int PR = kBlurStageH;              /* staged rows y0-3 .. y0+BH+2 */
static_assert(BW % 16 == 0 && BH % 4 == 0 && PR % 2 == 0, "tile shape");

/* taps k0..k6 = 18 34 48 56 48 34 18 laid over three aligned words (bytes x-4..x-1 | x..x+3 | x+4..x+7) for the
 * outputs x, x+1, x+2, x+3; byte i of a constant multiplies byte i of the pixel word */
#define CW4(b0, b1, b2, b3) ((uint32_t)(b0) | ((uint32_t)(b1) << 8) | ((uint32_t)(b2) << 16) | ((uint32_t)(b3) << 24))

__device__ __forceinline__ void hpass4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t (&o)[4])
{
    o[0] = __dp4a(w1, CW4(56, 48, 34, 18), __dp4a(w0, CW4(0, 18, 34, 48), 0u));
    o[1] = __dp4a(w2, CW4(18, 0, 0, 0), __dp4a(w1, CW4(48, 56, 48, 34), __dp4a(w0, CW4(0, 0, 18, 34), 0u)));
    o[2] = __dp4a(w2, CW4(34, 18, 0, 0), __dp4a(w1, CW4(34, 48, 56, 48), __dp4a(w0, CW4(0, 0, 0, 18), 0u)));
    o[3] = __dp4a(w2, CW4(48, 34, 18, 0), __dp4a(w1, CW4(18, 34, 48, 56), 0u));
}

__global__ void __launch_bounds__(256)
k_blur(const __grid_constant__ Geom g, const TileRef* __restrict__ tiles, const __grid_constant__ LevelMaps maps,
       uint8_t* __restrict__ blur)
{
    __shared__ __align__(128) uint8_t px[PR * PXB];
    __shared__ __align__(16) uint32_t hz2[(PR / 2) * BW];     /* [row pair][column]: row 2p in the low half, 2p+1 in the high half */
    __shared__ __align__(8) uint64_t bar;
    const TileRef t = tiles[blockIdx.x];
    const LevelGeom& L = g.L[t.level];
    const int x0 = t.tx * BW, y0 = t.ty * BH, tid = threadIdx.x;
    /* rows y0-3 .. y0+BH+2, columns x0-16 .. x0+BW+15 of the bordered level: ONE TMA box load (the x origin is a multiple
     * of 16 bytes because tiles sit on a 128-column grid and interior column 0 is 32-byte aligned); the level's frame
     * supplies the halo, parts of the box beyond the bordered extent are zero-filled and only feed unwritten outputs */
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, PR * PXB);
        tma_load_3d(px, &maps.m[t.level], kLeftPad + x0 - 16, kEdge + y0 - 3, blockIdx.y, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    /* horizontal pass: 2 rows x 4 adjacent outputs per thread, one 16-byte store */
    for (int i = tid; i < (PR / 2) * (BW / 4); i += 256) {
        const int rp = i / (BW / 4), xq = i - rp * (BW / 4);
        const uint32_t* a = reinterpret_cast<const uint32_t*>(px + (2 * rp) * PXB + 12) + xq;       /* word of columns x-4..x-1 */
        const uint32_t* b = reinterpret_cast<const uint32_t*>(px + (2 * rp + 1) * PXB + 12) + xq;
        uint32_t lo[4], hi[4];
        hpass4(a[0], a[1], a[2], lo);
        hpass4(b[0], b[1], b[2], hi);
        reinterpret_cast<uint4*>(hz2 + rp * BW)[xq] =
            make_uint4(lo[0] | (hi[0] << 16), lo[1] | (hi[1] << 16), lo[2] | (hi[2] << 16), lo[3] | (hi[3] << 16));
    }
    __syncthreads();
    /* vertical pass: 4 adjacent columns x 4 rows per thread.  Output row y = 2m (+1) reads staged rows y..y+6 =
     * pairs m..m+3; even rows use taps (k0,k1)(k2,k3)(k4,k5)(k6,0), odd rows (0,k0)(k1,k2)(k3,k4)(k5,k6). */
    uint8_t* out = blur + (size_t)blockIdx.y * g.frameBytes + L.off;
    {
        const int xq = tid & (BW / 4 - 1), rg = tid / (BW / 4);      /* rows 4*rg .. 4*rg+3 */
        const int gx = x0 + 4 * xq;
        uint4 P[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) P[k] = reinterpret_cast<const uint4*>(hz2 + (2 * rg + k) * BW)[xq];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint32_t e[4], o[4];
            const uint32_t* p0 = reinterpret_cast<const uint32_t*>(&P[h]);
            const uint32_t* p1 = reinterpret_cast<const uint32_t*>(&P[h + 1]);
            const uint32_t* p2 = reinterpret_cast<const uint32_t*>(&P[h + 2]);
            const uint32_t* p3 = reinterpret_cast<const uint32_t*>(&P[h + 3]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                e[c] = __dp2a_lo(p3[c], CW4(18, 0, 0, 0), __dp2a_lo(p2[c], CW4(48, 34, 0, 0),
                       __dp2a_lo(p1[c], CW4(48, 56, 0, 0), __dp2a_lo(p0[c], CW4(18, 34, 0, 0), 32768u))));
                o[c] = __dp2a_lo(p3[c], CW4(34, 18, 0, 0), __dp2a_lo(p2[c], CW4(56, 48, 0, 0),
                       __dp2a_lo(p1[c], CW4(34, 48, 0, 0), __dp2a_lo(p0[c], CW4(0, 18, 0, 0), 32768u))));
            }
            const int gy = y0 + 4 * rg + 2 * h;
            /* byte 2 of each accumulator; columns past the level's width land in the right frame / pad of the
             * blurred buffer, which nobody reads */
            if (gx < L.w && gy < L.h)
                *reinterpret_cast<uint32_t*>(out + (long long)gy * L.pitch + gx) =
                    __byte_perm(__byte_perm(e[0], e[1], 0x0062), __byte_perm(e[2], e[3], 0x0062), 0x5410);
            if (gx < L.w && gy + 1 < L.h)
                *reinterpret_cast<uint32_t*>(out + (long long)(gy + 1) * L.pitch + gx) =
                    __byte_perm(__byte_perm(o[0], o[1], 0x0062), __byte_perm(o[2], o[3], 0x0062), 0x5410);
        }
    }
}
#undef CW4

cudaError_t launch_blur(const Geom& g, const TileRef* tiles, int ntiles, const void* tmaMaps,
                        uint8_t* dBlur, int nframes, cudaStream_t st)
{
    dim3 grid(ntiles, nframes);
    k_blur<<<grid, 256, 0, st>>>(g, tiles, static_cast<const TmaMaps*>(tmaMaps)->blurTile, dBlur);
    return cudaGetLastError();
}

}  // namespace sdyn
