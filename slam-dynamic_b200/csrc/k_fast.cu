/* FAST-9/16 detection with per-cell semantics (K2).
 *   reference: ComputeKeyPointsOctTree's per-cell cv::FAST calls, src/ORBextractor.cc:789-829
 *   arithmetic: cv::FAST(TYPE_9_16, nonmax=true) — SURVEY A-4 (OpenCV, un-vendored dependency)
 *
 * The reference calls cv::FAST once (or twice) per 30-ish pixel cell on a window that overlaps its
 * neighbours by 6 px; the 3-px margin cv::FAST ignores makes the cells' *interiors* tile the FAST window
 * exactly.  So: one score map V(p) for the whole level (the score does not depend on the threshold), 3x3
 * strict-maximum suppression that only looks at neighbours inside the same cell interior, every maximum
 * with V >= minTh is emitted, and a per-cell flag records whether the cell has a maximum with V >= iniTh
 * (the octree stage then applies "ini, else min" per cell).  Emission order is irrelevant: the selection
 * stage breaks response ties with an explicit (cell, y, x) key.
 *
 * Bound: integer ALU (min/max network), not HBM — ~1.4 M pixels/frame at ~100 ops each against 1 byte each.
 */
#include "sdyn_internal.h"

namespace sdyn {

constexpr int TW = kFastTileW, TH = kFastTileH;
constexpr int PW = TW + 8, PH = TH + 8;     /* pixels: tile + 1 (NMS halo) + 3 (ring) on each side */
constexpr int SW = TW + 2, SH = TH + 2;     /* scores: tile + NMS halo */

/* V(p) = max over the 16 contiguous 9-arcs of max(min d, min -d) - 1, d_k = I(p) - I(ring_k). */
__device__ __forceinline__ int fast_score(const uint8_t* p)
{
    const int c = p[0];
    int d[16];
    d[0] = c - p[3 * PW];      d[1] = c - p[3 * PW + 1];   d[2] = c - p[2 * PW + 2];   d[3] = c - p[PW + 3];
    d[4] = c - p[3];           d[5] = c - p[-PW + 3];      d[6] = c - p[-2 * PW + 2];  d[7] = c - p[-3 * PW + 1];
    d[8] = c - p[-3 * PW];     d[9] = c - p[-3 * PW - 1];  d[10] = c - p[-2 * PW - 2]; d[11] = c - p[-PW - 3];
    d[12] = c - p[-3];         d[13] = c - p[PW - 3];      d[14] = c - p[2 * PW - 2];  d[15] = c - p[3 * PW - 1];
    /* sliding min / max over windows of 9 via windows of 3: w3[k] = op(d[k], d[k+1], d[k+2]) */
    int lo3[16], hi3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        lo3[k] = min(d[k], min(d[(k + 1) & 15], d[(k + 2) & 15]));
        hi3[k] = max(d[k], max(d[(k + 1) & 15], d[(k + 2) & 15]));
    }
    int best = -256, worst = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        best = max(best, min(lo3[k], min(lo3[(k + 3) & 15], lo3[(k + 6) & 15])));
        worst = min(worst, max(hi3[k], max(hi3[(k + 3) & 15], hi3[(k + 6) & 15])));
    }
    return max(best, -worst) - 1;
}

__global__ void __launch_bounds__(256)
k_fast(const __grid_constant__ Geom g, const TileRef* __restrict__ tiles, const uint8_t* __restrict__ pyr,
       int iniTh, int minTh, uint8_t* __restrict__ cellFlag, uint32_t* __restrict__ cand,
       int32_t* __restrict__ candCount)
{
    __shared__ uint8_t px[PH * PW];
    __shared__ uint8_t sc[SH * SW];

    const TileRef t = tiles[blockIdx.x];
    const int f = blockIdx.y;
    const LevelGeom& L = g.L[t.level];
    const int x0 = t.tx * TW, y0 = t.ty * TH;                 /* window-relative origin of the tile */
    const uint8_t* img = pyr + (size_t)f * g.frameBytes + L.off + (long long)kFastBorder * L.pitch + kFastBorder;
    const int tid = threadIdx.x;

    /* stage pixels [x0-4, x0+TW+4) x [y0-4, y0+TH+4); the 19-px border keeps every read inside the level
     * buffer as long as the coordinates are clamped to the bordered extent */
    const int loX = -kFastBorder - kEdge, hiX = L.w + kEdge - 1 - kFastBorder;
    const int loY = -kFastBorder - kEdge, hiY = L.h + kEdge - 1 - kFastBorder;
    for (int i = tid; i < PH * PW; i += 256) {
        const int yy = i / PW, xx = i - yy * PW;
        const int gx = min(max(x0 - 4 + xx, loX), hiX), gy = min(max(y0 - 4 + yy, loY), hiY);
        px[i] = img[(long long)gy * L.pitch + gx];
    }
    __syncthreads();

    /* scores on [x0-1, x0+TW+1) x [y0-1, y0+TH+1); 0 outside the cv::FAST-valid range of the window */
    for (int i = tid; i < SH * SW; i += 256) {
        const int yy = i / SW, xx = i - yy * SW;
        const int wx = x0 - 1 + xx, wy = y0 - 1 + yy;
        int v = 0;
        if (wx >= 3 && wx < L.fw - 3 && wy >= 3 && wy < L.fh - 3) {
            v = fast_score(&px[(yy + 3) * PW + xx + 3]);
            v = v < minTh ? 0 : v;
        }
        sc[i] = (uint8_t)v;
    }
    __syncthreads();

    uint32_t* out = cand + (size_t)f * g.candPerFrame + L.candOff;
    int32_t* cnt = candCount + f * SDYN_MAX_LEVELS + t.level;
    uint8_t* flags = cellFlag + (size_t)f * g.cellsPerFrame + L.cellOff;
    for (int i = tid; i < TW * TH; i += 256) {
        const int yy = i / TW, xx = i - yy * TW;
        const int wx = x0 + xx, wy = y0 + yy;
        const int s = sc[(yy + 1) * SW + xx + 1];
        bool keep = s > 0 && wx < L.fw - 3 && wy < L.fh - 3;   /* s > 0 already implies wx,wy >= 3 */
        if (keep) {
            /* neighbours count only inside the same cell interior: interiors start at 3 + j*wCell */
            const int cx = (wx - 3) / L.wCell, cy = (wy - 3) / L.hCell;
            const int lx = wx - 3 - cx * L.wCell, ly = wy - 3 - cy * L.hCell;
            const bool hasL = lx > 0, hasR = lx < L.wCell - 1, hasU = ly > 0, hasD = ly < L.hCell - 1;
            const uint8_t* c = &sc[(yy + 1) * SW + xx + 1];
            int m = 0;
            if (hasU) { m = max(m, (int)c[-SW]); if (hasL) m = max(m, (int)c[-SW - 1]); if (hasR) m = max(m, (int)c[-SW + 1]); }
            if (hasD) { m = max(m, (int)c[SW]);  if (hasL) m = max(m, (int)c[SW - 1]);  if (hasR) m = max(m, (int)c[SW + 1]); }
            if (hasL) m = max(m, (int)c[-1]);
            if (hasR) m = max(m, (int)c[1]);
            keep = s > m;
            if (keep) {
                if (s >= iniTh) flags[cy * L.nCols + cx] = 1;
                const int slot = atomicAdd(cnt, 1);
                if (slot < L.candCap) out[slot] = (uint32_t)wx | ((uint32_t)wy << 12) | ((uint32_t)s << 24);
            }
        }
    }
}

cudaError_t launch_fast(const Geom& g, const TileRef* tiles, int ntiles, const uint8_t* dPyr,
                        int iniTh, int minTh, uint8_t* dCellFlag, uint32_t* dCand, int32_t* dCandCount,
                        int nframes, cudaStream_t st)
{
    dim3 grid(ntiles, nframes);
    k_fast<<<grid, 256, 0, st>>>(g, tiles, dPyr, iniTh, minTh, dCellFlag, dCand, dCandCount);
    return cudaGetLastError();
}

}  // namespace sdyn
