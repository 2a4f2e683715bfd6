/* FAST-9/16 detection with per-cell semantics (K2).
 *   reference: ComputeKeyPointsOctTree's per-cell cv::FAST calls, src/ORBextractor.cc:789-829
 *   arithmetic: cv::FAST(TYPE_9_16, nonmax=true) — SURVEY A-4 (OpenCV, un-vendored dependency)
 *
 * The reference calls cv::FAST once (or twice) per 30-ish pixel cell on a window that overlaps its
 * neighbours by 6 px; the 3-px margin cv::FAST ignores makes the cells' *interiors* tile the FAST window
 * exactly.  So: one score map V(p) for the whole level (the score does not depend on the threshold), 3x3
 * strict-maximum suppression that only looks at neighbours inside the same cell interior, every maximum
 * with V >= min(iniTh, minTh) is emitted, and a per-cell flag records whether the cell has a maximum with
 * V >= iniTh (the octree stage then applies "ini, else min" per cell).  Emission order is irrelevant: the
 * selection stage breaks response ties with an explicit (cell, y, x) key.
 *
 * Bound: integer ALU, not HBM (1 byte per pixel against a ~100-op min/max network).  The kernel therefore
 * spends its effort on NOT scoring: (A) a 4-point compass test that is an exact necessary condition for a
 * 9-arc at the low threshold rejects most pixels; it runs on whole words (4 columns x 4 rows per thread, bytes
 * widened to u16x2 lanes, two positions per VIMNMX.U16x2); (B) survivors are compacted into a dense list so the
 * full 16-tap score (VIMNMX3 sliding-window network) runs with all lanes busy; (C) the 3x3 suppression only touches
 * pixels with a non-zero score.  The tile is staged by ONE TMA box load; the tile grid (sdyn_internal.h) keeps score
 * column 0 word aligned in the staged rows.
 */
#include "sdyn_internal.h"
#include "tma.h"

namespace sdyn {

constexpr int TW = kFastTileW, TH = kFastTileH;
constexpr int PH = kFastStageH;             /* pixel rows: tile + 1 (NMS halo) + 3 (ring) on each side */
constexpr int PWB = kFastStageW;            /* staged bytes per row */
constexpr int SW = 128, SH = TH + 2;        /* score positions a CTA computes: tile + NMS halo (+ 2 spare columns) */
static_assert(TW + 2 <= SW && TW % 4 == 0 && (kFastLead & 3) == 3, "tile grid keeps score column 0 word aligned");
static_assert(PWB % 16 == 0 && PWB >= 16 + SW + 4, "box covers the aligned lead-in, the positions and the last lane's right word");
constexpr int FT = 256;                     /* threads */

/* V(p) = max over the 16 contiguous 9-arcs of max(min d, min -d) - 1, d_k = I(p) - I(ring_k).
 * Both polarities ride in one register as biased unsigned 16-bit lanes: low = d + 256, high = -d + 256.
 * With K = 1 - 65536 the pair is (c*K + 0x01000100) + r*0xFFFF — ONE integer multiply-add per ring pixel (FMA
 * pipe) — and the sliding-window minima / final maximum are 40 three-input VIMNMX3.U16x2 ops (ALU pipe) for both
 * polarities at once, instead of ~100 scalar min/max. */
__device__ __forceinline__ int fast_score(const uint8_t* p)
{
    const uint32_t cK = (uint32_t)p[0] * 0xFFFF0001u + 0x01000100u;
    uint32_t d[16];
#define RING(k, off) d[k] = (uint32_t)p[off] * 0xFFFFu + cK
    RING(0, 3 * PWB);      RING(1, 3 * PWB + 1);   RING(2, 2 * PWB + 2);   RING(3, PWB + 3);
    RING(4, 3);            RING(5, -PWB + 3);      RING(6, -2 * PWB + 2);  RING(7, -3 * PWB + 1);
    RING(8, -3 * PWB);     RING(9, -3 * PWB - 1);  RING(10, -2 * PWB - 2); RING(11, -PWB - 3);
    RING(12, -3);          RING(13, PWB - 3);      RING(14, 2 * PWB - 2);  RING(15, 3 * PWB - 1);
#undef RING
    uint32_t lo3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) lo3[k] = __vimin3_u16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    uint32_t m9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m9[k] = __vimin3_u16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
    uint32_t a = __vimax3_u16x2(m9[0], m9[1], m9[2]), b = __vimax3_u16x2(m9[3], m9[4], m9[5]);
    uint32_t c = __vimax3_u16x2(m9[6], m9[7], m9[8]), e = __vimax3_u16x2(m9[9], m9[10], m9[11]);
    uint32_t f = __vimax3_u16x2(m9[12], m9[13], m9[14]);
    a = __vimax3_u16x2(a, b, c);
    e = __vimax3_u16x2(e, f, m9[15]);
    a = __vmaxu2(a, e);
    return (int)max(a & 0xffffu, a >> 16) - 257;
}

constexpr int SWP = 136;                    /* score row pitch: position xx lives at byte xx+3, so tile column 0 is word-aligned */
constexpr int ROWR = SH / (FT / 32), COLR = SW / 32;   /* score rows per warp, 32-column chunks per row */
static_assert(SW % 32 == 0 && SH % (FT / 32) == 0, "score positions tile the CTA exactly");
static_assert(ROWR == 4 && COLR == 4, "the compass test owns 4 columns x 4 rows per thread");
static_assert(SW + 3 + 1 <= SWP && SWP % 4 == 0, "score row pitch");

__global__ void __launch_bounds__(FT, 8)
k_fast(const __grid_constant__ Geom g, const TileRef* __restrict__ tiles, const __grid_constant__ LevelMaps maps,
       int iniTh, int lowTh, uint8_t* __restrict__ cellFlag, uint32_t* __restrict__ cand,
       int32_t* __restrict__ candCount)
{
    __shared__ __align__(128) uint8_t px[PH * PWB];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) uint8_t sc[SH * SWP];
    __shared__ uint16_t list[SH * SW];          /* survivors of the compass test */
    __shared__ uint16_t corners[TH * TW];       /* tile positions with V >= lowTh (input of the suppression) */
    __shared__ uint16_t colInfo[TW], rowInfo[TH];   /* cell index | hasLow << 14 | hasHigh << 15 */
    __shared__ int nList, nCorners;

    const TileRef t = tiles[blockIdx.x];
    const int f = blockIdx.y;
    const LevelGeom& L = g.L[t.level];
    const int x0 = t.tx * TW - kFastLead, y0 = t.ty * TH;     /* window-relative origin of the tile */
    const int tid = threadIdx.x, lane = tid & 31;

    /* stage rows [gy0, gy0+PH) of the bordered level, PWB bytes from the 16-byte aligned padded column at or before the
     * word left of score column 0's centre: that centre (window x0-1 = padded column 47+x0, a multiple of 4 by the tile
     * grid) lands at byte cbase = 4, 8, 12 or 16 of a staged row and the compass test reads whole words.  One thread
     * issues the TMA box load; parts of the box beyond the bordered level are zero-filled and only feed masked-out
     * positions. */
    const int gy0 = kFastBorder + y0 - 4;
    const int xc = kLeftPad + kFastBorder + x0 - 1;           /* padded column of score column 0's centre */
    const int bx0 = (xc - 4) & ~15, cbase = xc - bx0;
    if (tid == 0) {
        nList = 0; nCorners = 0;
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, PH * PWB);
        tma_load_3d(px, &maps.m[t.level], bx0, kEdge + gy0, f, &bar);
    }
    for (int i = tid; i < SH * SWP / 4; i += FT) reinterpret_cast<uint32_t*>(sc)[i] = 0;
    /* which neighbours of a tile column / row lie in the same cell interior (interiors start at 3 + j*wCell):
     * one runtime division per tile column / row instead of two per pixel */
    if (tid < TW) {
        const int w3 = x0 + tid - 3;
        const int cx = max(w3, 0) / L.wCell, lx = w3 - cx * L.wCell;
        colInfo[tid] = (uint16_t)(cx | ((lx > 0) << 14) | ((lx < L.wCell - 1) << 15));
    } else if (tid < TW + TH) {
        const int r = tid - TW, h3 = y0 + r - 3;
        const int cy = max(h3, 0) / L.hCell, ly = h3 - cy * L.hCell;
        rowInfo[r] = (uint16_t)(cy | ((ly > 0) << 14) | ((ly < L.hCell - 1) << 15));
    }
    __syncthreads();                 /* barrier initialised, scores cleared */
    mbar_wait(&bar, 0);              /* tile landed */

    /* (A) compass test at the low threshold on [x0-1, x0+TW+1) x [y0-1, y0+TH+1): a contiguous 9-arc contains one
     * pixel of every antipodal pair, so  min(max(N,S), max(E,W)) > c + t  (bright arc)  or
     * max(min(N,S), min(E,W)) < c - t  (dark arc)  is an exact necessary condition.  A thread owns 4 adjacent
     * columns of 4 rows; per row it reads five aligned words (N, S and three centre-row words), widens the bytes to
     * u16x2 lanes with PRMT and evaluates two positions per VIMNMX.U16x2 — ~11 instructions per position instead of
     * ~30 scalar.  The comparisons use bit 15 of each lane as a borrow guard: ((X | 0x8000) - M) keeps bit 15 iff
     * X >= M.  Survivors are kept as bits; ONE warp scan + one shared atomic per warp reserves list space. */
    const int xlo = max(0, 3 - (x0 - 1)), xhi = min(TW + 2, L.fw - 3 - (x0 - 1));    /* valid score columns */
    const int ylo = max(0, 3 - (y0 - 1)), yhi = min(SH, L.fh - 3 - (y0 - 1));
    const uint32_t tG = (uint32_t)lowTh * 0x10001u + 0x80008000u;
    uint32_t bits = 0;
    const int warp = tid >> 5;
    /* bit layout of a row r (hit word >> 2r): column 0 -> bit 15, 2 -> 14, 1 -> 31, 3 -> 30 */
    uint32_t colMask = 0;
    {
        const int c0 = 4 * lane;
        colMask = ((uint32_t)(c0 >= xlo && c0 < xhi) << 15) | ((uint32_t)(c0 + 2 >= xlo && c0 + 2 < xhi) << 14) |
                  ((uint32_t)(c0 + 1 >= xlo && c0 + 1 < xhi) << 31) | ((uint32_t)(c0 + 3 >= xlo && c0 + 3 < xhi) << 30);
    }
    uint32_t valid = 0;
    /* a thread's four rows are ADJACENT (pitch 40 words: banks +0, +8, +16, +24).  Rows eight apart would put all 16
     * positions of a thread — consecutive list entries, scored by adjacent lanes in (B) — in one bank */
#pragma unroll
    for (int r = 0; r < ROWR; ++r) {
        const int yy = ROWR * warp + r;
        if (yy >= ylo && yy < yhi) valid |= colMask >> (2 * r);
        const uint32_t* rc = reinterpret_cast<const uint32_t*>(px + (yy + 3) * PWB + cbase) + lane;
        const uint32_t wl = rc[-1], wc = rc[0], wr = rc[1], wn = rc[-3 * (PWB / 4)], ws = rc[3 * (PWB / 4)];
        const uint32_t we = __byte_perm(wc, wr, 0x6543), ww = __byte_perm(wl, wc, 0x4321);     /* columns +3 / -3 */
        uint32_t h[2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t sel = half ? 0x4342u : 0x4140u;                                      /* bytes (2,3) / (0,1) -> u16 lanes */
            const uint32_t C = __byte_perm(wc, 0, sel), N = __byte_perm(wn, 0, sel), S = __byte_perm(ws, 0, sel);
            const uint32_t E = __byte_perm(we, 0, sel), W = __byte_perm(ww, 0, sel);
            const uint32_t mx = __vminu2(__vmaxu2(N, S), __vmaxu2(E, W));
            const uint32_t mn = __vmaxu2(__vminu2(N, S), __vminu2(E, W));
            const uint32_t rb = (C + tG) - mx;            /* bit 15 of a lane cleared  <=>  mx > c + t */
            const uint32_t rd = (mn + tG) - C;            /* bit 15 of a lane cleared  <=>  c > mn + t */
            h[half] = ~(rb & rd) & 0x80008000u;
        }
        bits |= (h[0] | (h[1] >> 1)) >> (2 * r);
    }
    bits &= valid;
    {
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int base = 0;
        if (lane == 31 && incl) base = atomicAdd(&nList, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int k = 15 - (b & 15);                  /* 2r + (column >= 2) */
            const int yy = ROWR * warp + (k >> 1), xx = 4 * lane + 2 * (k & 1) + (b >> 4);
            list[base++] = (uint16_t)(yy * SW + xx);
        }
    }
    __syncthreads();

    /* (B) full score of the survivors, dense over the list; corners inside the tile are appended to a second
     * list so that the suppression below is dense as well */
    const int nl = nList;
    for (int e0 = 0; e0 < nl; e0 += FT) {
        const int e = e0 + tid;
        bool corner = false;
        int pos = 0;
        if (e < nl) {
            const int i = list[e];
            const int yy = i / SW, xx = i - yy * SW;
            const int v = fast_score(&px[(yy + 3) * PWB + cbase + xx]);
            if (v >= lowTh) {
                sc[yy * SWP + xx + 3] = (uint8_t)v;
                corner = yy >= 1 && yy <= TH && xx >= 1 && xx <= TW;
                pos = (yy - 1) * TW + (xx - 1);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, corner);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&nCorners, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (corner) corners[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)pos;
        }
    }
    __syncthreads();

    /* (C) cell-confined 3x3 strict maximum over the corners */
    uint32_t* out = cand + (size_t)f * g.candPerFrame + L.candOff;
    int32_t* cnt = candCount + f * SDYN_MAX_LEVELS + t.level;
    uint8_t* flags = cellFlag + (size_t)f * g.cellsPerFrame + L.cellOff;
    const int nc = nCorners;
    for (int e = tid; e < nc; e += FT) {
        const int i = corners[e];
        const int yy = i / TW, xx = i - yy * TW;
        const uint32_t ri = rowInfo[yy], ci = colInfo[xx];
        const bool hasU = ri & 0x4000, hasD = ri & 0x8000, hasL = ci & 0x4000, hasR = ci & 0x8000;
        const uint8_t* c = &sc[(yy + 1) * SWP + 4 + xx];
        const int s = c[0];
        int m = 0;
        if (hasU) { m = max(m, (int)c[-SWP]); if (hasL) m = max(m, (int)c[-SWP - 1]); if (hasR) m = max(m, (int)c[-SWP + 1]); }
        if (hasD) { m = max(m, (int)c[SWP]);  if (hasL) m = max(m, (int)c[SWP - 1]);  if (hasR) m = max(m, (int)c[SWP + 1]); }
        if (hasL) m = max(m, (int)c[-1]);
        if (hasR) m = max(m, (int)c[1]);
        if (s > m) {
            if (s >= iniTh) flags[(ri & 0x3fff) * L.nCols + (ci & 0x3fff)] = 1;
            const int slot = atomicAdd(cnt, 1);
            if (slot < L.candCap) out[slot] = (uint32_t)(x0 + xx) | ((uint32_t)(y0 + yy) << 12) | ((uint32_t)s << 24);
        }
    }
}

cudaError_t launch_fast(const Geom& g, const TileRef* tiles, int ntiles, const void* tmaMaps,
                        int iniTh, int minTh, uint8_t* dCellFlag, uint32_t* dCand, int32_t* dCandCount,
                        int nframes, cudaStream_t st)
{
    dim3 grid(ntiles, nframes);
    k_fast<<<grid, FT, 0, st>>>(g, tiles, static_cast<const TmaMaps*>(tmaMaps)->fastTile, iniTh, iniTh < minTh ? iniTh : minTh, dCellFlag, dCand, dCandCount);
    return cudaGetLastError();
}

}  // namespace sdyn
