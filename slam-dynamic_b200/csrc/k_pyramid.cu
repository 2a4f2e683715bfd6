/* Pyramid kernels (K1): level 0 = input + REFLECT_101 frame; level l>0 = fixed-point bilinear resize of
 * level l-1, frame produced in the same pass.
 *   reference: ORBextractor::ComputePyramid, src/ORBextractor.cc:1107-1132
 *   arithmetic: cv::resize INTER_LINEAR 8-bit path + cv::copyMakeBorder REFLECT_101 (SURVEY A-1, A-2)
 *
 * Bound: instruction issue first (60-75 % issue-active at 15-25 % of DRAM throughput), HBM second.  Algorithmic bytes
 * per frame: read W*H once, write every bordered level once.
 * Every thread produces 4 horizontally adjacent bytes of a bordered row and stores them with one aligned
 * 32-bit store; the rows of a level are 32-byte aligned at interior column 0.
 */
#include "sdyn_internal.h"
#include "tma.h"

namespace sdyn {

__device__ __forceinline__ int refl(int i, int n)
{
    /* valid for -n < i < 2n-1, which 19-pixel borders satisfy for n >= 20; loop keeps tiny levels correct */
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

/* Level 0: one thread per 16-byte chunk of a padded destination row.  Chunks that lie inside the image are
 * realigned copies: the source row starts at an arbitrary byte address (row stride = W for packed frames), so the
 * chunk is read as aligned 32-bit words and shifted into place with funnel shifts, then stored with one 16-byte
 * store.  Only the few chunks that touch the REFLECT_101 frame or the padding take the per-byte path. */
__global__ void __launch_bounds__(256)
k_level0(const __grid_constant__ Geom g, const uint8_t* __restrict__ in, size_t inFrameStride, int inRowStride,
         uint8_t* __restrict__ pyr, int nframes)
{
    const LevelGeom& L = g.L[0];
    const int nchunks = L.pitch / 16;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int row = idx / nchunks, c = idx - row * nchunks;   /* bordered row, chunk inside the padded row */
    if (row >= L.h + 2 * kEdge) return;
    const uint8_t* src = in + (size_t)blockIdx.z * inFrameStride + (size_t)refl(row - kEdge, L.h) * inRowStride;
    const int col0 = c * 16 - kLeftPad;                        /* interior coordinate of the chunk's first byte */
    uint4 out;
    bool fast = col0 >= 0 && col0 + 16 <= L.w;
    if (fast) {
        const uintptr_t s = reinterpret_cast<uintptr_t>(src + col0);
        const unsigned a = (unsigned)(s & 3);
        const uintptr_t lo = s - a, hi = lo + (a ? 20 : 16);    /* aligned words actually read */
        const uintptr_t bufLo = reinterpret_cast<uintptr_t>(in);
        const uintptr_t bufHi = bufLo + (size_t)(nframes - 1) * inFrameStride + (size_t)(L.h - 1) * inRowStride + L.w;
        fast = lo >= bufLo && hi <= bufHi;
        if (fast) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(lo);
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2), w3 = __ldg(p + 3);
            const uint32_t w4 = a ? __ldg(p + 4) : 0u;
            const unsigned sh = 8 * a;
            out = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh),
                             __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
        }
    }
    if (!fast) {
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int col = col0 + i;
            col = max(-kEdge, min(col, L.w + kEdge - 1));      /* padding bytes replicate the frame edge */
            o[i >> 2] |= (uint32_t)__ldg(src + refl(col, L.w)) << (8 * (i & 3));
        }
        out = make_uint4(o[0], o[1], o[2], o[3]);
    }
    uint8_t* dst = pyr + (size_t)blockIdx.z * g.frameBytes + L.off + (long long)(row - kEdge) * L.pitch - kLeftPad;
    reinterpret_cast<uint4*>(dst)[c] = out;
}

/* Bilinear resize of one level, tile-staged and separable.  A CTA produces RT_W x RT_H bytes of the bordered (and
 * left-padded) destination:
 *   1. the source rectangle its taps touch (mirrored border tiles included: the rectangle comes from the taps, tabulated
 *      per tile column / row on the host) is staged in shared memory by one TMA box load;
 *   2. horizontal pass ONCE per source row: H = (S[s0]*c0 + S[s1]*c1) >> 4 as uint16 (fits: 255*2048 >> 4);
 *   3. vertical pass per output: ((b0*H0) >> 16) + ((b1*H1) >> 16) + 2 >> 2, four bytes per 32-bit store.
 * Consecutive output rows share source rows (scale 1.2), so the horizontal work is 1.2 rows per output row
 * instead of 2. */
constexpr int RT_W = 128, RT_H = 32;

/* a*c0 + b*c1 as two IMADs (FMA pipe): the compiler's own choice, PRMT + PRMT + IDP.2A, loads the busier ALU pipe */
__device__ __forceinline__ uint32_t mad2(uint32_t a, uint32_t c0, uint32_t b, uint32_t c1)
{
    uint32_t r;
    asm("{ .reg .u32 t; mul.lo.u32 t, %3, %4; mad.lo.u32 %0, %1, %2, t; }" : "=r"(r) : "r"(a), "r"(c0), "r"(b), "r"(c1));
    return r;
}

/* SPITCH > 0: the staging pitch as a compile-time constant (row offsets become immediates of the shared loads);
 * SPITCH == 0: run-time pitch L.rsPitch (scale factors whose source window is wider than 192 bytes). */
template <int SPITCH>
__global__ void __launch_bounds__(256)
k_resize(const __grid_constant__ Geom g, int level, const uint8_t* __restrict__ tables, const __grid_constant__ LevelMaps maps,
         uint8_t* __restrict__ pyr)
{
    extern __shared__ __align__(128) uint8_t smemR[];
    __shared__ ResizeTap sx[RT_W], sy[RT_H];
    __shared__ int sMinC, sMaxC, sMinR, sMaxR;
    __shared__ __align__(8) uint64_t bar;
    const LevelGeom& L = g.L[level];
    const LevelGeom& P = g.L[level - 1];
    const int tid = threadIdx.x;
    const int pc0 = blockIdx.x * RT_W, row0 = blockIdx.y * RT_H;     /* padded column / bordered row of the tile */
    const int bw = L.w + 2 * kEdge, bh = L.h + 2 * kEdge;
    const int srcPitch = SPITCH > 0 ? SPITCH : L.rsPitch;
    uint8_t* src = smemR;                                             /* rsRows x rsPitch bytes */
    uint16_t* hz = reinterpret_cast<uint16_t*>(smemR + (size_t)L.rsRows * srcPitch);   /* rsRows x RT_W */
    uint8_t* frame = pyr + (size_t)blockIdx.z * g.frameBytes;
    int ax0, minR, nrows;
    if (SPITCH > 0) {
        /* common case: the source rectangle is a fixed SPITCH x rsRows box whose origin the host tabulated per tile column
         * and tile row (xTile / yTile) — ONE TMA load, issued before the taps are fetched so the two overlap */
        ax0 = reinterpret_cast<const int16_t*>(tables + L.xTile)[blockIdx.x];
        const int16_t* yt = reinterpret_cast<const int16_t*>(tables + L.yTile) + 2 * blockIdx.y;
        minR = yt[0]; nrows = yt[1];
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_expect_tx(&bar, (uint32_t)(SPITCH * L.rsRows));
            tma_load_3d(src, &maps.m[level], kLeftPad + ax0, kEdge + minR, blockIdx.z, &bar);
        }
        if (tid < RT_W) {
            const int bc = max(0, min(pc0 + tid - (kLeftPad - kEdge), bw - 1));   /* padding bytes replicate the frame edge */
            sx[tid] = reinterpret_cast<const ResizeTap*>(tables + L.xtab)[bc];
        } else if (tid < RT_W + RT_H) {
            sy[tid - RT_W] = reinterpret_cast<const ResizeTap*>(tables + L.ytab)[min(row0 + tid - RT_W, bh - 1)];
        }
        __syncthreads();
        mbar_wait(&bar, 0);
    } else {
        /* wide source windows (other scale factors): rectangle from a block min/max over the tile's taps, staged with
         * aligned 16-byte loads */
        if (tid == 0) { sMinC = 1 << 30; sMaxC = -1; sMinR = 1 << 30; sMaxR = -1; }
        __syncthreads();
        if (tid < RT_W) {
            const int bc = max(0, min(pc0 + tid - (kLeftPad - kEdge), bw - 1));
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
        ax0 = sMinC & ~15; minR = sMinR;
        const int n16 = (sMaxC - ax0) / 16 + 1;
        nrows = sMaxR - minR + 1;
        const uint8_t* sbase = frame + P.off + (long long)minR * P.pitch + ax0;
        for (int q = tid & 15; q < n16; q += 16)
            for (int r = tid >> 4; r < nrows; r += 16)
                reinterpret_cast<uint4*>(src + r * srcPitch)[q] = __ldg(reinterpret_cast<const uint4*>(sbase + (long long)r * P.pitch) + q);
        __syncthreads();
    }
    /* horizontal pass: thread = (column c of the tile, source rows r = tid/128, +2, ...), four rows per iteration so the
     * loop and address arithmetic are paid once per four values */
    {
        const int c = tid & (RT_W - 1);
        const ResizeTap t = sx[c];
        const uint32_t c0 = (uint32_t)t.c0, c1 = (uint32_t)t.c1;
        const int r0 = tid >> 7;
        const uint8_t* s0 = src + (t.s0 - ax0) + r0 * srcPitch, * s1 = src + (t.s1 - ax0) + r0 * srcPitch;
        uint16_t* h = hz + c + r0 * RT_W;
        const int p2 = 2 * srcPitch, p4 = 4 * srcPitch, p6 = 6 * srcPitch;
        int r = r0;
        for (; r + 6 < nrows; r += 8) {
            const uint32_t a0 = s0[0], a1 = s0[p2], a2 = s0[p4], a3 = s0[p6];
            const uint32_t b0 = s1[0], b1 = s1[p2], b2 = s1[p4], b3 = s1[p6];
            h[0] = (uint16_t)(mad2(a0, c0, b0, c1) >> 4);
            h[2 * RT_W] = (uint16_t)(mad2(a1, c0, b1, c1) >> 4);
            h[4 * RT_W] = (uint16_t)(mad2(a2, c0, b2, c1) >> 4);
            h[6 * RT_W] = (uint16_t)(mad2(a3, c0, b3, c1) >> 4);
            s0 += 8 * srcPitch; s1 += 8 * srcPitch; h += 8 * RT_W;
        }
        for (; r < nrows; r += 2) {
            h[0] = (uint16_t)((s0[0] * c0 + s1[0] * c1) >> 4);
            s0 += p2; s1 += p2; h += 2 * RT_W;
        }
    }
    __syncthreads();
    /* vertical pass: thread = 4 adjacent columns of output rows tid/32 + {0, 8, 16, 24}.  (b*H) >> 16 is the high half
     * of H * (b << 16): one IMAD.HI per product instead of a multiply and a shift. */
    const int gx = tid & 31;
    if (pc0 + gx * 4 >= L.pitch) return;
    {
        const int rr0 = tid >> 5;
        uint8_t* dst = frame + L.off + (long long)(row0 + rr0 - kEdge) * L.pitch - kLeftPad + pc0 + gx * 4;
        const long long dstep = 8LL * L.pitch;
        const uint16_t* hcol = hz + gx * 4;
#pragma unroll
        for (int k = 0; k < RT_H / 8; ++k) {
            const int rr = rr0 + 8 * k;
            if (row0 + rr >= bh) break;
            const ResizeTap ty = sy[rr];
            const uint2 h0 = *reinterpret_cast<const uint2*>(hcol + (ty.s0 - minR) * RT_W);
            const uint2 h1 = *reinterpret_cast<const uint2*>(hcol + (ty.s1 - minR) * RT_W);
            const uint32_t b0 = (uint32_t)ty.c0 << 16, b1 = (uint32_t)ty.c1 << 16;
            const uint32_t v0 = (__umulhi(h0.x & 0xffffu, b0) + __umulhi(h1.x & 0xffffu, b1) + 2) >> 2;
            const uint32_t v1 = (__umulhi(h0.x >> 16, b0) + __umulhi(h1.x >> 16, b1) + 2) >> 2;
            const uint32_t v2 = (__umulhi(h0.y & 0xffffu, b0) + __umulhi(h1.y & 0xffffu, b1) + 2) >> 2;
            const uint32_t v3 = (__umulhi(h0.y >> 16, b0) + __umulhi(h1.y >> 16, b1) + 2) >> 2;
            /* coefficients are non-negative and sum to 2048, so 0 <= v <= 255 without clamping */
            *reinterpret_cast<uint32_t*>(dst + k * dstep) =
                __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410);
        }
    }
}

cudaError_t launch_level0(const Geom& g, const uint8_t* dIn, size_t inFrameStride, int inRowStride,
                          uint8_t* dPyr, int nframes, cudaStream_t st)
{
    const LevelGeom& L = g.L[0];
    const int items = (L.pitch / 16) * (L.h + 2 * kEdge);
    dim3 grid((items + 255) / 256, 1, nframes);
    k_level0<<<grid, 256, 0, st>>>(g, dIn, inFrameStride, inRowStride, dPyr, nframes);
    return cudaGetLastError();
}

cudaError_t launch_resize(const Geom& g, int level, const uint8_t* dTables, const void* tmaMaps, uint8_t* dPyr, int nframes,
                          cudaStream_t st)
{
    const LevelGeom& L = g.L[level];
    const bool fixed = L.rsPitch <= kResizePitch && L.rsRows <= 256;      /* a TMA box is at most 256 rows (tma_host.cpp) */
    const int pitch = fixed ? kResizePitch : L.rsPitch;
    const size_t smem = (size_t)pitch * L.rsRows + (size_t)L.rsRows * RT_W * sizeof(uint16_t);
    if (smem + 2048 > 48 * 1024) {      /* static shared memory counts against the 48 KB default; the opt-in is per device */
        const int most = 200 * 1024;
        cudaError_t e = fixed ? cudaFuncSetAttribute(k_resize<kResizePitch>, cudaFuncAttributeMaxDynamicSharedMemorySize, most)
                              : cudaFuncSetAttribute(k_resize<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, most);
        if (e != cudaSuccess) return e;
    }
    const LevelMaps& maps = static_cast<const TmaMaps*>(tmaMaps)->resizeSrc;
    dim3 grid((L.pitch + RT_W - 1) / RT_W, (L.h + 2 * kEdge + RT_H - 1) / RT_H, nframes);
    if (fixed) k_resize<kResizePitch><<<grid, 256, smem, st>>>(g, level, dTables, maps, dPyr);
    else k_resize<0><<<grid, 256, smem, st>>>(g, level, dTables, maps, dPyr);
    return cudaGetLastError();
}

}  // namespace sdyn
