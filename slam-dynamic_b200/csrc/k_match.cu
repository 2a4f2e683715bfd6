/* Hamming search kernels behind ORBmatcher (K7/K8).
 *   reference: Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea   src/Frame.cc:463-478, 735-800
 *              ORBmatcher::SearchByProjection(F, MapPoints)                  src/ORBmatcher.cc:45-129
 *              ORBmatcher::SearchByProjection(Cur, Last[, pairs])            src/ORBmatcher.cc:1485-1627, 407-559
 *              ORBmatcher::SearchForInitialization                           src/ORBmatcher.cc:562-677
 *              ORBmatcher::SearchByBoW(KF, F)                                src/ORBmatcher.cc:159-288
 *              ORBmatcher::DescriptorDistance / ComputeThreeMaxima           src/ORBmatcher.cc:1804-1820, 1758-1799
 *
 * Every search in the reference mutates its outputs while iterating (query k sees the claims of queries
 * < k), and breaks distance ties by candidate enumeration order.  The work is therefore split in three:
 *
 *  k_grid_build        the 64x48 grid as a CSR: keypoints sorted by (cell = ix*48+iy, index).  Because
 *                      GetFeaturesInArea walks ix outer / iy inner / in-cell insertion order, the candidates
 *                      of one ix are ONE contiguous CSR span, already in reference order.
 *  k_match_candidates  data-parallel, one thread per query: project, window, level / stereo gates, then
 *                      __popc over the 256-bit descriptors (8 x 32 bit, query held in registers).  Emits each
 *                      query's candidate list IN REFERENCE ORDER as packed records idx:16|dist:9|level:5.
 *                      This is where the integer / popc work is.
 *  k_match_resolve     one warp per search walks the queries in order; lanes scan the cached list, drop
 *                      candidates excluded by earlier claims, and a shuffle reduction yields the two smallest
 *                      (distance, position) keys — exactly the reference's first-strict-less best / second
 *                      best.  Ratio / threshold tests, claims, the rotation histogram and its three-maxima
 *                      cull follow the reference literally.
 */
#include "match_internal.h"
#include <algorithm>

namespace sdyn {

constexpr int kPoolLevelBits = 5, kPoolDistBits = 9;

__device__ __forceinline__ uint32_t pack_rec(int idx, int dist, int level)
{ return (uint32_t)idx | ((uint32_t)dist << 16) | ((uint32_t)level << 25); }
__device__ __forceinline__ int rec_idx(uint32_t r) { return r & 0xffff; }
__device__ __forceinline__ int rec_dist(uint32_t r) { return (r >> 16) & 0x1ff; }
__device__ __forceinline__ int rec_level(uint32_t r) { return (r >> 25) & 0x1f; }

__device__ __forceinline__ int job_n(const MatchJob& J) { return J.nPtr ? min(*J.nPtr, J.n) : J.n; }
__device__ __forceinline__ int job_nq(const MatchJob& J) { return J.nqPtr ? min(*J.nqPtr, J.nq) : J.nq; }

/* ---------------------------------------------------------------------------------------------------- grid */
constexpr int GB = 1024;   /* one CTA per frame: the kernel is a chain of short dependent phases, so it wants many threads */

__device__ __forceinline__ void grid_build_body(const MatchJob& J, int* smemG)
{
    if (J.mode == MM_BOW || J.mode == MM_TRI) return;
    int* cnt = smemG;                         /* kGridCells: counters, then insertion cursors */
    int* start = smemG + kGridCells;          /* kGridCells + 1: CSR offsets */
    int* sSorted = start + kGridCells + 1;    /* n: keypoint indices in cell order (sorted in shared memory) */
    __shared__ int warpSum[GB / 32 + 1];
    const int tid = threadIdx.x, n = job_n(J);
    for (int c = tid; c < kGridCells; c += GB) cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += GB) {
        const sdyn_keypoint kp = J.keysUn[i];
        /* PosInGrid: round() is half-away-from-zero; keypoints outside the grid are not indexed */
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, J.minX), J.gridWInv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, J.minY), J.gridHInv));
        int c = -1;
        if (px >= 0 && px < SDYN_GRID_COLS && py >= 0 && py < SDYN_GRID_ROWS) {
            c = px * SDYN_GRID_ROWS + py;
            atomicAdd(&cnt[c], 1);
        }
        J.cellOf[i] = c;
    }
    __syncthreads();
    /* exclusive scan of 3072 counters: 3 per thread */
    constexpr int PER = kGridCells / GB;
    static_assert(kGridCells % GB == 0, "cells per thread");
    int local[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { local[k] = cnt[tid * PER + k]; sum += local[k]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) warpSum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        const int v = warpSum[tid];
        int wi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += u; }
        warpSum[tid] = wi - v;
        if (tid == 31) warpSum[32] = wi;
    }
    __syncthreads();
    int run = warpSum[tid >> 5] + incl - sum;
#pragma unroll
    for (int k = 0; k < PER; ++k) { start[tid * PER + k] = run; cnt[tid * PER + k] = run; run += local[k]; }
    if (tid == GB - 1) start[kGridCells] = warpSum[32];
    __syncthreads();
    for (int i = tid; i < n; i += GB) {
        const int c = J.cellOf[i];
        if (c >= 0) sSorted[atomicAdd(&cnt[c], 1)] = i;
    }
    __syncthreads();
    /* in-cell order = insertion order of AssignFeaturesToGrid = ascending index */
    for (int c = tid; c < kGridCells; c += GB) {
        const int b = start[c], e = cnt[c];
        for (int i = b + 1; i < e; ++i) {
            const int v = sSorted[i];
            int j = i - 1;
            while (j >= b && sSorted[j] > v) { sSorted[j + 1] = sSorted[j]; --j; }
            sSorted[j + 1] = v;
        }
    }
    __syncthreads();
    for (int c = tid; c <= kGridCells; c += GB) J.cellOff[c] = start[c];
    /* everything the window / level gates need, next to each other in enumeration order */
    const int total = start[kGridCells];
    for (int p = tid; p < total; p += GB) {
        const int idx = sSorted[p];
        J.sorted[p] = idx;
        const sdyn_keypoint kp = J.keysUn[idx];
        J.gridEntry[p] = make_float4(kp.x, kp.y, __int_as_float(idx | (kp.octave << 24)), 0.f);
    }
}

__device__ __forceinline__ void query_order_body(const MatchJob& J);

/* One launch prepares a step's searches: CTAs [0, nGrid) build the grids of jobs[0..nGrid), CTAs [nGrid, nGrid + nOrder) sort
 * the map queries of orderJobs[0..nOrder) by window class — the two are independent and each is a chain of short dependent
 * phases, so they overlap instead of queueing behind each other. */
__global__ void __launch_bounds__(GB)
k_grid_build(const MatchJob* __restrict__ jobs, int nGrid, const MatchJob* __restrict__ orderJobs)
{
    extern __shared__ int smemG[];
    if ((int)blockIdx.x < nGrid) grid_build_body(jobs[blockIdx.x], smemG);
    else query_order_body(orderJobs[blockIdx.x - nGrid]);
}

/* ---------------------------------------------------------------------------------------------- candidates */
__device__ __forceinline__ int hamming256(const uint32_t (&q)[8], const uint8_t* d)
{
    const uint4 a = *reinterpret_cast<const uint4*>(d), b = *reinterpret_cast<const uint4*>(d + 16);
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
           __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

__device__ __forceinline__ void load_desc(const uint8_t* p, uint32_t (&q)[8])
{
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);   /* 4-byte aligned in every query layout */
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = w[k];
}

/* ------------------------------------------------------------------------------------------- query order
 * Counting sort of the map-point queries of a job by search-window size class (level, narrow / wide viewing-cosine
 * radius), largest windows first, inactive queries last.  Only the visiting order of k_match_candidates changes:
 * every result is stored under the query's own index. */
__device__ __forceinline__ void query_order_body(const MatchJob& J)
{
    if (J.mode != MM_MAP || !J.qperm) return;
    constexpr int NB = 2 * SDYN_MAX_LEVELS + 1;
    __shared__ int hist[NB];
    const int tid = threadIdx.x, nq = job_nq(J), NT = blockDim.x;
    const sdyn_mappoint_query* mps = reinterpret_cast<const sdyn_mappoint_query*>(J.queries);
    if (tid < NB) hist[tid] = 0;
    __syncthreads();
    auto key = [&](int q) {
        const sdyn_mappoint_query& m = mps[q];
        if (!m.track_in_view || m.bad) return NB - 1;
        const int lvl = min(max(m.level, 0), SDYN_MAX_LEVELS - 1);
        return 2 * (SDYN_MAX_LEVELS - 1 - lvl) + (((double)m.view_cos > 0.998) ? 1 : 0);
    };
    for (int q = tid; q < nq; q += NT) atomicAdd(&hist[key(q)], 1);
    __syncthreads();
    if (tid == 0) { int run = 0; for (int k = 0; k < NB; ++k) { const int c = hist[k]; hist[k] = run; run += c; } }
    __syncthreads();
    for (int q = tid; q < nq; q += NT) J.qperm[atomicAdd(&hist[key(q)], 1)] = q;
}

__global__ void __launch_bounds__(256)
k_query_order(const MatchJob* __restrict__ jobs) { query_order_body(jobs[blockIdx.x]); }

constexpr int kCandMaxThreads = 1024;
constexpr size_t kCandSmemBudget = 200 * 1024;
constexpr int kCellOffBytes = ((kGridCells + 1) * 4 + 15) / 16 * 16;

/* One THREAD per query; a CTA works through the queries of ONE job (search) in chunks of blockDim.x, gridDim.x
 * CTAs sharing a job.
 *  - A query touches a handful of grid columns and, after the level / window / stereo gates, evaluates only a few
 *    distances (2-3 on tracking frames): its cost is a chain of dependent loads, not arithmetic.  A warp per query
 *    leaves 31 lanes waiting on that chain; with a thread per query 32 chains overlap in one warp.
 *  - All queries of a CTA read the same frame, so the frame's grid (cell offsets + entries, ~45 KB) and, when they
 *    fit, its descriptors (64 KB) are staged in shared memory once per CTA: the ~45 entry reads per query become
 *    30-cycle shared-memory reads instead of L1-thrashing L2 reads (8 different frames per SM otherwise).
 *  - A thread walks its columns in order, so its list comes out in the reference's enumeration order by
 *    construction; candidate space is reserved with one atomic per warp batch. */
__global__ void __launch_bounds__(kCandMaxThreads)
k_match_candidates(const MatchJob* __restrict__ jobs, int stageGrid, int stageDesc, int stageCap)
{
    extern __shared__ __align__(16) uint8_t smemC[];
    /* the job descriptor is read dozens of times: one coalesced copy into shared memory per CTA */
    __shared__ __align__(16) MatchJob sJ;
    const int NT = blockDim.x;
    {
        const uint32_t* srcw = reinterpret_cast<const uint32_t*>(jobs + blockIdx.y);
        uint32_t* dstw = reinterpret_cast<uint32_t*>(&sJ);
        for (int i = threadIdx.x; i < (int)(sizeof(MatchJob) / 4); i += NT) dstw[i] = srcw[i];
    }
    __syncthreads();
    if (sJ.pose && threadIdx.x == 0) {
        /* per-frame pose of the batched front end: Tcw, and bForward / bBackward of ORBmatcher.cc:1495-1506 —
         * twc = -Rcw^T tcw (transposed operand: double accumulation), tlc = Rlw twc + tlw (3x3 float path) */
        const float* Tc = sJ.pose; const float* Tl = sJ.pose + 12;
        float twc[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            double a = 0;
#pragma unroll
            for (int k = 0; k < 3; ++k) a = __dadd_rn(a, __dmul_rn((double)Tc[4 * k + r], (double)Tc[4 * k + 3]));
            twc[r] = (float)(-1.0 * a);
        }
        const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Tl[8], twc[0]), __fmul_rn(Tl[9], twc[1])), __fmul_rn(Tl[10], twc[2])), Tl[11]);
        sJ.forward = (tz > sJ.mb) && !sJ.mono;
        sJ.backward = (-tz > sJ.mb) && !sJ.mono;
#pragma unroll
        for (int k = 0; k < 12; ++k) sJ.Tcw[k] = Tc[k];
    }
    if (sJ.pose) __syncthreads();
    const MatchJob& J = sJ;
    const int lane = threadIdx.x & 31;
    const int nq = job_nq(J);

    /* stage the searched frame: cell offsets | grid entries | descriptors */
    const int32_t* cellOff = J.cellOff;
    const float4* gridEntry = J.gridEntry;
    const uint8_t* desc = J.desc;
    const bool bowLike = J.mode == MM_BOW || J.mode == MM_TRI;     /* node-bucketed brute force: no grid */
    if (!bowLike && stageGrid) {
        int32_t* sOff = reinterpret_cast<int32_t*>(smemC);
        float4* sEnt = reinterpret_cast<float4*>(smemC + kCellOffBytes);
        for (int i = threadIdx.x; i <= kGridCells; i += NT) sOff[i] = J.cellOff[i];
        const int total = min(J.cellOff[kGridCells], stageCap);
        for (int i = threadIdx.x; i < total; i += NT) sEnt[i] = J.gridEntry[i];
        cellOff = sOff; gridEntry = sEnt;
    }
    if (stageDesc) {
        uint4* sDesc = reinterpret_cast<uint4*>(smemC + (stageGrid ? kCellOffBytes + (size_t)stageCap * 16 : 0));
        const int n2 = min(job_n(J), stageCap) * 2;
        const uint4* g = reinterpret_cast<const uint4*>(J.desc);
        for (int i = threadIdx.x; i < n2; i += NT) sDesc[i] = g[i];
        desc = reinterpret_cast<const uint8_t*>(sDesc);
    }
    __syncthreads();

  /* Warps pull batches of 32 queries from the job's counter: no CTA-wide barrier, so a warp with a long walk (large
   * window in a dense region) delays nobody, and with the largest windows first (k_query_order) the tail is short. */
  for (;;) {
    int first = 0;
    if (lane == 0) first = atomicAdd(J.qNext, 32);
    first = __shfl_sync(0xffffffffu, first, 0);
    if (first >= nq) break;
    const int slot = first + lane;
    /* queries are independent here, so they can be visited in any order: k_query_order groups queries of similar
     * window size (= walk length) so that the lanes of a warp finish together */
    const int q = slot < nq ? (J.qperm ? J.qperm[slot] : slot) : nq;

    uint32_t qd[8];
    bool active = q < nq;
    float x = 0.f, y = 0.f, r = 0.f, gate = 0.f, gateX = 0.f;   /* window centre / radius, stereo gate */
    int minLevel = 0, maxLevel = -1;
    BowQuery bq = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 8; ++k) qd[k] = 0;

    if (active) {
        if (bowLike) {
            bq = reinterpret_cast<const BowQuery*>(J.queries)[q];
            load_desc(J.qDesc + 32 * (size_t)bq.kfIdx, qd);
        } else if (J.mode == MM_FRAME) {
            const sdyn_last_point* lp = reinterpret_cast<const sdyn_last_point*>(J.queries) + q;
            active = lp->has_mp && !lp->outlier;
            if (active) {
                /* x3Dc = Rcw*x3Dw + tcw, evaluated like cv::Mat's 3x3 float product: left-to-right, no FMA */
                float pc[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float s = __fadd_rn(__fadd_rn(__fmul_rn(J.Tcw[4 * k], lp->world[0]), __fmul_rn(J.Tcw[4 * k + 1], lp->world[1])),
                                              __fmul_rn(J.Tcw[4 * k + 2], lp->world[2]));
                    pc[k] = __fadd_rn(s, J.Tcw[4 * k + 3]);
                }
                const float invzc = (float)(1.0 / (double)pc[2]);
                if (invzc < 0) active = false;
                x = __fadd_rn(__fmul_rn(__fmul_rn(J.fx, pc[0]), invzc), J.cx);
                y = __fadd_rn(__fmul_rn(__fmul_rn(J.fy, pc[1]), invzc), J.cy);
                if (x < J.minX || x > J.maxX || y < J.minY || y > J.maxY) active = false;
                const int oct = J.qKeys[q].octave;
                r = __fmul_rn(J.th, J.scale[oct]);
                if (J.forward) { minLevel = oct; maxLevel = -1; }
                else if (J.backward) { minLevel = 0; maxLevel = oct; }
                else { minLevel = oct - 1; maxLevel = oct + 1; }
                gate = r;
                gateX = __fsub_rn(x, __fmul_rn(J.bf, invzc));    /* ur = u - mbf*invzc */
                load_desc(lp->desc, qd);
            }
        } else if (J.mode == MM_POSE) {
            /* SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist) :1629-1756 (variant 0) and
             * SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) :290-403 (variant 1) */
            const sdyn_proj_point* pp = reinterpret_cast<const sdyn_proj_point*>(J.queries) + q;
            active = pp->valid;
            if (active) {
                float pc[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float s = __fadd_rn(__fadd_rn(__fmul_rn(J.Tcw[4 * k], pp->world[0]), __fmul_rn(J.Tcw[4 * k + 1], pp->world[1])),
                                              __fmul_rn(J.Tcw[4 * k + 2], pp->world[2]));
                    pc[k] = __fadd_rn(s, J.Tcw[4 * k + 3]);
                }
                if (J.poseVariant == SDYN_PROJ_FRAME_KEYFRAME) {
                    const float invzc = (float)(1.0 / (double)pc[2]);
                    x = __fadd_rn(__fmul_rn(__fmul_rn(J.fx, pc[0]), invzc), J.cx);
                    y = __fadd_rn(__fmul_rn(__fmul_rn(J.fy, pc[1]), invzc), J.cy);
                    if (x < J.minX || x > J.maxX || y < J.minY || y > J.maxY) active = false;
                } else {
                    if ((double)pc[2] < 0.0) active = false;
                    const float invz = __fdiv_rn(1.0f, pc[2]);
                    x = __fadd_rn(__fmul_rn(J.fx, __fmul_rn(pc[0], invz)), J.cx);
                    y = __fadd_rn(__fmul_rn(J.fy, __fmul_rn(pc[1], invz)), J.cy);
                    if (!(x >= J.minX && x < J.maxX && y >= J.minY && y < J.maxY)) active = false;    /* KeyFrame::IsInImage */
                }
                /* PO = p3Dw - Ow; cv::norm / Mat::dot accumulate float products in double */
                const float po0 = __fsub_rn(pp->world[0], J.Ow[0]), po1 = __fsub_rn(pp->world[1], J.Ow[1]), po2 = __fsub_rn(pp->world[2], J.Ow[2]);
                const double n2 = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)po0), __dmul_rn((double)po1, (double)po1)),
                                            __dmul_rn((double)po2, (double)po2));
                const float dist = (float)sqrt(n2);
                if (dist < pp->min_distance || dist > pp->max_distance) active = false;
                if (J.poseVariant == SDYN_PROJ_KEYFRAME_SIM3) {
                    const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)pp->normal[0]), __dmul_rn((double)po1, (double)pp->normal[1])),
                                                 __dmul_rn((double)po2, (double)pp->normal[2]));
                    if (dot < 0.5 * (double)dist) active = false;       /* PO.dot(Pn) < 0.5*dist, in double */
                }
                if (active) {
                    /* MapPoint::PredictScale: ceil(log(ratio) / mfLogScaleFactor) in float; the float logarithm is the
                     * correctly rounded value of the double one (glibc's logf in all but boundary-free cases) */
                    const float ratio = __fdiv_rn(pp->max_distance_raw, dist);
                    int nScale = (int)ceilf(__fdiv_rn((float)log((double)ratio), J.logScaleFactor));
                    nScale = nScale < 0 ? 0 : (nScale >= J.predLevels ? J.predLevels - 1 : nScale);
                    r = __fmul_rn(J.th, J.scale[nScale]);
                    minLevel = nScale - 1;
                    maxLevel = J.poseVariant == SDYN_PROJ_FRAME_KEYFRAME ? nScale + 1 : nScale;
                    load_desc(pp->desc, qd);
                }
            }
        } else if (J.mode == MM_BEST) {
            /* Fuse(pKF, vpMapPoints, th) :982-1130, Fuse(pKF, Scw, ...) :1132-1257 and the two passes of SearchBySim3
             * :1259-1483: project, gate, predict the level — the keypoint choice needs no claim bookkeeping */
            const sdyn_proj_point* pp = reinterpret_cast<const sdyn_proj_point*>(J.queries) + q;
            active = pp->valid;
            if (active) {
                float pc[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float s = __fadd_rn(__fadd_rn(__fmul_rn(J.Tcw[4 * k], pp->world[0]), __fmul_rn(J.Tcw[4 * k + 1], pp->world[1])),
                                              __fmul_rn(J.Tcw[4 * k + 2], pp->world[2]));
                    pc[k] = __fadd_rn(s, J.Tcw[4 * k + 3]);
                }
                if (J.useT2) {                           /* p3Dc2 = sR21*p3Dc1 + t21 */
                    float p2[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float s = __fadd_rn(__fadd_rn(__fmul_rn(J.T2[4 * k], pc[0]), __fmul_rn(J.T2[4 * k + 1], pc[1])),
                                                  __fmul_rn(J.T2[4 * k + 2], pc[2]));
                        p2[k] = __fadd_rn(s, J.T2[4 * k + 3]);
                    }
                    pc[0] = p2[0]; pc[1] = p2[1]; pc[2] = p2[2];
                }
                if (pc[2] < 0.0f) active = false;
                const float invz = J.invzDouble ? (float)(1.0 / (double)pc[2]) : __fdiv_rn(1.0f, pc[2]);
                x = __fadd_rn(__fmul_rn(J.fx, __fmul_rn(pc[0], invz)), J.cx);
                y = __fadd_rn(__fmul_rn(J.fy, __fmul_rn(pc[1], invz)), J.cy);
                if (!(x >= J.minX && x < J.maxX && y >= J.minY && y < J.maxY)) active = false;    /* KeyFrame::IsInImage */
                gateX = __fsub_rn(x, __fmul_rn(J.bf, invz));                                       /* ur = u - bf*invz */
                float po0, po1, po2;
                if (J.distFromCamera) { po0 = pc[0]; po1 = pc[1]; po2 = pc[2]; }
                else { po0 = __fsub_rn(pp->world[0], J.Ow[0]); po1 = __fsub_rn(pp->world[1], J.Ow[1]); po2 = __fsub_rn(pp->world[2], J.Ow[2]); }
                const double n2 = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)po0), __dmul_rn((double)po1, (double)po1)),
                                            __dmul_rn((double)po2, (double)po2));
                const float dist = (float)sqrt(n2);
                if (dist < pp->min_distance || dist > pp->max_distance) active = false;
                if (J.checkNormal) {
                    const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)pp->normal[0]), __dmul_rn((double)po1, (double)pp->normal[1])),
                                                 __dmul_rn((double)po2, (double)pp->normal[2]));
                    if (dot < 0.5 * (double)dist) active = false;
                }
                if (active) {
                    const float ratio = __fdiv_rn(pp->max_distance_raw, dist);
                    int nScale = (int)ceilf(__fdiv_rn((float)log((double)ratio), J.logScaleFactor));
                    nScale = nScale < 0 ? 0 : (nScale >= J.predLevels ? J.predLevels - 1 : nScale);
                    r = __fmul_rn(J.th, J.scale[nScale]);
                    minLevel = nScale - 1; maxLevel = nScale;
                    load_desc(pp->desc, qd);
                }
            }
        } else if (J.mode == MM_MAP) {
            const sdyn_mappoint_query* mp = reinterpret_cast<const sdyn_mappoint_query*>(J.queries) + q;
            active = mp->track_in_view && !mp->bad;
            if (active) {
                float rr = ((double)mp->view_cos > 0.998) ? 2.5f : 4.0f;      /* RadiusByViewingCos */
                if (J.th != 1.0f) rr = __fmul_rn(rr, J.th);
                r = __fmul_rn(rr, J.scale[mp->level]);
                x = mp->proj_x; y = mp->proj_y;
                minLevel = mp->level - 1; maxLevel = mp->level;
                gate = r; gateX = mp->proj_xr;
                load_desc(mp->desc, qd);
            }
        } else {   /* MM_INIT */
            const sdyn_keypoint kp = J.qKeys[q];
            active = !(kp.octave > 0);
            if (active) {
                x = J.prevMatched[2 * q]; y = J.prevMatched[2 * q + 1];
                r = (float)J.window;
                minLevel = kp.octave; maxLevel = kp.octave;
                load_desc(J.qDesc + 32 * (size_t)q, qd);
            }
        }
    }

    int cx0 = 0, cx1 = -1, cy0 = 0, cy1 = -1;
    if (active && !bowLike) {
        cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, J.minX), r), J.gridWInv)));
        cx1 = min(SDYN_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, J.minX), r), J.gridWInv)));
        cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, J.minY), r), J.gridHInv)));
        cy1 = min(SDYN_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, J.minY), r), J.gridHInv)));
        if (cx0 >= SDYN_GRID_COLS || cx1 < 0 || cy0 >= SDYN_GRID_ROWS || cy1 < 0) active = false;
    }

    /* upper bound of this query's list (all keypoints of the touched cells); ONE pool reservation per CTA */
    int bound = 0;
    if (active) {
        if (J.mode == MM_TRI) bound = 0;
        else if (J.mode == MM_BOW) bound = bq.fCnt;
        else
            for (int ix = cx0; ix <= cx1; ++ix)
                bound += cellOff[ix * SDYN_GRID_ROWS + cy1 + 1] - cellOff[ix * SDYN_GRID_ROWS + cy0];
    }
    int incl = bound;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    int base = 0;
    if (lane == 31 && incl) base = atomicAdd(J.poolUsed, incl);
    base = __shfl_sync(0xffffffffu, base, 31);
    const int tot = __shfl_sync(0xffffffffu, incl, 31);
    const bool overflow = base + tot > J.poolCap;
    if (overflow && lane == 0) J.result[2] = 1;
    const bool work = active && !overflow;           /* active implies q < nq */
    const int off = base + incl - bound;
    uint32_t* out = J.pool + off;
    int cnt = 0;

    if (!work) {
    } else if (J.mode == MM_TRI) {
        /* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:814-980: KeyFrame-1 feature q against the KeyFrame-2
         * features of the same vocabulary node.  Nothing marks a feature as taken in the reference (vbMatched2 is never
         * set), so every query is independent: the winner is the candidate of smallest distance that passes the gates,
         * the LAST one among equals (`dist > bestDist` lets ties through).  Epipolar line l = x1' F12 (:140-157). */
        const sdyn_keypoint kp1 = J.qKeys[bq.kfIdx];
        const bool stereo1 = J.qURight && J.qURight[bq.kfIdx] >= 0;
        const float la = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, J.F12[0]), __fmul_rn(kp1.y, J.F12[3])), J.F12[6]);
        const float lb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, J.F12[1]), __fmul_rn(kp1.y, J.F12[4])), J.F12[7]);
        const float lc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, J.F12[2]), __fmul_rn(kp1.y, J.F12[5])), J.F12[8]);
        const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
        uint32_t bestKey = 0xffffffffu;
        for (int k = 0; k < bq.fCnt; ++k) {
            const int idx2 = (int)J.fIndex[bq.fOff + k];
            if (!J.fValid[idx2]) continue;                                   /* vbMatched2 (never set) || pMP2, bOnlyStereo */
            const int dist = hamming256(qd, desc + 32 * (size_t)idx2);
            if (dist > SDYN_TH_LOW) continue;
            const sdyn_keypoint kp2 = J.keysUn[idx2];
            const bool stereo2 = J.uRight && J.uRight[idx2] >= 0;
            if (!stereo1 && !stereo2) {                                      /* too close to the epipole (:886-892) */
                const float dx = __fsub_rn(J.epiX, kp2.x), dy = __fsub_rn(J.epiY, kp2.y);
                if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, J.scale[kp2.octave & (SDYN_MAX_LEVELS - 1)])) continue;
            }
            const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, kp2.x), __fmul_rn(lb, kp2.y)), lc);
            if (den == 0) continue;
            const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
            if (!((double)dsqr < 3.84 * (double)J.sigma2[kp2.octave & (SDYN_MAX_LEVELS - 1)])) continue;
            bestKey = min(bestKey, ((uint32_t)dist << 20) | (uint32_t)(0xfffff - k));
        }
        int best = -1, bin = 0;
        if (bestKey != 0xffffffffu) {
            best = (int)J.fIndex[bq.fOff + (0xfffff - (int)(bestKey & 0xfffff))];
            float rot = __fsub_rn(kp1.angle, J.keysUn[best].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            bin = (int)roundf(__fmul_rn(rot, 1.0f / SDYN_HISTO_LENGTH));
            if (bin == SDYN_HISTO_LENGTH) bin = 0;
        }
        J.qAccepted[q] = best; J.qBin[q] = bin;
    } else if (J.mode == MM_BOW) {
        for (int k = 0; k < bq.fCnt; ++k) {
            const int idx = (int)J.fIndex[bq.fOff + k];
            out[k] = pack_rec(idx, hamming256(qd, desc + 32 * (size_t)idx), 0);
        }
        cnt = bq.fCnt;
    } else {
        /* The candidates of grid column ix are ONE contiguous CSR span (k_grid_build), already in reference order. */
        /* bCheckLevels = (minLevel > 0) || (maxLevel >= 0); kept iff oct >= minLevel and (maxLevel < 0 or oct <= maxLevel) */
        const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
        const int loLevel = checkLevels ? minLevel : -0x7fffffff, hiLevel = (checkLevels && maxLevel >= 0) ? maxLevel : 0x7fffffff;
        const float* __restrict__ uRight = (J.mode == MM_FRAME || J.mode == MM_MAP) ? J.uRight : nullptr;
        const bool stereoGate = uRight != nullptr;
        const bool chi2Gate = J.mode == MM_BEST && J.chi2Gate;
        /* Pass 1: walk the spans and keep what survives the level / window / stereo gates.  The entries of a span are
         * fetched four at a time before any of them is tested, so four loads are in flight per thread instead of one
         * dependent load per iteration (the walk is latency bound: ~45 entries per query, 2-3 survivors). */
        for (int ix = cx0; ix <= cx1; ++ix) {
            const int b = cellOff[ix * SDYN_GRID_ROWS + cy0], e = cellOff[ix * SDYN_GRID_ROWS + cy1 + 1];
            for (int p0 = b; p0 < e; p0 += 4) {
                float4 ge[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ge[k] = gridEntry[min(p0 + k, e - 1)];
                /* branch-free gates: predicates instead of divergent `continue`s (32 independent walks per warp) */
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int io = __float_as_int(ge[k].z);
                    const int oct = io >> 24;
                    bool ok = (p0 + k < e) & (oct >= loLevel) & (oct <= hiLevel) &
                              (fabsf(__fsub_rn(ge[k].x, x)) < r) & (fabsf(__fsub_rn(ge[k].y, y)) < r);
                    if (stereoGate) {                       /* uniform over the job */
                        const float ur = ok ? uRight[io & 0xffffff] : -1.0f;
                        ok &= !(ur > 0 && fabsf(__fsub_rn(gateX, ur)) > gate);
                    }
                    if (chi2Gate) {                         /* Fuse: reprojection error against the keypoint (:1067-1092) */
                        const float kpr = (ok && J.uRight) ? J.uRight[io & 0xffffff] : -1.0f;
                        const float ex = __fsub_rn(x, ge[k].x), ey = __fsub_rn(y, ge[k].y);
                        float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                        const float s2 = J.invSigma2[oct & (SDYN_MAX_LEVELS - 1)];
                        if (kpr >= 0) {
                            const float er = __fsub_rn(gateX, kpr);
                            e2 = __fadd_rn(e2, __fmul_rn(er, er));
                            ok &= !((double)__fmul_rn(e2, s2) > 7.8);
                        } else {
                            ok &= !((double)__fmul_rn(e2, s2) > 5.99);
                        }
                    }
                    if (ok) out[cnt] = (uint32_t)io;
                    cnt += ok;
                }
            }
        }
        /* Pass 2: distances of the survivors (lanes with survivors left run together) */
        uint32_t bestKey = 0xffffffffu;                   /* BEST: first candidate of smallest distance */
        for (int k = 0; k < cnt; ++k) {
            const int io = (int)out[k];
            const int idx = io & 0xffffff;
            const int dist = hamming256(qd, desc + 32 * (size_t)idx);
            out[k] = pack_rec(idx, dist, io >> 24);
            bestKey = min(bestKey, ((uint32_t)dist << 20) | (uint32_t)k);
        }
        if (J.mode == MM_BEST) {
            J.qAccepted[q] = bestKey != 0xffffffffu ? rec_idx(out[bestKey & 0xfffff]) : -1;
            J.qBin[q] = bestKey != 0xffffffffu ? (int)(bestKey >> 20) : 256;
        }
    }
    if (q < nq && (J.mode == MM_BEST || J.mode == MM_TRI) && !work) { J.qAccepted[q] = -1; J.qBin[q] = 256; }
    if (q < nq) J.qspan[q] = work ? make_int2(off, cnt) : make_int2(0, 0);
    /* distance evaluations (statistics): one atomic per warp */
    int ev = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ev += __shfl_xor_sync(0xffffffffu, ev, o);
    if (lane == 0 && ev) atomicAdd(&J.result[3], ev);
  }
}

/* ------------------------------------------------------------------------------------------------- resolve */
struct Top2 { uint32_t a, b; };   /* keys (dist << 20 | position); a <= b */

__device__ __forceinline__ Top2 warp_top2(uint32_t a, uint32_t b)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
        const uint32_t lo = min(a, oa), hi = max(a, oa);
        b = min(hi, min(b, ob));
        a = lo;
    }
    return {a, b};
}

/* Sequential claim logic of SearchForInitialization (:562-677) and SearchByBoW (:159-288, :679-812): a query's decision depends
 * on what every earlier query did to vMatchedDistance / the matched flags, so queries are walked in order by ONE warp (the scan
 * of a query's candidate list is lane-parallel).  What made this slow was not the walk but its memory chain — per query a
 * dependent global load of the span, of the records and of the per-keypoint state.  Now: the per-keypoint state lives in shared
 * memory; a second warp streams the records of the NEXT group of 32 queries (contiguous in the pool: one reservation per
 * warp batch of k_match_candidates) into a double buffer while the walker works on the current group; empty queries (the
 * non-level-0 keypoints of SearchForInitialization) are skipped with one ballot per group. */
constexpr int RS_BUF = 8192;                 /* records per buffer */

constexpr int RS_PW = 8;                    /* prefetch warps */
constexpr int RS_T = 32 * (1 + RS_PW);

__global__ void __launch_bounds__(RS_T)
k_match_resolve(const MatchJob* __restrict__ jobs, int stageState)
{
    extern __shared__ int smemRes[];
    const MatchJob& J = jobs[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nq = job_nq(J), n = job_n(J);
    __shared__ int hist[SDYN_HISTO_LENGTH];
    __shared__ int keep[3];
    __shared__ int2 sSpan[2][32];            /* per query of the group: (offset into the packed buffer | pool, count) */
    __shared__ int sStagedOk[2];
    __shared__ int sNmatches, sNpairs;
    constexpr uint32_t NONE = 0xffffffffu;
    /* job fields the walk reads: copied out once (`J.x` inside the loop would be a dependent global load per query) */
    const int mode = J.mode, checkOri = J.checkOri, strictLow = J.strictLow;
    const float nnratio = J.nnratio;
    const uint32_t* pool = J.pool;
    const int2* qspan = J.qspan;
    int32_t* qAccepted = J.qAccepted; int32_t* qBin = J.qBin; int32_t* assignOut = J.assign;
    const sdyn_keypoint* qKeys = J.qKeys; const sdyn_keypoint* keysUn = J.keysUn;
    const BowQuery* bowq = reinterpret_cast<const BowQuery*>(J.queries);
    const float histFactor = 1.0f / SDYN_HISTO_LENGTH;
    /* per-keypoint state of the searched frame: INIT: vMatchedDistance + vnMatches21; BOW: the matched flag (assign != -1) */
    uint32_t* buf0 = reinterpret_cast<uint32_t*>(smemRes);
    uint32_t* buf1 = buf0 + RS_BUF;
    int* stA = smemRes + 2 * RS_BUF;                       /* INIT: mdist, BOW: assign */
    int* stB = stA + (stageState ? J.n : 0);               /* INIT: m21 */
    int* mdist = mode == MM_INIT ? (stageState ? stA : J.matchedDist) : nullptr;
    int* m21 = mode == MM_INIT ? (stageState ? stB : J.m21) : nullptr;
    int* occ = mode == MM_INIT ? nullptr : (stageState ? stA : J.assign);
    if (stageState) {
        for (int k = threadIdx.x; k < n; k += RS_T) {
            if (mode == MM_INIT) { stA[k] = J.matchedDist[k]; stB[k] = J.m21[k]; }
            else stA[k] = J.assign[k];
        }
    }
    if (threadIdx.x == 0) { sNmatches = 0; sNpairs = 0; }
    __syncthreads();

    const int ngroups = (nq + 31) / 32;
    /* prefetch warps 1..RS_PW: the candidate records of group g, PACKED back to back (a query's pool slot is as large as its
     * un-gated bound, its written records are few), into buffer g & 1; warp w copies queries 4(w-1) .. 4(w-1)+3 */
    auto prefetch = [&](int g) {
        const int q = g * 32 + lane;
        const int2 sp = q < nq ? qspan[q] : make_int2(0, 0);
        int incl = sp.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31), off = incl - sp.y;
        const bool fits = total <= RS_BUF;
        if (warp == 1) {
            sSpan[g & 1][lane] = make_int2(fits ? off : sp.x, sp.y);
            if (lane == 0) sStagedOk[g & 1] = fits;
        }
        if (fits) {
            uint32_t* dst = (g & 1) ? buf1 : buf0;
            for (int j = 4 * (warp - 1); j < 4 * warp; ++j) {
                const int sx = __shfl_sync(0xffffffffu, sp.x, j), sy = __shfl_sync(0xffffffffu, sp.y, j), so = __shfl_sync(0xffffffffu, off, j);
                for (int i = lane; i < sy; i += 32) dst[so + i] = __ldg(pool + sx + i);
            }
        }
    };
    if (warp >= 1 && ngroups > 0) prefetch(0);
    __syncthreads();

    int nmatches = 0, npairs = 0;
    for (int g = 0; g < ngroups; ++g) {
        if (warp >= 1) { if (g + 1 < ngroups) prefetch(g + 1); }
        else {
            const int2 mySpan = sSpan[g & 1][lane];
            const uint32_t* recs = sStagedOk[g & 1] ? ((g & 1) ? buf1 : buf0) : pool;   /* sSpan offsets index whichever it is */
            unsigned todo = __ballot_sync(0xffffffffu, mySpan.y > 0);
            const int qme = g * 32 + lane;
            if (qme < nq && mySpan.y == 0) { qAccepted[qme] = -1; qBin[qme] = 0; }
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1;
                const int q = g * 32 + j;
                const int sx = __shfl_sync(0xffffffffu, mySpan.x, j), sy = __shfl_sync(0xffffffffu, mySpan.y, j);
                uint32_t a = NONE, b = NONE;
                for (int p = lane; p < sy; p += 32) {
                    const uint32_t rec = recs[sx + p];
                    const int idx = rec_idx(rec), dist = rec_dist(rec);
                    const bool ok = mode == MM_INIT ? !(mdist[idx] <= dist) : occ[idx] == -1;
                    if (ok) {
                        const uint32_t key = ((uint32_t)dist << 20) | (uint32_t)p;
                        if (key < a) { b = a; a = key; } else if (key < b) b = key;
                    }
                }
                const Top2 t = warp_top2(a, b);
                int accepted = -1, bin = 0;
                if (t.a != NONE) {
                    const uint32_t r1 = recs[sx + (t.a & 0xfffff)];
                    const int bestDist = (int)(t.a >> 20), bestIdx = rec_idx(r1);
                    /* second best: the reference starts from INT_MAX in SearchForInitialization, 256 in SearchByBoW */
                    int bestDist2 = mode == MM_INIT ? 0x7fffffff : 256;
                    if (t.b != NONE) bestDist2 = (int)(t.b >> 20);
                    bool ok;
                    if (mode == MM_INIT) ok = bestDist <= SDYN_TH_LOW && (float)bestDist < __fmul_rn((float)bestDist2, nnratio);
                    else ok = (strictLow ? bestDist < SDYN_TH_LOW : bestDist <= SDYN_TH_LOW) &&
                              (float)bestDist < __fmul_rn(nnratio, (float)bestDist2);
                    if (ok) {
                        accepted = bestIdx;
                        if (lane == 0) {
                            if (mode == MM_INIT) {
                                const int prev = m21[bestIdx];
                                if (prev >= 0) { assignOut[prev] = -1; --nmatches; }
                                assignOut[q] = bestIdx; m21[bestIdx] = q; mdist[bestIdx] = bestDist;
                            } else {
                                occ[bestIdx] = bowq[q].kfIdx;
                            }
                        }
                        ++nmatches; ++npairs;
                    }
                }
                if (lane == 0) qAccepted[q] = accepted;      /* the rotation bin follows after the walk, in parallel */
                (void)bin;
                __syncwarp();
            }
        }
        __syncthreads();                                   /* buffer (g + 1) & 1 is filled, buffer g & 1 is free */
    }
    if (warp == 0 && lane == 0) { sNmatches = nmatches; sNpairs = npairs; }
    __syncthreads();
    /* rotation bin of every accepted query (an entry stays in the histogram even if the keypoint is stolen later, as in the
     * reference's rotHist): the two angle loads are off the sequential walk */
    if (checkOri)
        for (int q = threadIdx.x; q < nq; q += RS_T) {
            const int acc = qAccepted[q];
            if (acc < 0) { qBin[q] = 0; continue; }
            float rot = __fsub_rn(mode == MM_INIT ? qKeys[q].angle : qKeys[bowq[q].kfIdx].angle, keysUn[acc].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, histFactor));
            if (bin == SDYN_HISTO_LENGTH) bin = 0;
            qBin[q] = bin;
        }
    /* write the staged per-keypoint state back (the walk is over: both warps help) */
    if (stageState) {
        for (int k = threadIdx.x; k < n; k += RS_T) {
            if (mode == MM_INIT) { J.matchedDist[k] = stA[k]; J.m21[k] = stB[k]; }
            else J.assign[k] = stA[k];
        }
    }
    __syncthreads();
    if (warp != 0) return;
    nmatches = sNmatches; npairs = sNpairs;

    /* rotation-consistency cull (ComputeThreeMaxima) */
    if (checkOri) {
        for (int i = lane; i < SDYN_HISTO_LENGTH; i += 32) hist[i] = 0;
        __syncwarp();
        for (int q = lane; q < nq; q += 32) if (J.qAccepted[q] >= 0) atomicAdd(&hist[J.qBin[q]], 1);
        __syncwarp();
        if (lane == 0) {
            int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < SDYN_HISTO_LENGTH; ++i) {
                const int sz = hist[i];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; i3 = i2; i2 = i1; i1 = i; }
                else if (sz > max2) { max3 = max2; max2 = sz; i3 = i2; i2 = i; }
                else if (sz > max3) { max3 = sz; i3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) i3 = -1;
            keep[0] = i1; keep[1] = i2; keep[2] = i3;
        }
        __syncwarp();
        int dec = 0;
        for (int q = lane; q < nq; q += 32) {
            const int acc = J.qAccepted[q];
            if (acc < 0) continue;
            const int bin = J.qBin[q];
            if (bin == keep[0] || bin == keep[1] || bin == keep[2]) continue;
            if (mode == MM_INIT) { if (J.assign[q] >= 0) { J.assign[q] = -1; ++dec; } }
            else { J.assign[acc] = -1; if (J.locked) J.locked[acc] = 0; ++dec; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dec += __shfl_xor_sync(0xffffffffu, dec, o);
        nmatches -= dec;
    }
    __syncwarp();
    if (mode == MM_INIT) {   /* update vbPrevMatched (:671-674) */
        for (int q = lane; q < nq; q += 32) {
            const int m = J.assign[q];
            if (m >= 0) { J.prevMatched[2 * q] = J.keysUn[m].x; J.prevMatched[2 * q + 1] = J.keysUn[m].y; }
        }
    }
    if (lane == 0) { J.result[0] = nmatches; J.result[1] = npairs; }
}

/* ------------------------------------------------------------------------------- resolve, FRAME / MAP modes
 * The two per-frame searches only interact through "keypoint already holds a map point with observations"
 * (src/ORBmatcher.cc:87-89, 1560-1562).  Let T[kp] = the first query that locks kp (-1: locked on entry,
 * INT_MAX: never).  Query q may take kp iff T[kp] >= q, and T is the minimum over accepted, observed claims:
 * a monotone system whose unique fixpoint is the sequential result (induction on q: queries < q fix T as seen
 * by q).  Jacobi iteration from "nothing locked" makes queries 0..t-1 final after sweep t, so it stops after
 * (longest conflict chain + 1) sweeps — 2-4 on real frames — and every sweep is data-parallel over queries.
 * One CTA per search; T and the per-query decisions live in shared memory. */
constexpr int RF = 1024;

__device__ __forceinline__ void resolve_fix_body(const MatchJob& J, int poolCap, int* smemFix)
{
    const int tid = threadIdx.x;
    const int n = job_n(J), nq = job_nq(J);
    int* lockT = smemFix;                 /* n */
    int* acc = smemFix + J.n;             /* nq: accepted keypoint or -1 */
    __shared__ int hist[SDYN_HISTO_LENGTH];
    __shared__ int keep[3];
    __shared__ int sCount, sDec, warpTot[RF / 32];
    constexpr uint32_t NONE = 0xffffffffu;
    const sdyn_last_point* lps = reinterpret_cast<const sdyn_last_point*>(J.queries);
    const sdyn_mappoint_query* mps = reinterpret_cast<const sdyn_mappoint_query*>(J.queries);
    const sdyn_proj_point* pps = reinterpret_cast<const sdyn_proj_point*>(J.queries);
    /* job fields the loops test: read once (every barrier is a memory fence, so `J.x` inside a loop is a fresh global load) */
    const int mode = J.mode, distTh = J.distTh, checkOri = J.checkOri, assignBase = J.assignBase;
    const float nnratio = J.nnratio;
    const bool frameLike = mode == MM_FRAME || mode == MM_POSE;      /* best-only acceptance + rotation histogram */
    /* does query q lock the keypoint it takes (the MapPoint has observations; every claim of the pose searches does) */
    auto locks = [&](int q) -> bool { return mode == MM_FRAME ? lps[q].obs_positive : (mode == MM_POSE ? true : mps[q].obs_positive); };
    auto query_angle = [&](int q) -> float { return mode == MM_POSE ? pps[q].angle : J.qKeysUn[q].angle; };

    /* Everything a sweep reads is brought into shared memory once: the initial lock state of the keypoints, the
     * "does this query lock" flags, and the candidate lists compacted back to back (a prefix sum over the list lengths
     * gives the offsets).  A sweep then touches shared memory only — the sweeps were a chain of dependent L2 round
     * trips before.  Lists that do not fit (pc records) stay in global memory and are read through the same code. */
    int* baseT = acc + J.nq;               /* n: lock state on entry (-1 locked, INT_MAX free) */
    int* sOff = baseT + J.n;               /* nq + 1: offsets of the compacted lists */
    uint32_t* seen = reinterpret_cast<uint32_t*>(sOff + J.nq + 1);       /* nq: best | second-best keypoint of the last evaluation (map search) */
    uint32_t* sPool = seen + J.nq;
    uint8_t* qLock = reinterpret_cast<uint8_t*>(sPool + poolCap);      /* nq */
    uint16_t* act = reinterpret_cast<uint16_t*>(qLock + ((J.nq + 1) & ~1));   /* queries with a non-empty list (any order) */
    __shared__ int sStaged, sNAct;
    for (int k = tid; k < n; k += RF) baseT[k] = (J.assign[k] != -1 && J.locked[k]) ? -1 : 0x7fffffff;
    for (int q = tid; q < nq; q += RF) { acc[q] = -2; qLock[q] = locks(q); }
    if (tid == 0) sNAct = 0;
    {
        /* exclusive scan of the list lengths: a contiguous chunk of queries per thread */
        const int per = (nq + RF - 1) / RF, q0 = min(tid * per, nq), q1 = min(q0 + per, nq);
        int sum = 0;
        for (int q = q0; q < q1; ++q) { const int len = J.qspan[q].y; sOff[q] = len; sum += len; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
        if ((tid & 31) == 31) warpTot[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            const int v = warpTot[tid];
            int wi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += u; }
            warpTot[tid] = wi - v;
            if (tid == 31) sStaged = wi <= poolCap;
        }
        __syncthreads();
        int run = warpTot[tid >> 5] + incl - sum;
        for (int q = q0; q < q1; ++q) { const int len = sOff[q]; sOff[q] = run; run += len; }
        if (q1 == nq && q0 < nq) sOff[nq] = run;
        if (nq == 0 && tid == 0) sOff[0] = 0;
    }
    __syncthreads();
    const bool staged = sStaged != 0;
    /* only queries that have candidates take part in the sweeps (the others are decided: no match); one shared atomic
     * per warp appends them to the active list */
    for (int q0 = 0; q0 < nq; q0 += RF) {
        const int q = q0 + tid;
        const bool has = q < nq && sOff[q + 1] > sOff[q];
        if (q < nq && !has) acc[q] = -1;
        const unsigned m = __ballot_sync(0xffffffffu, has);
        int base = 0;
        if ((tid & 31) == 0 && m) base = atomicAdd(&sNAct, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (has) act[base + __popc(m & ((1u << (tid & 31)) - 1))] = (uint16_t)q;
    }
    if (staged)
        for (int q = tid; q < nq; q += RF) {
            const int2 span = J.qspan[q];
            const int o = sOff[q];
#pragma unroll 4
            for (int p = 0; p < span.y; ++p) sPool[o + p] = __ldg(J.pool + span.x + p);
        }
    __syncthreads();
    const int nAct = sNAct;

    /* Best-only searches (FRAME, POSE) are monotone: a keypoint, once unavailable to a query, stays unavailable, so a
     * query only ever moves DOWN its list (or to "no match") and never returns.  T can then be kept in place with
     * atomicMin — a claim a query has since abandoned is harmless, because it abandoned the keypoint only after an
     * earlier query locked it — and a sweep merely re-checks "is my keypoint still mine to take" (two shared loads),
     * re-picking for the few queries that lost theirs.  The map search's ratio test against the second-best candidate
     * is not monotone (a query can go from accepted to rejected while its keypoint stays free), so it rebuilds T from
     * the current claims every sweep (below). */
    if (frameLike) {
        for (int k = tid; k < n; k += RF) lockT[k] = baseT[k];
        __syncthreads();
        for (int sweep = 0; sweep <= nq; ++sweep) {
            int changed = 0;
            for (int i = tid; i < nAct; i += RF) {
                const int q = act[i], cur = acc[q];
                if (cur == -1 || (cur >= 0 && lockT[cur] >= q)) continue;      /* rejected for good / still holds */
                int2 span;
                const uint32_t* P;
                if (staged) { span.x = sOff[q]; span.y = sOff[q + 1] - span.x; P = sPool; }
                else { span = J.qspan[q]; P = J.pool; }
                uint32_t a = NONE;                 /* best unlocked candidate: distance << 20 | list position */
                for (int p = 0; p < span.y; ++p) {
                    const uint32_t rec = P[span.x + p];
                    if (rec == NONE || lockT[rec_idx(rec)] < q) continue;
                    a = min(a, ((uint32_t)rec_dist(rec) << 20) | (uint32_t)p);
                }
                int res = -1;
                if (a != NONE && (int)(a >> 20) <= distTh) res = rec_idx(P[span.x + (a & 0xfffff)]);
                acc[q] = res;
                changed = 1;
                if (res >= 0 && qLock[q]) atomicMin(&lockT[res], q);
            }
            if (!__syncthreads_or(changed)) break;
        }
    } else {
        /* Map search.  T is rebuilt from the current claims only when it has to be: a query that gives up a keypoint it
         * was the FIRST to lock (possible only through the ratio test, when its second-best candidate gets locked)
         * leaves T too low, so the next sweep starts from a rebuilt T and re-evaluates every query.  All other sweeps
         * are incremental: T only decreases (atomicMin of new claims), and a query whose best and second-best candidates
         * of its last evaluation are both still available has nothing to re-decide (availability only shrinks between
         * rebuilds).  When a sweep changes nothing and no rebuild is pending, the smallest claim ever made on each
         * keypoint is a current one, i.e. T is exact and the state is the fixpoint. */
        bool rebuild = true;
        for (int sweep = 0; sweep <= 2 * nq + 1; ++sweep) {
            const bool full = rebuild;
            if (rebuild) {
                for (int k = tid; k < n; k += RF) lockT[k] = baseT[k];
                __syncthreads();
                for (int q = tid; q < nq; q += RF) {
                    const int a = acc[q];
                    if (a >= 0 && qLock[q]) atomicMin(&lockT[a], q);
                }
                __syncthreads();
                rebuild = false;
            }
            int changed = 0, retract = 0;
            for (int i = tid; i < nAct; i += RF) {
                const int q = act[i], cur = acc[q];
                if (!full && cur != -2) {
                    const uint32_t s2 = seen[q];
                    const int b1 = (int)(s2 & 0xffffu), b2 = (int)(s2 >> 16);
                    if ((b1 == 0xffff || lockT[b1] >= q) && (b2 == 0xffff || lockT[b2] >= q)) continue;
                }
                int2 span;
                const uint32_t* P;
                if (staged) { span.x = sOff[q]; span.y = sOff[q + 1] - span.x; P = sPool; }
                else { span = J.qspan[q]; P = J.pool; }
                /* best and second-best unlocked candidate as keys distance << 20 | list position (first of equals wins) */
                uint32_t a = NONE, b = NONE;
                for (int p = 0; p < span.y; ++p) {
                    const uint32_t rec = P[span.x + p];
                    if (rec == NONE || lockT[rec_idx(rec)] < q) continue;
                    const uint32_t key = ((uint32_t)rec_dist(rec) << 20) | (uint32_t)p;
                    if (key < a) { b = a; a = key; } else if (key < b) b = key;
                }
                int res = -1;
                uint32_t s2 = 0xffffffffu;
                if (a != NONE) {
                    const uint32_t r1 = P[span.x + (a & 0xfffff)];
                    const int bestDist = (int)(a >> 20);
                    int bestDist2 = 256, bestLevel2 = -1;
                    s2 = (uint32_t)rec_idx(r1) | 0xffff0000u;
                    if (b != NONE) {
                        const uint32_t r2 = P[span.x + (b & 0xfffff)];
                        bestDist2 = (int)(b >> 20); bestLevel2 = rec_level(r2);
                        s2 = (uint32_t)rec_idx(r1) | ((uint32_t)rec_idx(r2) << 16);
                    }
                    if (bestDist <= distTh && !(rec_level(r1) == bestLevel2 && (float)bestDist > __fmul_rn(nnratio, (float)bestDist2)))
                        res = rec_idx(r1);
                }
                seen[q] = s2;
                if (res != cur) {
                    if (cur >= 0 && qLock[q] && lockT[cur] == q) retract = 1;      /* gives up a keypoint it was first to lock */
                    acc[q] = res;
                    changed = 1;
                    if (res >= 0 && qLock[q]) atomicMin(&lockT[res], q);
                }
            }
            const int anyChanged = __syncthreads_or(changed);
            rebuild = __syncthreads_or(retract) != 0;
            if (!anyChanged) break;
        }
    }

    /* final owners: the last accepted claimant of every keypoint */
    if (tid == 0) { sCount = 0; sDec = 0; }
    for (int i = tid; i < SDYN_HISTO_LENGTH; i += RF) hist[i] = 0;
    for (int k = tid; k < n; k += RF) lockT[k] = -1;          /* reuse: max claimant */
    __syncthreads();
    int mine = 0;
    for (int q = tid; q < nq; q += RF) {
        const int a = acc[q];
        if (a < 0) { if (frameLike) J.qBin[q] = -1; continue; }
        ++mine;
        atomicMax(&lockT[a], q);
        if (frameLike && checkOri) {
            float rot = __fsub_rn(query_angle(q), J.keysUn[a].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, 1.0f / SDYN_HISTO_LENGTH));
            if (bin == SDYN_HISTO_LENGTH) bin = 0;
            J.qBin[q] = bin;
            atomicAdd(&hist[bin], 1);
        }
    }
    atomicAdd(&sCount, mine);
    __syncthreads();
    for (int k = tid; k < n; k += RF) {
        const int q = lockT[k];
        if (q >= 0) {
            J.assign[k] = assignBase + q;
            J.locked[k] = qLock[q];
        }
    }
    /* point pairs of the fork's overload, in query order (before the cull, Appendix B-8) */
    if (mode == MM_FRAME && J.pairs) {
        int base = 0;
        for (int q0 = 0; q0 < nq; q0 += RF) {
            const int q = q0 + tid;
            const bool ok = q < nq && acc[q] >= 0;
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if ((tid & 31) == 0) warpTot[tid >> 5] = __popc(bal);
            __syncthreads();
            int pos = base, tot = 0;
            for (int w = 0; w < RF / 32; ++w) { if (w < (tid >> 5)) pos += warpTot[w]; tot += warpTot[w]; }
            pos += __popc(bal & ((1u << (tid & 31)) - 1));
            if (ok) {
                const int a = acc[q];
                J.pairs[4 * pos] = J.qKeysUn[q].x; J.pairs[4 * pos + 1] = J.qKeysUn[q].y;
                J.pairs[4 * pos + 2] = J.keysUn[a].x; J.pairs[4 * pos + 3] = J.keysUn[a].y;
            }
            base += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    if (frameLike && checkOri) {
        if (tid == 0) {
            int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < SDYN_HISTO_LENGTH; ++i) {
                const int sz = hist[i];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; i3 = i2; i2 = i1; i1 = i; }
                else if (sz > max2) { max3 = max2; max2 = sz; i3 = i2; i2 = i; }
                else if (sz > max3) { max3 = sz; i3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) i3 = -1;
            keep[0] = i1; keep[1] = i2; keep[2] = i3;
        }
        __syncthreads();
        int dec = 0;
        for (int q = tid; q < nq; q += RF) {
            const int a = acc[q];
            if (a < 0) continue;
            const int bin = J.qBin[q];
            if (bin == keep[0] || bin == keep[1] || bin == keep[2]) continue;
            J.assign[a] = -1; J.locked[a] = 0; ++dec;      /* every histogram entry of a culled bin nulls its keypoint */
        }
        atomicAdd(&sDec, dec);
        __syncthreads();
    }
    if (tid == 0) { J.result[0] = sCount - sDec; J.result[1] = sCount; }
}

/* second > 0: the CTA resolves jobs[b] and then jobs[b + second] — the frame search and the map search of the same frame,
 * which only interact through that frame's assign / locked arrays (one launch instead of two dependent ones). */
__global__ void __launch_bounds__(RF, 1)
k_match_resolve_fix(const MatchJob* __restrict__ jobs, int poolCap, int second)
{
    extern __shared__ int smemFix[];
    resolve_fix_body(jobs[blockIdx.x], poolCap, smemFix);
    if (second > 0) {
        __syncthreads();                   /* assign / locked of the first search are visible to the whole CTA */
        resolve_fix_body(jobs[blockIdx.x + second], poolCap, smemFix);
    }
}

cudaError_t launch_grid_build(const MatchJob* dJobs, int njobs, int maxN, cudaStream_t st, const MatchJob* orderJobs, int nOrder)
{
    const size_t smem = (size_t)(2 * kGridCells + 1 + std::max(maxN, 1)) * sizeof(int);
    if (smem + 2048 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_grid_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_grid_build<<<njobs + (orderJobs ? nOrder : 0), GB, smem, st>>>(dJobs, njobs, orderJobs);
    return cudaGetLastError();
}

cudaError_t launch_query_order(const MatchJob* dJobs, int njobs, cudaStream_t st)
{
    k_query_order<<<njobs, 256, 0, st>>>(dJobs);
    return cudaGetLastError();
}

cudaError_t launch_match_candidates(const MatchJob* dJobs, int njobs, int maxQueries, int maxN, cudaStream_t st)
{
    if (maxQueries <= 0) return cudaSuccess;
    /* many jobs: one 1024-thread CTA per job (one job per SM, everything staged once); few jobs (the single-search
     * entry points): 256-thread CTAs, one per chunk, so a lone search still spreads over the GPU */
    const int threads = njobs >= 32 ? kCandMaxThreads : 256;
    const int ctasPerJob = njobs >= 32 ? 1 : (maxQueries + threads - 1) / threads;
    const size_t gridBytes = (size_t)kCellOffBytes + (size_t)maxN * 16, descBytes = (size_t)maxN * 32;
    const int stageGrid = gridBytes <= kCandSmemBudget, stageDesc = (stageGrid ? gridBytes : 0) + descBytes <= kCandSmemBudget;
    const size_t smem = (stageGrid ? gridBytes : 0) + (stageDesc ? descBytes : 0);
    /* the opt-in is per device and cheap: no process-wide cache (contexts of several devices / threads share this code) */
    if (smem + 2048 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_match_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCandSmemBudget);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(ctasPerJob, njobs);
    k_match_candidates<<<grid, threads, smem, st>>>(dJobs, stageGrid, stageDesc, maxN);
    return cudaGetLastError();
}

cudaError_t launch_match_resolve(const MatchJob* dJobs, int njobs, int mode, int maxN, int maxQ, cudaStream_t st, int second)
{
    if (mode == MM_FRAME || mode == MM_MAP || mode == MM_POSE) {
        /* lockT + baseT (maxN each), acc + sOff + seen (maxQ each), qLock bytes, the active list, and whatever is left for
         * the compacted lists */
        const size_t fixedB = (size_t)(2 * maxN + 3 * maxQ + 1) * sizeof(int) + (size_t)maxQ * 3 + 16;   /* + qLock bytes + active list */
        const size_t budget = 200 * 1024;
        if (fixedB + 4096 > budget) return cudaErrorInvalidValue;
        const int poolCap = (int)((budget - fixedB) / 4);
        const size_t smem = fixedB + (size_t)poolCap * 4;
        cudaError_t e = cudaFuncSetAttribute(k_match_resolve_fix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_match_resolve_fix<<<njobs, RF, smem, st>>>(dJobs, poolCap, second);
    } else {
        /* record double buffer + the per-keypoint state (2 ints per keypoint for SearchForInitialization) when it fits */
        const size_t stateB = (size_t)2 * maxN * sizeof(int);
        const int stage = 2 * RS_BUF * 4 + stateB <= 200 * 1024;
        const size_t smem = (size_t)2 * RS_BUF * 4 + (stage ? stateB : 0);
        cudaError_t e = cudaFuncSetAttribute(k_match_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_match_resolve<<<njobs, RS_T, smem, st>>>(dJobs, stage);
    }
    return cudaGetLastError();
}

}  // namespace sdyn
