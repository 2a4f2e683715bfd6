/* Device-resident dynamic-mask stage of the batched front end (sdyn_track_batch_device).
 *   reference: Frame::firstSeparate (+ tail split)   src/Frame.cc:555-604, 337-367
 *              Tracking::Separate + classifyF        src/Tracking.cc:1093-1239, 1311-1367
 *              Frame::UpdateFrame                    src/Frame.cc:607-641
 * mask[i] = in_box(i) && !readmitted(i) per extracted keypoint.  The box bookkeeping of firstSeparate —
 * including its erase-while-iterating behaviour (SURVEY B-4) — is replayed per frame from the 64-bit
 * "which boxes contain a keypoint" word, so box slot s of Tracking::Separate gets the geometry / reference
 * join of the s-th SURVIVING box and the keypoints of the s-th OCCUPIED box, exactly as the reference does.
 */
#include "match_internal.h"
#include <algorithm>

namespace sdyn {

__global__ void __launch_bounds__(256)
k_box_occupancy(const uint64_t* __restrict__ mask, const int32_t* __restrict__ count, int cap,
                unsigned long long* __restrict__ has, const int32_t* __restrict__ nBoxes, long long pitch, int8_t* __restrict__ slotMap)
{
    const int f = blockIdx.x, n = min(count[f], cap);
    unsigned long long m = 0;
    for (int i = threadIdx.x; i < n; i += 256) m |= mask[(size_t)f * cap + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    __shared__ unsigned long long w[8];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) m |= w[k];
        has[f] = m;
        /* replay of firstSeparate's erase loop (Frame.cc:585-592), once per frame: `i` advances after an erase and hasKpts keeps
         * its indexing, so slot s of Tracking::Separate gets the geometry / reference join of the s-th SURVIVING box
         * (slotMap[s]) and the keypoints of the s-th OCCUPIED box (slotMap[64 + s]); -1 = no such slot */
        const int nb = min(*frame_part(nBoxes, f, 4, pitch), 64);
        int8_t* sm = slotMap + (size_t)f * 128;
        int8_t ids[64];
        int size = nb;
        for (int i = 0; i < nb; ++i) ids[i] = (int8_t)i;
        for (int i = 0; i < size; ++i) {
            if ((m >> i) & 1ull) continue;
            for (int k = i; k + 1 < size; ++k) ids[k] = ids[k + 1];
            --size;
        }
        int seen = 0;
        for (int s = 0; s < 64; ++s) { sm[s] = s < size ? ids[s] : (int8_t)-1; sm[64 + s] = -1; }
        for (int b = 0; b < nb; ++b) if ((m >> b) & 1ull) sm[64 + seen++] = (int8_t)b;
    }
}

__device__ __forceinline__ int hamming_bytes2(const uint8_t* a, const uint8_t* b)
{
    const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 16);
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 16);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

constexpr int BS = 256;

/* one CTA per (box slot, frame) */
__global__ void __launch_bounds__(BS)
k_box_stage(const sdyn_track_inputs in, const sdyn_keypoint* __restrict__ kp, const sdyn_keypoint* __restrict__ kpUn,
            const uint8_t* __restrict__ desc,
            const int32_t* __restrict__ count, int cap, const uint64_t* __restrict__ mask,
            const int8_t* __restrict__ slotMap, int32_t* __restrict__ boxList, int32_t* __restrict__ nnQ,
            int32_t* __restrict__ nnT, int nnTStride, int32_t* __restrict__ readmit, int32_t* __restrict__ staticExit)
{
    const int s = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const long long P = in.frame_pitch;
    const int nb = min(*frame_part(in.n_boxes, f, 4, P), 64);
    if (s >= nb) return;
    __shared__ int sCount, sBase, warpCnt[BS / 32], sStatic;
    const int surv = slotMap[(size_t)f * 128 + s], occ = slotMap[(size_t)f * 128 + 64 + s];
    if (surv < 0 || occ < 0) return;                 /* slot beyond objects.size(), or no keypoints in it */
    if (tid == 0) { sCount = 0; sBase = 0; sStatic = 0; }
    __syncthreads();
    const int r = frame_part(in.ref_box, f, 64 * 4, P)[surv];
    if (r < 0) return;                               /* box id not present in the reference frame */
    const int n = min(count[f], cap);
    /* classifyF reads mvdynKeysUn (Tracking.cc:1129-1131: undistorted points); the box test above used mvKeys (Frame.cc:562) */
    const sdyn_keypoint* K = kpUn + (size_t)f * cap;
    const uint8_t* D = desc + (size_t)f * cap * 32;
    const uint64_t* M = mask + (size_t)f * cap;
    int32_t* list = boxList + ((size_t)f * 64 + s) * cap;
    int32_t* nq_ = nnQ + ((size_t)f * 64 + s) * cap;
    int32_t* nt_ = nnT + ((size_t)f * 64 + s) * nnTStride;

    /* keypoints of box `occ`, ascending index = order of mvdynKeys[slot].  A thread owns a contiguous run of the frame's
     * keypoints: count, one block scan, write — the mask words are independent loads (a strided loop with a barrier per
     * 128 keypoints paid one global-memory latency per iteration and was most of this kernel's time) */
    {
        const int run = (n + BS - 1) / BS, i0 = tid * run, i1 = min(i0 + run, n);
        int mine = 0;
        unsigned long long flags = 0;                /* bit k: keypoint i0 + k is in the box (runs of up to 64 keypoints) */
        for (int b = i0; b < i1; b += 8) {
            uint64_t m[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = b + k < i1 ? M[b + k] : 0ull;     /* eight loads in flight */
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned long long in = (m[k] >> occ) & 1ull;
                mine += (int)in;
                if (b - i0 + k < 64) flags |= in << (b - i0 + k);
            }
        }
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
        if ((tid & 31) == 31) warpCnt[tid >> 5] = incl;
        __syncthreads();
        int pos = incl - mine;
        for (int w = 0; w < (tid >> 5); ++w) pos += warpCnt[w];
        while (flags) { const int k = __ffsll((long long)flags) - 1; flags &= flags - 1; list[pos++] = i0 + k; }
        for (int i = i0 + 64; i < i1; ++i)           /* only for runs longer than 64 keypoints */
            if ((M[i] >> occ) & 1ull) list[pos++] = i;
        if (tid == BS - 1) sBase = pos;
        __syncthreads();
    }
    const int nq = sBase;
    const int32_t* refOff = frame_part(in.ref_off, f, 65 * 4, P);
    const int to = refOff[r], nt = min(refOff[r + 1] - to, nnTStride);
    if (nq == 0 || nt == 0) return;                  /* mdynDescriptors[..].cols == 0 */
    const uint8_t* TD = frame_part(in.ref_desc, f, (size_t)in.ref_stride * 32, P) + (size_t)to * 32;
    const float* TX = frame_part(in.ref_xy, f, (size_t)in.ref_stride * 8, P) + (size_t)to * 2;

    /* BFMatcher(NORM_HAMMING, crossCheck = true).  A warp owns a query (then a target) and its lanes share out the other side:
     * the best match is the minimum of (distance << 20 | index) — smallest distance, then smallest index, which is what the
     * sequential scan with a strict comparison keeps — reduced over the warp by shuffles. */
    {
        const int lane = tid & 31, wid = tid >> 5;
        for (int i = wid; i < nq; i += BS / 32) {
            const uint4* q = reinterpret_cast<const uint4*>(D + 32 * (size_t)list[i]);
            const uint4 q0 = q[0], q1 = q[1];
            unsigned best = 0xffffffffu;
            for (int j = lane; j < nt; j += 32) {
                const uint4* t = reinterpret_cast<const uint4*>(TD + 32 * (size_t)j);
                const uint4 t0 = t[0], t1 = t[1];
                const int d = __popc(q0.x ^ t0.x) + __popc(q0.y ^ t0.y) + __popc(q0.z ^ t0.z) + __popc(q0.w ^ t0.w) +
                              __popc(q1.x ^ t1.x) + __popc(q1.y ^ t1.y) + __popc(q1.z ^ t1.z) + __popc(q1.w ^ t1.w);
                best = min(best, ((unsigned)d << 20) | (unsigned)j);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            if (lane == 0) nq_[i] = (int)(best & 0xfffffu);
        }
        for (int j = wid; j < nt; j += BS / 32) {
            const uint4* t = reinterpret_cast<const uint4*>(TD + 32 * (size_t)j);
            const uint4 t0 = t[0], t1 = t[1];
            unsigned best = 0xffffffffu;
            for (int i = lane; i < nq; i += 32) {
                const uint4* q = reinterpret_cast<const uint4*>(D + 32 * (size_t)list[i]);
                const uint4 q0 = q[0], q1 = q[1];
                const int d = __popc(q0.x ^ t0.x) + __popc(q0.y ^ t0.y) + __popc(q0.z ^ t0.z) + __popc(q0.w ^ t0.w) +
                              __popc(q1.x ^ t1.x) + __popc(q1.y ^ t1.y) + __popc(q1.z ^ t1.z) + __popc(q1.w ^ t1.w);
                best = min(best, ((unsigned)d << 20) | (unsigned)i);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            if (lane == 0) nt_[j] = (int)(best & 0xfffffu);
        }
    }
    __syncthreads();

    /* classifyF on the mutual matches; num0 = #static */
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = frame_part(in.fmat, f, 36, P)[k];
    int nmatch = 0, nstatic = 0;
    for (int i = tid; i < nq; i += BS) {
        const int tr = nq_[i];
        if (tr < 0 || nt_[tr] != i) continue;
        ++nmatch;
        const float u1 = TX[2 * tr], v1 = TX[2 * tr + 1];
        const float u2 = K[list[i]].x, v2 = K[list[i]].y;
        const float th = 5.841f;
        const float a2 = __fadd_rn(__fadd_rn(__fmul_rn(m[0], u1), __fmul_rn(m[1], v1)), m[2]);
        const float b2 = __fadd_rn(__fadd_rn(__fmul_rn(m[3], u1), __fmul_rn(m[4], v1)), m[5]);
        const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(m[6], u1), __fmul_rn(m[7], v1)), m[8]);
        const float num2 = __fadd_rn(__fadd_rn(__fmul_rn(a2, u2), __fmul_rn(b2, v2)), c2);
        const float d1 = __fdiv_rn(__fmul_rn(num2, num2), __fadd_rn(__fmul_rn(a2, a2), __fmul_rn(b2, b2)));
        const float a1 = __fadd_rn(__fadd_rn(__fmul_rn(m[0], u2), __fmul_rn(m[3], v2)), m[6]);
        const float b1 = __fadd_rn(__fadd_rn(__fmul_rn(m[1], u2), __fmul_rn(m[4], v2)), m[7]);
        const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(m[2], u2), __fmul_rn(m[5], v2)), m[8]);
        const float num1 = __fadd_rn(__fadd_rn(__fmul_rn(a1, u1), __fmul_rn(b1, v1)), c1);
        const float d2 = __fdiv_rn(__fmul_rn(num1, num1), __fadd_rn(__fmul_rn(a1, a1), __fmul_rn(b1, b1)));
        const bool st = d1 <= th && d2 <= th;
        nq_[i] = st ? -2 - tr : tr;                  /* mark static matches for the second pass */
        nstatic += st;
    }
    atomicAdd(&sCount, nmatch);
    atomicAdd(&sStatic, nstatic);
    __syncthreads();
    const int good = sCount, num0 = sStatic;
    /* gates of Tracking::Separate (:1125, :1151) */
    if (good < 3 || (double)good < 0.2 * (double)nq) return;
    for (int i = tid; i < nq; i += BS)
        if (nq_[i] <= -2) atomicMax(&readmit[(size_t)f * cap + list[i]], 64 - s);   /* dynStatus[slot][m] != -1; keeps the FIRST slot */
    if (tid == 0 && (double)num0 > fmax(1.0, 0.2 * (double)good)) atomicOr(&staticExit[f], 1);
}

__global__ void __launch_bounds__(256)
k_dyn_finalize(const uint64_t* __restrict__ mask, const int32_t* __restrict__ readmit,
               const int32_t* __restrict__ staticExit, const int32_t* __restrict__ count, int cap,
               uint8_t* __restrict__ dynMask, int32_t* __restrict__ counts)
{
    const int f = blockIdx.x, n = min(count[f], cap);
    const bool upd = staticExit[f] != 0;             /* if (Separate(...) == 1) UpdateFrame(dynStatus) */
    int inBox = 0, masked = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const bool in = mask[(size_t)f * cap + i] != 0;
        const bool dyn = in && !(upd && readmit[(size_t)f * cap + i]);
        dynMask[(size_t)f * cap + i] = dyn;
        inBox += in; masked += dyn;
    }
    __shared__ int a[2];
    if (threadIdx.x == 0) a[0] = a[1] = 0;
    __syncthreads();
    atomicAdd(&a[0], inBox); atomicAdd(&a[1], masked);
    __syncthreads();
    if (threadIdx.x == 0) { counts[f * 4 + 2] = a[0]; counts[f * 4 + 3] = a[1]; }
}

cudaError_t launch_dyn_stage(const sdyn_track_inputs& in, const sdyn_keypoint* kp, const sdyn_keypoint* kpUn, const uint8_t* desc, const int32_t* count,
                             int cap, uint64_t* mask, unsigned long long* has, int8_t* slotMap, int32_t* boxList, int32_t* nnQ, int32_t* nnT,
                             int nnTStride, int32_t* readmit, int32_t* staticExit, uint8_t* dynMask, int32_t* counts,
                             int nframes, cudaStream_t st)
{
    cudaError_t e = launch_box_mask(kp, count, cap, cap, in.boxes, in.n_boxes, 64, 64, mask, nframes, st,
                                    in.frame_pitch > 0 ? (size_t)in.frame_pitch : 0, in.frame_pitch > 0 ? (size_t)in.frame_pitch : 0);
    if (e != cudaSuccess) return e;
    k_box_occupancy<<<nframes, 256, 0, st>>>(mask, count, cap, has, in.n_boxes, in.frame_pitch, slotMap);
    dim3 grid(64, nframes);
    k_box_stage<<<grid, BS, 0, st>>>(in, kp, kpUn, desc, count, cap, mask, slotMap, boxList, nnQ, nnT, nnTStride, readmit, staticExit);
    k_dyn_finalize<<<nframes, 256, 0, st>>>(mask, readmit, staticExit, count, cap, dynMask, counts);
    return cudaGetLastError();
}

/* ---- RGB-D-constructor form: the frame the reference tracks with ------------------------------------------------------
 * Frame::firstSeparate reorders the keypoints "static first" and the tail split moves the in-box ones out of the frame
 * (src/Frame.cc:555-604, 337-367); after Tracking::Separate, Frame::UpdateFrame appends the re-admitted ones (:607-641) in
 * push order: box slot ascending, position in the box's list ascending, first occurrence wins (the class_id set).  A box
 * list is in extraction order, so the push order is the lexicographic order of (first re-admitting slot, extraction index).
 * One CTA per frame writes that list: order[], keypoints (class_id = extraction index for re-admitted ones, Frame.cc:565-568),
 * undistorted keypoints, descriptors, N and N_s. */
constexpr int FC = 256;

__global__ void __launch_bounds__(FC)
k_frame_compact(const sdyn_keypoint* __restrict__ kp, const sdyn_keypoint* __restrict__ kpUn, const uint8_t* __restrict__ desc,
                const int32_t* __restrict__ count, int cap, const uint64_t* __restrict__ mask, const int32_t* __restrict__ readmit,
                const int32_t* __restrict__ staticExit, sdyn_keypoint* __restrict__ fKp, sdyn_keypoint* __restrict__ fKpUn,
                uint8_t* __restrict__ fDesc, int32_t* __restrict__ fOrder, int32_t* __restrict__ fCount, int32_t* __restrict__ fStatic)
{
    extern __shared__ int32_t sR[];                 /* re-admitted: idx[cap], slot[cap] */
    __shared__ int warpCnt[FC / 32], sBase;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = min(count[f], cap);
    const uint64_t* M = mask + (size_t)f * cap;
    const int32_t* RA = readmit + (size_t)f * cap;
    int32_t* ord = fOrder + (size_t)f * cap;
    const bool upd = staticExit[f] != 0;
    int32_t* rIdx = sR; int32_t* rSlot = sR + cap;
    /* two stable compactions in extraction order: pass 0 = outside every box, pass 1 = re-admitted */
    int ns = 0, nr = 0;
    for (int pass = 0; pass < 2; ++pass) {
        if (tid == 0) sBase = 0;
        __syncthreads();
        if (pass == 1 && !upd) break;
        for (int i0 = 0; i0 < n; i0 += FC) {
            const int i = i0 + tid;
            const bool ok = i < n && (pass == 0 ? M[i] == 0 : (M[i] != 0 && RA[i] != 0));
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) warpCnt[wid] = __popc(bal);
            __syncthreads();
            int pos = sBase;
            for (int w = 0; w < wid; ++w) pos += warpCnt[w];
            pos += __popc(bal & ((1u << lane) - 1));
            if (ok) { if (pass == 0) ord[pos] = i; else { rIdx[pos] = i; rSlot[pos] = 64 - RA[i]; } }
            __syncthreads();
            if (tid == 0) { int t = 0; for (int w = 0; w < FC / 32; ++w) t += warpCnt[w]; sBase += t; }
            __syncthreads();
        }
        if (pass == 0) ns = sBase; else nr = sBase;
        __syncthreads();
    }
    /* rank of every re-admitted keypoint in (slot, index) order */
    for (int t = tid; t < nr; t += FC) {
        const int st = rSlot[t];
        int rank = 0;
        for (int u = 0; u < nr; ++u) rank += (rSlot[u] < st) || (rSlot[u] == st && u < t);
        ord[ns + rank] = rIdx[t];
    }
    __syncthreads();
    const int N = ns + nr;
    if (tid == 0) { fCount[f] = N; fStatic[f] = ns; }
    const sdyn_keypoint* K = kp + (size_t)f * cap; const sdyn_keypoint* KU = kpUn + (size_t)f * cap;
    for (int j = tid; j < N; j += FC) {
        const int src = ord[j];
        sdyn_keypoint a = K[src], b = KU[src];
        if (j >= ns) { a.class_id = src; b.class_id = src; }
        fKp[(size_t)f * cap + j] = a;
        if (fKpUn != fKp) fKpUn[(size_t)f * cap + j] = b;
    }
    const uint4* D = reinterpret_cast<const uint4*>(desc + (size_t)f * cap * 32);
    uint4* FD = reinterpret_cast<uint4*>(fDesc + (size_t)f * cap * 32);
    for (int w = tid; w < 2 * N; w += FC) FD[w] = D[2 * ord[w >> 1] + (w & 1)];
}

cudaError_t launch_frame_compact(const sdyn_keypoint* kp, const sdyn_keypoint* kpUn, const uint8_t* desc, const int32_t* count, int cap,
                                 const uint64_t* mask, const int32_t* readmit, const int32_t* staticExit, sdyn_keypoint* fKp,
                                 sdyn_keypoint* fKpUn, uint8_t* fDesc, int32_t* fOrder, int32_t* fCount, int32_t* fStatic, int nframes,
                                 cudaStream_t st)
{
    const size_t smem = (size_t)cap * 8;
    if (smem > 48 * 1024) {             /* the opt-in is per device and per function: set whenever it is needed */
        cudaError_t e = cudaFuncSetAttribute(k_frame_compact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_frame_compact<<<nframes, FC, smem, st>>>(kp, kpUn, desc, count, cap, mask, readmit, staticExit, fKp, fKpUn, fDesc, fOrder, fCount, fStatic);
    return cudaGetLastError();
}

/* ---- resident query forms: materialise the searches' query records from ids + the MapPoint table --------------------
 * blockIdx.y = frame; blockIdx.z = 0: LastFrame points (sdyn_last_point), 1: local-map points (sdyn_mappoint_query). */
__global__ void __launch_bounds__(256)
k_gather_queries(const sdyn_map_point* __restrict__ table, int tableCap,
                 const int32_t* __restrict__ lastIds, const uint8_t* __restrict__ lastFlags, const int32_t* __restrict__ nLast, int lastStride,
                 sdyn_last_point* __restrict__ gLast,
                 const int32_t* __restrict__ mapIds, const sdyn_map_proj* __restrict__ mapProj, const int32_t* __restrict__ nMap, int mapStride,
                 sdyn_mappoint_query* __restrict__ gMap, const uint8_t* __restrict__ mapFlags, const float* __restrict__ poses,
                 const FrustumParams fp, long long P)
{
    const int f = blockIdx.y, i = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.z == 0) {
        if (!lastIds || i >= min(nLast[f], lastStride)) return;          /* nLast: the resident LastFrame's own counts */
        const size_t at = (size_t)f * lastStride + i;
        const int id = frame_part(lastIds, f, (size_t)lastStride * 4, P)[i];
        sdyn_last_point o;
        const uint8_t fl = frame_part(lastFlags, f, (size_t)lastStride, P)[i];
        const bool has = id >= 0 && id < tableCap;
        o.has_mp = has; o.outlier = fl & SDYN_LP_OUTLIER ? 1 : 0; o.obs_positive = fl & SDYN_LP_OBS_POSITIVE ? 1 : 0; o.pad = 0;
        uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
        o.world[0] = o.world[1] = o.world[2] = 0.f;
        if (has) {
            const sdyn_map_point* p = table + id;
            o.world[0] = p->world[0]; o.world[1] = p->world[1]; o.world[2] = p->world[2];
            d0 = *reinterpret_cast<const uint4*>(p->desc); d1 = *reinterpret_cast<const uint4*>(p->desc + 16);
        }
        uint32_t* w = reinterpret_cast<uint32_t*>(o.desc);
        w[0] = d0.x; w[1] = d0.y; w[2] = d0.z; w[3] = d0.w; w[4] = d1.x; w[5] = d1.y; w[6] = d1.z; w[7] = d1.w;
        gLast[at] = o;
    } else {
        if (!mapIds || i >= min(*frame_part(nMap, f, 4, P), mapStride)) return;
        const size_t at = (size_t)f * mapStride + i;
        const int id = frame_part(mapIds, f, (size_t)mapStride * 4, P)[i];
        sdyn_map_proj pr;
        if (mapProj) pr = frame_part(mapProj, f, (size_t)mapStride * sizeof(sdyn_map_proj), P)[i];
        else {
            /* Frame::isInFrustum (src/Frame.cc:677-733) for table entry `id` under this frame's pose */
            const uint8_t fl = frame_part(mapFlags, f, (size_t)mapStride, P)[i];
            pr.proj_x = pr.proj_y = pr.proj_xr = pr.view_cos = 0.f; pr.level = 0; pr.pad = 0;
            pr.bad = fl & SDYN_MP_BAD ? 1 : 0; pr.obs_positive = fl & SDYN_MP_OBS_POSITIVE ? 1 : 0;
            bool in = !(fl & SDYN_MP_SKIP) && id >= 0 && id < tableCap;
            if (in) {
                const float* T = frame_part(poses, f, 96, P);
                const sdyn_map_point* p = table + id;
                const float X = p->world[0], Y = p->world[1], Z = p->world[2];
                /* Pc = mRcw*P + mtcw: 3x3 float product left to right, the addend joined in double (cv::gemm small-matrix path) */
                float pc[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float t = __fadd_rn(__fadd_rn(__fmul_rn(T[4 * k], X), __fmul_rn(T[4 * k + 1], Y)), __fmul_rn(T[4 * k + 2], Z));
                    pc[k] = (float)__dadd_rn((double)t, (double)T[4 * k + 3]);
                }
                if (pc[2] < 0.0f) in = false;
                const float invz = __fdiv_rn(1.0f, pc[2]);
                const float u = __fadd_rn(__fmul_rn(__fmul_rn(fp.fx, pc[0]), invz), fp.cx);
                const float v = __fadd_rn(__fmul_rn(__fmul_rn(fp.fy, pc[1]), invz), fp.cy);
                if (u < fp.minX || u > fp.maxX || v < fp.minY || v > fp.maxY) in = false;
                /* mOw = -mRcw.t()*mtcw (transposed operand: double accumulation) */
                float ow[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double a = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) a = __dadd_rn(a, __dmul_rn((double)T[4 * k + r], (double)T[4 * k + 3]));
                    ow[r] = (float)(-1.0 * a);
                }
                const float po0 = __fsub_rn(X, ow[0]), po1 = __fsub_rn(Y, ow[1]), po2 = __fsub_rn(Z, ow[2]);
                const double n2 = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)po0), __dmul_rn((double)po1, (double)po1)), __dmul_rn((double)po2, (double)po2));
                const float dist = (float)sqrt(n2);
                const float maxD = __fmul_rn(1.2f, p->max_distance), minD = __fmul_rn(0.8f, p->min_distance);
                if (dist < minD || dist > maxD) in = false;
                const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)p->normal[0]), __dmul_rn((double)po1, (double)p->normal[1])),
                                             __dmul_rn((double)po2, (double)p->normal[2]));
                const float viewCos = (float)__ddiv_rn(dot, (double)dist);
                if (viewCos < fp.cosLimit) in = false;
                if (in) {
                    const float ratio = __fdiv_rn(p->max_distance, dist);
                    int nScale = (int)ceilf(__fdiv_rn((float)log((double)ratio), fp.logScaleFactor));      /* MapPoint::PredictScale */
                    nScale = nScale < 0 ? 0 : (nScale >= fp.nlevels ? fp.nlevels - 1 : nScale);
                    pr.proj_x = u; pr.proj_y = v; pr.proj_xr = __fsub_rn(u, __fmul_rn(fp.bf, invz)); pr.view_cos = viewCos; pr.level = nScale;
                }
            }
            pr.track_in_view = in;
        }
        sdyn_mappoint_query o;
        o.proj_x = pr.proj_x; o.proj_y = pr.proj_y; o.proj_xr = pr.proj_xr; o.view_cos = pr.view_cos; o.level = pr.level;
        const bool has = id >= 0 && id < tableCap;
        o.track_in_view = has ? pr.track_in_view : 0; o.bad = pr.bad; o.obs_positive = pr.obs_positive; o.pad = 0;
        uint4 d0 = make_uint4(0, 0, 0, 0), d1 = d0;
        if (has) { d0 = *reinterpret_cast<const uint4*>(table[id].desc); d1 = *reinterpret_cast<const uint4*>(table[id].desc + 16); }
        uint32_t* w = reinterpret_cast<uint32_t*>(o.desc);
        w[0] = d0.x; w[1] = d0.y; w[2] = d0.z; w[3] = d0.w; w[4] = d1.x; w[5] = d1.y; w[6] = d1.z; w[7] = d1.w;
        gMap[at] = o;
    }
}

cudaError_t launch_gather_queries(const sdyn_map_point* table, int tableCap, const int32_t* lastIds, const uint8_t* lastFlags,
                                  const int32_t* nLast, int lastStride, sdyn_last_point* gLast, const int32_t* mapIds,
                                  const sdyn_map_proj* mapProj, const int32_t* nMap, int mapStride, sdyn_mappoint_query* gMap,
                                  const uint8_t* mapFlags, const float* poses, const FrustumParams& fp,
                                  int nframes, long long framePitch, cudaStream_t st)
{
    const int m = std::max(lastIds ? lastStride : 0, mapIds ? mapStride : 0);
    if (m == 0) return cudaSuccess;
    dim3 grid((m + 255) / 256, nframes, 2);
    k_gather_queries<<<grid, 256, 0, st>>>(table, tableCap, lastIds, lastFlags, nLast, lastStride, gLast, mapIds, mapProj, nMap, mapStride, gMap, mapFlags, poses, fp, framePitch);
    return cudaGetLastError();
}

}  // namespace sdyn
