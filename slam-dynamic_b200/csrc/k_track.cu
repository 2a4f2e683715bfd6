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

namespace sdyn {

__global__ void __launch_bounds__(256)
k_box_occupancy(const uint64_t* __restrict__ mask, const int32_t* __restrict__ count, int cap,
                unsigned long long* __restrict__ has)
{
    const int f = blockIdx.x, n = min(count[f], cap);
    unsigned long long m = 0;
    for (int i = threadIdx.x; i < n; i += 256) m |= mask[(size_t)f * cap + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    __shared__ unsigned long long w[8];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) { for (int k = 1; k < 8; ++k) m |= w[k]; has[f] = m; }
}

__device__ __forceinline__ int hamming_bytes2(const uint8_t* a, const uint8_t* b)
{
    const uint4 a0 = *reinterpret_cast<const uint4*>(a), a1 = *reinterpret_cast<const uint4*>(a + 16);
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 16);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

constexpr int BS = 128;

/* one CTA per (box slot, frame) */
__global__ void __launch_bounds__(BS)
k_box_stage(const sdyn_track_inputs in, const sdyn_keypoint* __restrict__ kp, const sdyn_keypoint* __restrict__ kpUn,
            const uint8_t* __restrict__ desc,
            const int32_t* __restrict__ count, int cap, const uint64_t* __restrict__ mask,
            const unsigned long long* __restrict__ has, int32_t* __restrict__ boxList, int32_t* __restrict__ nnQ,
            int32_t* __restrict__ nnT, int nnTStride, uint8_t* __restrict__ readmit, int32_t* __restrict__ staticExit)
{
    const int s = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int nb = min(in.n_boxes[f], 64);
    if (s >= nb) return;
    __shared__ int sSurv, sOcc, sCount, sBase, warpCnt[BS / 32], sStatic;
    if (tid == 0) {
        /* replay of the erase loop (Frame.cc:585-592): `i` advances after an erase, hasKpts keeps its indexing */
        const unsigned long long h = has[f];
        int ids[64], size = nb;
        for (int i = 0; i < nb; ++i) ids[i] = i;
        for (int i = 0; i < size; ++i) {
            if ((h >> i) & 1ull) continue;
            for (int k = i; k + 1 < size; ++k) ids[k] = ids[k + 1];
            --size;
        }
        sSurv = s < size ? ids[s] : -1;
        int occ = -1, seen = 0;
        for (int b = 0; b < nb; ++b) if ((h >> b) & 1ull) { if (seen == s) { occ = b; break; } ++seen; }
        sOcc = occ; sCount = 0; sBase = 0; sStatic = 0;
    }
    __syncthreads();
    const int surv = sSurv, occ = sOcc;
    if (surv < 0 || occ < 0) return;                 /* slot beyond objects.size(), or no keypoints in it */
    const int r = in.ref_box[f * 64 + surv];
    if (r < 0) return;                               /* box id not present in the reference frame */
    const int n = min(count[f], cap);
    /* classifyF reads mvdynKeysUn (Tracking.cc:1129-1131: undistorted points); the box test above used mvKeys (Frame.cc:562) */
    const sdyn_keypoint* K = kpUn + (size_t)f * cap;
    const uint8_t* D = desc + (size_t)f * cap * 32;
    const uint64_t* M = mask + (size_t)f * cap;
    int32_t* list = boxList + ((size_t)f * 64 + s) * cap;
    int32_t* nq_ = nnQ + ((size_t)f * 64 + s) * cap;
    int32_t* nt_ = nnT + ((size_t)f * 64 + s) * nnTStride;

    /* keypoints of box `occ`, ascending index = order of mvdynKeys[slot] */
    for (int i0 = 0; i0 < n; i0 += BS) {
        const int i = i0 + tid;
        const bool ok = i < n && ((M[i] >> occ) & 1ull);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if ((tid & 31) == 0) warpCnt[tid >> 5] = __popc(bal);
        __syncthreads();
        int pos = sBase;
        for (int w = 0; w < (tid >> 5); ++w) pos += warpCnt[w];
        if (ok) list[pos + __popc(bal & ((1u << (tid & 31)) - 1))] = i;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < BS / 32; ++w) t += warpCnt[w]; sBase += t; }
        __syncthreads();
    }
    const int nq = sBase;
    const int to = in.ref_off[f * 65 + r], nt = min(in.ref_off[f * 65 + r + 1] - to, nnTStride);
    if (nq == 0 || nt == 0) return;                  /* mdynDescriptors[..].cols == 0 */
    const uint8_t* TD = in.ref_desc + ((size_t)f * in.ref_stride + to) * 32;
    const float* TX = in.ref_xy + ((size_t)f * in.ref_stride + to) * 2;

    /* BFMatcher(NORM_HAMMING, crossCheck = true) */
    for (int i = tid; i < nq; i += BS) {
        int best = 1 << 30, bj = -1;
        const uint8_t* q = D + 32 * (size_t)list[i];
        for (int j = 0; j < nt; ++j) { const int d = hamming_bytes2(q, TD + 32 * (size_t)j); if (d < best) { best = d; bj = j; } }
        nq_[i] = bj;
    }
    for (int j = tid; j < nt; j += BS) {
        int best = 1 << 30, bi = -1;
        const uint8_t* t = TD + 32 * (size_t)j;
        for (int i = 0; i < nq; ++i) { const int d = hamming_bytes2(D + 32 * (size_t)list[i], t); if (d < best) { best = d; bi = i; } }
        nt_[j] = bi;
    }
    __syncthreads();

    /* classifyF on the mutual matches; num0 = #static */
    float m[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) m[k] = in.fmat[f * 9 + k];
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
        if (nq_[i] <= -2) readmit[(size_t)f * cap + list[i]] = 1;       /* dynStatus[slot][m] != -1 */
    if (tid == 0 && (double)num0 > fmax(1.0, 0.2 * (double)good)) atomicOr(&staticExit[f], 1);
}

__global__ void __launch_bounds__(256)
k_dyn_finalize(const uint64_t* __restrict__ mask, const uint8_t* __restrict__ readmit,
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
                             int cap, uint64_t* mask, unsigned long long* has, int32_t* boxList, int32_t* nnQ, int32_t* nnT,
                             int nnTStride, uint8_t* readmit, int32_t* staticExit, uint8_t* dynMask, int32_t* counts,
                             int nframes, cudaStream_t st)
{
    cudaError_t e = launch_box_mask(kp, count, cap, cap, in.boxes, in.n_boxes, 64, 64, mask, nframes, st);
    if (e != cudaSuccess) return e;
    k_box_occupancy<<<nframes, 256, 0, st>>>(mask, count, cap, has);
    dim3 grid(64, nframes);
    k_box_stage<<<grid, BS, 0, st>>>(in, kp, kpUn, desc, count, cap, mask, has, boxList, nnQ, nnT, nnTStride, readmit, staticExit);
    k_dyn_finalize<<<nframes, 256, 0, st>>>(mask, readmit, staticExit, count, cap, dynMask, counts);
    return cudaGetLastError();
}

}  // namespace sdyn
