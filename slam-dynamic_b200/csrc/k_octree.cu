/* Quad-tree keypoint selection (K3): one CTA per (frame, level), block-cooperative.
 *   reference: ORBextractor::DistributeOctTree + ExtractorNode::DivideNode, src/ORBextractor.cc:481-763
 *
 * The reference is a sequential std::list algorithm; its OUTPUT ORDER matters downstream (keypoint indices
 * feed the matchers' tie-breaks), so this kernel reproduces the list exactly, with data-parallel steps:
 *
 *  - A node is a row of a table held in LIST ORDER; a key only stores the list position of its node.
 *  - Every std::list insertion in the reference is a push_front, so list order = reverse creation order.
 *    One "round" (a breadth pass, or one sorted pass of the largest-first phase) therefore maps the old
 *    list to: [children of the LAST divided parent (n4..n1), ..., children of the FIRST divided parent,
 *    then all undivided nodes in their old order].  New positions are prefix sums over the processing order.
 *  - Breadth pass: every multi-key node is divided, processing order = list order.
 *  - Largest-first pass (entered when size + 3*nToExpand > N, :673): parents are processed by
 *    (size desc, creation desc); creation-desc = list position asc (the pointer tie-break of std::sort
 *    on pair<int,Node*> is pinned to creation order, SURVEY B-1).  Processing stops after the division
 *    that brings the list to >= N nodes (:730), i.e. parent r is divided iff size + sum_{r'<r} gain < N.
 *  - Key order inside a node only matters for "first maximum wins" in the final per-leaf selection
 *    (:744-760); that order is the FAST emission order = (cell row, cell col, y, x), so the best key of
 *    a leaf is argmax(response) with ties broken by the smallest such key — no stable partition needed.
 *
 * Small, latency-bound kernel (a few thousand keys, ~10 rounds); throughput comes from running
 * frames x levels CTAs concurrently.
 */
#include "sdyn_internal.h"

namespace sdyn {

constexpr int OT = 512;   /* threads per CTA */

struct Bounds { int16_t ulx, urx, uly, bry; };

__device__ __forceinline__ void split_point(const Bounds b, int& mx, int& my)
{
    /* halfX = ceil(float(UR.x-UL.x)/2) (DivideNode :483-484); extents are never negative */
    mx = b.ulx + ((b.urx - b.ulx + 1) >> 1);
    my = b.uly + ((b.bry - b.uly + 1) >> 1);
}

/* In-place exclusive scan of a[0..n) in shared memory by the whole CTA; returns the total. */
__device__ int block_exscan(int* a, int n, int* warpTmp)
{
    const int tid = threadIdx.x;
    const int per = (n + OT - 1) / OT;
    const int beg = min(tid * per, n), end = min(beg + per, n);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += a[i];
    /* scan of the per-thread sums */
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    __syncthreads();                       /* warpTmp may still be read from a previous call */
    if ((tid & 31) == 31) warpTmp[tid >> 5] = incl;
    __syncthreads();
    /* the 16 warp totals are scanned by the first warp (two shared reads per thread afterwards instead of 16) */
    if (tid < 32) {
        const int v = tid < OT / 32 ? warpTmp[tid] : 0;
        int wi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= o) wi += u;
        }
        if (tid < OT / 32) warpTmp[tid] = wi - v;
        if (tid == 31) warpTmp[OT / 32] = wi;
    }
    __syncthreads();
    const int base = warpTmp[tid >> 5], total = warpTmp[OT / 32];
    int run = base + incl - sum;
    for (int i = beg; i < end; ++i) { const int v = a[i]; a[i] = run; run += v; }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(OT)
k_octree(const __grid_constant__ Geom g, int iniTh, int minTh, const uint8_t* __restrict__ cellFlag,
         uint32_t* __restrict__ cand, const int32_t* __restrict__ candCount, int32_t* __restrict__ candNode,
         int32_t* __restrict__ selCount,
         LevelKp* __restrict__ levelKp, int32_t* __restrict__ levelCount, int32_t* __restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int level = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const LevelGeom& L = g.L[level];
    const int cap = g.maxNodeCap;
    const int N = L.quota;

    Bounds* bnd[2]; int* cnt[2];
    bnd[0] = reinterpret_cast<Bounds*>(smemRaw);
    bnd[1] = bnd[0] + cap;
    cnt[0] = reinterpret_cast<int*>(bnd[1] + cap);
    cnt[1] = cnt[0] + cap;
    int* cc = cnt[1] + cap;          /* 4 child counts per node */
    int* base = cc + 4 * cap;        /* per node: new position (undivided) or position of its first child slot */
    int* proc = base + cap;          /* per node: processing index among expandable nodes, or -1 */
    int* byProc = proc + cap;        /* processing index -> node */
    int* scanA = byProc + cap;       /* scan scratch */
    int* scanB = scanA + cap;
    __shared__ int warpTmp[OT / 32 + 1];
    __shared__ int sN, sExpand, sErr;

    uint32_t* K = cand + (size_t)f * g.candPerFrame + L.candOff;
    int32_t* node = candNode + (size_t)f * g.candPerFrame + L.candOff;
    const uint8_t* flags = cellFlag + (size_t)f * g.cellsPerFrame + L.cellOff;
    const int nRaw = min(candCount[f * SDYN_MAX_LEVELS + level], L.candCap);

    /* ---- per-cell threshold fallback (:809-816) and compaction of the surviving keys ---------------- */
    if (tid == 0) { sN = 0; sErr = 0; }
    __syncthreads();
    /* K is compacted in place through `node` as scratch: first mark, then move (two passes, so a slot is
     * never overwritten before it has been read). */
    for (int i0 = 0; i0 < nRaw; i0 += OT) {
        const int i = i0 + tid;
        uint32_t k = 0; bool ok = false;
        if (i < nRaw) {
            k = K[i];
            const int x = k & 4095, y = (k >> 12) & 4095, v = k >> 24;
            const int cell = ((y - 3) / L.hCell) * L.nCols + (x - 3) / L.wCell;
            ok = v >= iniTh || (!flags[cell] && v >= minTh);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        int wbase = 0;
        if ((tid & 31) == 0 && m) wbase = atomicAdd(&sN, __popc(m));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (ok) node[wbase + __popc(m & ((1u << (tid & 31)) - 1))] = (int32_t)k;
    }
    __syncthreads();
    const int n = sN;
    for (int i = tid; i < n; i += OT) K[i] = (uint32_t)node[i];
    __syncthreads();

    /* ---- roots (:543-585) ------------------------------------------------------------------------------ */
    const int nIni = L.nIni;
    const float hX = L.hX;
    for (int p = tid; p < cap; p += OT) { cnt[0][p] = 0; }
    for (int p = tid; p < nIni; p += OT) {
        Bounds b;
        b.ulx = (int16_t)(int)(hX * (float)p);
        b.urx = (int16_t)(int)(hX * (float)(p + 1));
        b.uly = 0; b.bry = (int16_t)L.fh;
        bnd[0][p] = b;
    }
    __syncthreads();
    for (int i = tid; i < n; i += OT) {
        const int x = K[i] & 4095;
        int r = (int)((float)x / hX);
        r = min(r, nIni - 1);
        node[i] = r;
        atomicAdd(&cnt[0][r], 1);
    }
    __syncthreads();
    /* drop empty roots, keep order */
    for (int p = tid; p < nIni; p += OT) scanA[p] = cnt[0][p] > 0;
    __syncthreads();
    int M = block_exscan(scanA, nIni, warpTmp);
    for (int p = tid; p < nIni; p += OT)
        if (cnt[0][p] > 0) { bnd[1][scanA[p]] = bnd[0][p]; cnt[1][scanA[p]] = cnt[0][p]; }
    __syncthreads();
    for (int i = tid; i < n; i += OT) node[i] = scanA[node[i]];
    __syncthreads();
    int cur = 1;

    /* ---- rounds ---------------------------------------------------------------------------------------- */
    bool sorted = false;
    for (int round = 0; round < 64; ++round) {
        const Bounds* B = bnd[cur]; const int* C = cnt[cur];
        Bounds* NB = bnd[cur ^ 1]; int* NC = cnt[cur ^ 1];

        /* 1. child key counts of every expandable node */
        for (int p = tid; p < M; p += OT) {
            cc[4 * p] = cc[4 * p + 1] = cc[4 * p + 2] = cc[4 * p + 3] = 0;
            scanA[p] = C[p] > 1;
        }
        __syncthreads();
        for (int i = tid; i < n; i += OT) {
            const int p = node[i];
            if (C[p] > 1) {
                const uint32_t k = K[i];
                int mx, my; split_point(B[p], mx, my);
                const int c = ((int)(k & 4095) >= mx) + 2 * ((int)((k >> 12) & 4095) >= my);
                atomicAdd(&cc[4 * p + c], 1);
            }
        }
        __syncthreads();

        /* 2. processing order of the expandable nodes */
        const int m = block_exscan(scanA, M, warpTmp);     /* scanA[p] = index among expandable, list order */
        if (!sorted) {
            for (int p = tid; p < M; p += OT) {
                const bool e = C[p] > 1;
                proc[p] = e ? scanA[p] : -1;
                if (e) byProc[scanA[p]] = p;
            }
        } else {
            for (int p = tid; p < M; p += OT) if (C[p] > 1) scanB[scanA[p]] = p;   /* compact list */
            __syncthreads();
            for (int k = tid; k < m; k += OT) {
                const int p = scanB[k], c = C[p];
                int rank = 0;
                for (int j = 0; j < m; ++j) {
                    const int cj = C[scanB[j]];
                    rank += (cj > c) || (cj == c && j < k);
                }
                byProc[rank] = p;
            }
            __syncthreads();
            for (int p = tid; p < M; p += OT) proc[p] = -1;
            __syncthreads();
            for (int r = tid; r < m; r += OT) proc[byProc[r]] = r;
        }
        __syncthreads();

        /* 3. which parents are divided, and where their children land */
        for (int r = tid; r < m; r += OT) {
            const int p = byProc[r];
            const int nch = (cc[4 * p] > 0) + (cc[4 * p + 1] > 0) + (cc[4 * p + 2] > 0) + (cc[4 * p + 3] > 0);
            scanA[r] = nch - 1;                               /* growth of the list when p is divided */
        }
        __syncthreads();
        block_exscan(scanA, m, warpTmp);                      /* scanA[r] = growth before p is processed */
        for (int r = tid; r < m; r += OT) {
            const int p = byProc[r];
            const bool divide = !sorted || (M + scanA[r] < N);
            const int nch = (cc[4 * p] > 0) + (cc[4 * p + 1] > 0) + (cc[4 * p + 2] > 0) + (cc[4 * p + 3] > 0);
            scanB[r] = divide ? nch : 0;
            if (!divide) proc[p] = -1;
        }
        __syncthreads();
        const int totalChildren = block_exscan(scanB, m, warpTmp);   /* scanB[r] = children created before p's */
        for (int p = tid; p < M; p += OT) scanA[p] = proc[p] < 0;     /* undivided nodes keep relative order */
        __syncthreads();
        const int nKeep = block_exscan(scanA, M, warpTmp);
        const int newM = totalChildren + nKeep;
        if (newM > cap) { if (tid == 0) sErr = 1; __syncthreads(); break; }   /* cannot happen; guards smem */

        /* 4. build the new table */
        if (tid == 0) sExpand = 0;
        __syncthreads();
        int myExpand = 0;
        for (int p = tid; p < M; p += OT) {
            const int r = proc[p];
            if (r < 0) {
                const int q = totalChildren + scanA[p];
                NB[q] = B[p]; NC[q] = C[p];
                base[p] = q;
            } else {
                const int nch = (cc[4 * p] > 0) + (cc[4 * p + 1] > 0) + (cc[4 * p + 2] > 0) + (cc[4 * p + 3] > 0);
                /* children created later sit further to the front: this parent's block starts after the
                 * blocks of all parents processed after it */
                const int first = totalChildren - scanB[r] - nch;
                base[p] = first;
                const Bounds b = B[p];
                int mx, my; split_point(b, mx, my);
                int q = first;
                for (int c = 3; c >= 0; --c) {               /* n4, n3, n2, n1 from the front */
                    const int k = cc[4 * p + c];
                    if (k == 0) continue;
                    Bounds nb;
                    nb.ulx = (c & 1) ? (int16_t)mx : b.ulx;  nb.urx = (c & 1) ? b.urx : (int16_t)mx;
                    nb.uly = (c & 2) ? (int16_t)my : b.uly;  nb.bry = (c & 2) ? b.bry : (int16_t)my;
                    NB[q] = nb; NC[q] = k;
                    myExpand += k > 1;
                    ++q;
                }
            }
        }
        if (myExpand) atomicAdd(&sExpand, myExpand);
        __syncthreads();

        /* 5. move the keys */
        for (int i = tid; i < n; i += OT) {
            const int p = node[i];
            if (proc[p] < 0) { node[i] = base[p]; continue; }
            const uint32_t k = K[i];
            int mx, my; split_point(B[p], mx, my);
            const int c = ((int)(k & 4095) >= mx) + 2 * ((int)((k >> 12) & 4095) >= my);
            int q = base[p];
            for (int c2 = 3; c2 > c; --c2) q += cc[4 * p + c2] > 0;
            node[i] = q;
        }
        __syncthreads();

        const int prevM = M;
        const int nToExpand = sExpand;
        M = newM;
        cur ^= 1;
        /* :664-673, :735-736 */
        if (M >= N || M == prevM) break;
        if (!sorted && M + 3 * nToExpand > N) sorted = true;
    }

    /* ---- keep the best key of every leaf (:744-760) ------------------------------------------------- */
    int* bestV = scanA; int* bestKey = scanB;
    for (int p = tid; p < M; p += OT) { bestV[p] = -1; bestKey[p] = 0x7fffffff; }
    __syncthreads();
    for (int i = tid; i < n; i += OT) atomicMax(&bestV[node[i]], (int)(K[i] >> 24));
    __syncthreads();
    for (int i = tid; i < n; i += OT) {
        const uint32_t k = K[i];
        const int p = node[i];
        if ((int)(k >> 24) == bestV[p]) {
            const int x = k & 4095, y = (k >> 12) & 4095;
            const int cx = (x - 3) / L.wCell, cy = (y - 3) / L.hCell;
            const int key = ((cy * L.nCols + cx) << 12) | ((y - 3 - cy * L.hCell) << 6) | (x - 3 - cx * L.wCell);
            atomicMin(&bestKey[p], key);
        }
    }
    __syncthreads();
    LevelKp* out = levelKp + (size_t)f * g.kpPerFrame + L.kpOff;
    for (int p = tid; p < M; p += OT) {
        const int key = bestKey[p];
        const int cell = key >> 12, cy = cell / L.nCols, cx = cell - cy * L.nCols;
        LevelKp o;
        o.x = (int16_t)(cx * L.wCell + (key & 63) + 3 + kFastBorder);
        o.y = (int16_t)(cy * L.hCell + ((key >> 6) & 63) + 3 + kFastBorder);
        o.score = bestV[p];
        out[p] = o;
    }
    if (tid == 0) {
        levelCount[f * SDYN_MAX_LEVELS + level] = M;
        selCount[f * SDYN_MAX_LEVELS + level] = n;
        if (sErr || candCount[f * SDYN_MAX_LEVELS + level] > L.candCap) atomicOr(&status[f], 1);
    }
}

size_t octree_smem_bytes(int nodeCap)
{
    return (size_t)nodeCap * (2 * sizeof(Bounds) + 2 * sizeof(int) + 4 * sizeof(int) + 5 * sizeof(int)) + 64;
}

cudaError_t launch_octree(const Geom& g, int iniTh, int minTh, const uint8_t* dCellFlag, uint32_t* dCand,
                          const int32_t* dCandCount, int32_t* dCandNode, int32_t* dSelCount, LevelKp* dLevelKp,
                          int32_t* dLevelCount, int32_t* dStatus, int nframes, cudaStream_t st)
{
    const size_t smem = octree_smem_bytes(g.maxNodeCap);
    cudaError_t e = cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid(g.nlevels, nframes);
    k_octree<<<grid, OT, smem, st>>>(g, iniTh, minTh, dCellFlag, dCand, dCandCount, dCandNode, dSelCount, dLevelKp,
                                     dLevelCount, dStatus);
    return cudaGetLastError();
}

}  // namespace sdyn
