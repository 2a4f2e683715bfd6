/* Bag-of-words transform (K11): Frame::ComputeBoW -> DBoW2 TemplatedVocabulary::transform, src/Frame.cc:803-810,
 * Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1263 (the step in front of SearchByBoW).
 *
 * The vocabulary tree lives in HBM once per process (ORBvoc: k = 10, L = 6, ~1.1 M nodes x 32-byte descriptors =
 * 35 MB, L2-resident on B200).  k_bow_descend walks one descriptor per thread down the tree: per level the
 * children's descriptors are independent 32-byte loads and 8 __popc each; the first child of smallest distance wins,
 * exactly as the reference's strict `d < best_d` over the children in order.  It emits per feature the word id, the
 * word's weight and the node at level L - levelsup; the std::map containers of the reference (BowVector with L1
 * normalisation, FeatureVector) are assembled from those arrays on the host (sdyn_bow_assemble) because they ARE
 * host containers in the drop-in. */
#include "sdyn_internal.h"
#include "bow_internal.h"

namespace sdyn {

__global__ void __launch_bounds__(128)
k_bow_descend(VocabView v, const uint8_t* __restrict__ desc, const int32_t* __restrict__ count, int cap, int levelsup,
              uint32_t* __restrict__ wordId, double* __restrict__ weight, uint32_t* __restrict__ nodeId)
{
    const int f = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    const int n = count ? min(count[f], cap) : cap;
    if (i >= n) return;
    const size_t o = (size_t)f * cap + i;
    uint32_t q[8];
    {
        const uint4 a = reinterpret_cast<const uint4*>(desc + o * 32)[0], b = reinterpret_cast<const uint4*>(desc + o * 32)[1];
        q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w; q[4] = b.x; q[5] = b.y; q[6] = b.z; q[7] = b.w;
    }
    const int nidLevel = v.L - levelsup;
    uint32_t nid = 0, cur = 0;                       /* nid_level <= 0 -> root */
    int level = 0;
    do {
        ++level;
        const int b = v.childOff[cur], e = v.childOff[cur + 1];
        uint32_t bestKey = 0xffffffffu;              /* distance << 16 | position among the children: first minimum */
        for (int c0 = b; c0 < e; c0 += 4) {
            uint32_t key[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = min(c0 + k, e - 1);
                const uint4* d = reinterpret_cast<const uint4*>(v.desc + (size_t)v.childIdx[c] * 32);
                const uint4 x = __ldg(d), y = __ldg(d + 1);
                const int dist = __popc(q[0] ^ x.x) + __popc(q[1] ^ x.y) + __popc(q[2] ^ x.z) + __popc(q[3] ^ x.w) +
                                 __popc(q[4] ^ y.x) + __popc(q[5] ^ y.y) + __popc(q[6] ^ y.z) + __popc(q[7] ^ y.w);
                key[k] = c0 + k < e ? ((uint32_t)dist << 16) | (uint32_t)(c0 + k - b) : 0xffffffffu;
            }
            bestKey = min(min(bestKey, min(key[0], key[1])), min(key[2], key[3]));
        }
        cur = v.childIdx[b + (bestKey & 0xffff)];
        if (level == nidLevel) nid = cur;
    } while (v.childOff[cur + 1] > v.childOff[cur]);
    wordId[o] = v.wordOf[cur];
    weight[o] = v.weight[cur];
    nodeId[o] = nid;
}

cudaError_t launch_bow_descend(const VocabView& v, const uint8_t* dDesc, const int32_t* dCount, int cap, int nframes,
                               int levelsup, uint32_t* dWord, double* dWeight, uint32_t* dNode, cudaStream_t st)
{
    if (cap <= 0 || nframes <= 0) return cudaSuccess;
    dim3 grid((cap + 127) / 128, nframes);
    k_bow_descend<<<grid, 128, 0, st>>>(v, dDesc, dCount, cap, levelsup, dWord, dWeight, dNode);
    return cudaGetLastError();
}

}  // namespace sdyn
