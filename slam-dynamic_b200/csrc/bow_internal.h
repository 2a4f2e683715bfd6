/* Device view of a DBoW2 vocabulary tree (k_bow.cu) — shared with the host glue. */
#pragma once
#include "sdyn_internal.h"

namespace sdyn {

struct VocabView {
    int nnodes, L;
    const int32_t* childOff;     /* nnodes + 1: children of node i are childIdx[childOff[i] .. childOff[i+1]) in DBoW2 order */
    const uint32_t* childIdx;
    const uint8_t* desc;         /* nnodes x 32 */
    const double* weight;        /* nnodes */
    const uint32_t* wordOf;      /* nnodes: word id of a leaf, 0 otherwise (Node::word_id default) */
};

cudaError_t launch_bow_descend(const VocabView& v, const uint8_t* dDesc, const int32_t* dCount, int cap, int nframes,
                               int levelsup, uint32_t* dWord, double* dWeight, uint32_t* dNode, cudaStream_t st);

}  // namespace sdyn
