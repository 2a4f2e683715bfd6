/* Device-side argument block of the stereo association kernels (k_stereo.cu) — shared with the host glue. */
#pragma once
#include "sdyn_internal.h"

namespace sdyn {

struct StereoArgs {
    /* extraction results of the left / right context (device) */
    const sdyn_keypoint* keysL; const uint8_t* descL; const int32_t* countL; int capL;
    const sdyn_keypoint* keysR; const uint8_t* descR; const int32_t* countR; int capR;
    const uint8_t* pyrL; const uint8_t* pyrR; size_t frameBytesL, frameBytesR;
    int nRows, maxRows;                  /* mvImagePyramid[0].rows; row-table stride - 1 */
    float scale[SDYN_MAX_LEVELS], invScale[SDYN_MAX_LEVELS];
    float mb, mbf;
    /* scratch */
    int32_t* rowOff;                     /* [B][maxRows + 1] */
    int32_t* rowList; int listCap;       /* [B][listCap] right keypoint indices, grouped by row */
    int2* rightKey;                      /* [B][capR]: (bits of pt.x, octave | minr << 8 | maxr << 20) */
    int32_t* sad;                        /* [B][capL] best window distance, -1 = no stereo match */
    int32_t* status;                     /* [B] 1 = a right keypoint's row span left the image (reference: out-of-range write) */
    /* outputs */
    float* uRight; float* depth;         /* [B][capL]  mvuRight, mvDepth */
    int32_t* kept;                       /* [B] stereo points after the median cut */
};

cudaError_t launch_stereo(const Geom& g, const StereoArgs& a, int nframes, int maxKpL, cudaStream_t st);

}  // namespace sdyn
