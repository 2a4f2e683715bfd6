/* Seeded integer-only synthetic inputs for benchmarks and parity tests. */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Renders the WxH window at world offset (ox,oy) of sequence `seq_seed` at time step t.
 * `frame_seed` drives the per-frame sensor noise.  nrect rectangles; every 16th one moves on its own. */
void sdyn_synth_frame(uint64_t seq_seed, uint64_t frame_seed, int W, int H, int nrect,
                      int ox, int oy, int t, uint8_t* out, int stride);

/* Detection boxes (x, y, w, h doubles, cv::Rect2d layout) around the moving rectangles of the same
 * sequence at time t, grown by `margin` pixels; returns the count written (<= cap). */
int sdyn_synth_boxes(uint64_t seq_seed, int W, int H, int nrect, int ox, int oy, int t, int margin,
                     double* xywh, int cap);

/* Same, also returning the rectangle id of every box (stable across frames: the tracking identity). */
int sdyn_synth_boxes_ids(uint64_t seq_seed, int W, int H, int nrect, int ox, int oy, int t, int margin,
                         double* xywh, int* ids, int cap);

uint64_t sdyn_synth_hash(uint64_t seed, uint64_t tag, int64_t a, int64_t b);

#ifdef __cplusplus
}
#endif
