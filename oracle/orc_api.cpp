/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * Flat C entry points so tests/ and bench.py's cpu_baseline leg can drive the oracle through ctypes. */
#include "orc_prims.h"
#include "orc_extractor.h"
#include <cstring>
#include <algorithm>

using namespace orc;

extern "C" {

void orc_resize_linear_u8(const uint8_t* s, int sw, int sh, int ss, uint8_t* d, int dw, int dh, int ds)
{ resize_linear_u8(s, sw, sh, ss, d, dw, dh, ds); }

void orc_border_reflect101(const uint8_t* s, int w, int h, int ss, uint8_t* d, int b, int ds)
{ border_reflect101(s, w, h, ss, d, b, ds, false); }

void orc_gaussian_blur7(const uint8_t* s, int w, int h, int ss, uint8_t* d, int ds)
{ gaussian_blur7_s2(s, w, h, ss, d, ds); }

int orc_fast_nms(const uint8_t* img, int w, int h, int stride, int th, int* xyv, int cap)
{ return fast_nms(img, w, h, stride, th, xyv, cap); }

/* score map over the interior (border rows/cols left 0), for stage-level parity tests */
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, int16_t* out)
{
    std::memset(out, 0, sizeof(int16_t) * (size_t)w * h);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) out[(size_t)y * w + x] = (int16_t)fast_score(img + (size_t)y * stride + x, stride);
}

void orc_fast_atan2(const float* y, const float* x, float* out, int n)
{ for (int i = 0; i < n; ++i) out[i] = fast_atan2(y[i], x[i]); }

void* orc_extractor_create(int nf, float sf, int nl, int ini, int mn) { return new Extractor(nf, sf, nl, ini, mn); }
void orc_extractor_destroy(void* e) { delete (Extractor*)e; }

int orc_extractor_run(void* e, const uint8_t* img, int w, int h, int stride, KeyPoint* kps, uint8_t* desc, int cap)
{
    Extractor* E = (Extractor*)e;
    std::vector<KeyPoint> k; std::vector<uint8_t> d;
    int n = E->run(img, w, h, stride, k, d);
    if (n < 0) return n;
    int m = std::min(n, cap);
    if (m > 0) { std::memcpy(kps, k.data(), sizeof(KeyPoint) * (size_t)m); std::memcpy(desc, d.data(), (size_t)32 * m); }
    return n;
}

void orc_extractor_tables(void* e, float* scale, float* inv, float* sig, float* invsig, int* quota, int* umax16)
{
    Extractor* E = (Extractor*)e;
    for (int i = 0; i < E->nlevels; ++i) { scale[i] = E->scale[i]; inv[i] = E->invScale[i]; sig[i] = E->sigma2[i]; invsig[i] = E->invSigma2[i]; quota[i] = E->quota[i]; }
    for (int i = 0; i < 16; ++i) umax16[i] = E->umax[i];
}

int orc_extractor_level_dims(void* e, int l, int* w, int* h)
{ Extractor* E = (Extractor*)e; if (l < 0 || l >= E->nlevels) return -1; *w = E->pyr[l].w; *h = E->pyr[l].h; return 0; }

/* copies the bordered level buffer, (w+38) x (h+38) */
int orc_extractor_level_copy(void* e, int l, uint8_t* out)
{ Extractor* E = (Extractor*)e; if (l < 0 || l >= E->nlevels) return -1; std::memcpy(out, E->pyr[l].buf.data(), E->pyr[l].buf.size()); return 0; }

int orc_extractor_candidates(void* e, int l, int* xyv, int cap)
{
    Extractor* E = (Extractor*)e; if (l < 0 || l >= (int)E->cand.size()) return -1;
    int n = (int)E->cand[l].size() / 3;
    std::memcpy(xyv, E->cand[l].data(), sizeof(int) * 3 * (size_t)std::min(n, cap));
    return n;
}

int orc_extractor_level_count(void* e, int l) { Extractor* E = (Extractor*)e; return (l < 0 || l >= (int)E->perLevel.size()) ? -1 : E->perLevel[l]; }

float orc_ic_angle(const uint8_t* center, int stride)
{ static Extractor E(1000, 1.2f, 8, 20, 7); return ic_angle(center, stride, E.umax); }

void orc_orb_descriptor(float angle, const uint8_t* center, int stride, uint8_t* out) { orb_descriptor(angle, center, stride, out); }

}  // extern "C"
