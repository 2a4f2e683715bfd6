/* Seeded, integer-only synthetic frame / box / map generator.
 *
 * There are no datasets in the build or GPU containers, so benchmarks and parity tests run on frames
 * produced here.  Everything is integer arithmetic on 64-bit hashes, so the bytes are identical on any
 * host.  Content lives in an unbounded "world" plane; a frame is a WxH window at an integer offset, so
 * consecutive frames of a sequence overlap exactly (real matches exist) while per-frame sensor noise
 * and the pyramid resampling phase still differ.
 *
 * Plain C, CPU only; part of the benchmark harness, not of the reference-facing API.
 */
#include "sdyn_synth.h"
#include <stdlib.h>
#include <string.h>

static inline uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

static inline uint64_t hash4(uint64_t seed, uint64_t tag, int64_t a, int64_t b)
{
    uint64_t h = mix64(seed ^ (tag * 0xD6E8FEB86659FD93ull));
    h = mix64(h ^ (uint64_t)a);
    h = mix64(h ^ ((uint64_t)b * 0xCA5A826395121157ull));
    return h;
}

/* floor division for possibly negative world coordinates */
static inline int64_t fdiv(int64_t a, int64_t c) { int64_t q = a / c; return (a % c != 0 && (a < 0)) ? q - 1 : q; }

/* bilinear value noise on a c-pixel lattice, 0..255, evaluated for a whole WxH window at (ox,oy);
 * the lattice values covering the window are hashed once and cached */
static void add_value_noise(uint64_t seed, uint64_t tag, int c, int weight, int W, int H, int ox, int oy, int* acc)
{
    int64_t ix0 = fdiv(ox, c), iy0 = fdiv(oy, c);
    int gw = (int)(fdiv((int64_t)ox + W - 1, c) - ix0) + 2, gh = (int)(fdiv((int64_t)oy + H - 1, c) - iy0) + 2;
    int* g = (int*)malloc(sizeof(int) * (size_t)gw * gh);
    for (int j = 0; j < gh; ++j)
        for (int i = 0; i < gw; ++i) g[(size_t)j * gw + i] = (int)(hash4(seed, tag, ix0 + i, iy0 + j) >> 56);
    for (int y = 0; y < H; ++y) {
        int64_t wy = (int64_t)y + oy;
        int64_t iy = fdiv(wy, c);
        int fy = (int)(wy - iy * c);
        const int* r0 = g + (size_t)(iy - iy0) * gw;
        const int* r1 = r0 + gw;
        for (int x = 0; x < W; ++x) {
            int64_t wx = (int64_t)x + ox;
            int64_t ix = fdiv(wx, c);
            int fx = (int)(wx - ix * c);
            int i = (int)(ix - ix0);
            int top = r0[i] * (c - fx) + r0[i + 1] * fx, bot = r1[i] * (c - fx) + r1[i + 1] * fx;
            acc[(size_t)y * W + x] += weight * ((top * (c - fy) + bot * fy) / (c * c));
        }
    }
    free(g);
}

typedef struct { int64_t x0, y0; int w, h, gray, alpha, vx, vy, moving; } rect_t;

static rect_t make_rect(uint64_t seed, int r, int W, int H, int t)
{
    rect_t q;
    uint64_t a = hash4(seed, 100, r, 0), b = hash4(seed, 101, r, 0), c = hash4(seed, 102, r, 0);
    int wmax = W / 6 > 9 ? W / 6 : 9, hmax = H / 4 > 9 ? H / 4 : 9;
    q.x0 = (int64_t)(a % (uint64_t)(W + 64)) - 32;
    q.y0 = (int64_t)((a >> 32) % (uint64_t)(H + 64)) - 32;
    q.w = 8 + (int)(b % (uint64_t)(wmax - 8));
    q.h = 8 + (int)((b >> 32) % (uint64_t)(hmax - 8));
    q.gray = (int)(c >> 56);
    q.alpha = 64 + (int)((c >> 20) % 192);
    q.moving = (r % 16) == 5;                    /* every 16th rectangle is an independently moving object */
    q.vx = q.moving ? (int)((c >> 8) % 13) - 6 : 0;
    q.vy = q.moving ? (int)((c >> 12) % 7) - 3 : 0;
    q.x0 += (int64_t)q.vx * t;
    q.y0 += (int64_t)q.vy * t;
    return q;
}

void sdyn_synth_frame(uint64_t seq_seed, uint64_t frame_seed, int W, int H, int nrect,
                      int ox, int oy, int t, uint8_t* out, int stride)
{
    /* 1. three octaves of value noise: (4*V64 + 3*V16 + V4) / 8 */
    int* acc = (int*)calloc((size_t)W * H, sizeof(int));
    add_value_noise(seq_seed, 1, 64, 4, W, H, ox, oy, acc);
    add_value_noise(seq_seed, 2, 16, 3, W, H, ox, oy, acc);
    add_value_noise(seq_seed, 3, 4, 1, W, H, ox, oy, acc);
    for (int y = 0; y < H; ++y) {
        uint8_t* row = out + (size_t)y * stride;
        for (int x = 0; x < W; ++x) row[x] = (uint8_t)(acc[(size_t)y * W + x] / 8);
    }
    free(acc);
    /* 2. alpha-blended axis-aligned rectangles, painted in index order; static ones are fixed in the
     *    world, moving ones additionally drift by (vx,vy) per time step t */
    for (int r = 0; r < nrect; ++r) {
        rect_t q = make_rect(seq_seed, r, W, H, t);
        int64_t fx0 = q.x0 - ox, fy0 = q.y0 - oy;
        int x0 = (int)(fx0 < 0 ? 0 : fx0), y0 = (int)(fy0 < 0 ? 0 : fy0);
        int64_t x1l = fx0 + q.w, y1l = fy0 + q.h;
        int x1 = (int)(x1l > W ? W : x1l), y1 = (int)(y1l > H ? H : y1l);
        for (int y = y0; y < y1; ++y) {
            uint8_t* row = out + (size_t)y * stride;
            for (int x = x0; x < x1; ++x) row[x] = (uint8_t)((row[x] * (256 - q.alpha) + q.gray * q.alpha) >> 8);
        }
    }
    /* 3. per-frame sensor noise in [-3,4], clipped */
    for (int y = 0; y < H; ++y) {
        uint8_t* row = out + (size_t)y * stride;
        for (int x = 0; x < W; ++x) {
            int v = row[x] + (int)(hash4(frame_seed, 7, x, y) >> 61) - 3;
            row[x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

int sdyn_synth_boxes_ids(uint64_t seq_seed, int W, int H, int nrect, int ox, int oy, int t, int margin,
                         double* xywh, int* ids, int cap);

int sdyn_synth_boxes(uint64_t seq_seed, int W, int H, int nrect, int ox, int oy, int t, int margin,
                     double* xywh, int cap)
{
    return sdyn_synth_boxes_ids(seq_seed, W, H, nrect, ox, oy, t, margin, xywh, 0, cap);
}

int sdyn_synth_boxes_ids(uint64_t seq_seed, int W, int H, int nrect, int ox, int oy, int t, int margin,
                         double* xywh, int* ids, int cap)
{
    int n = 0;
    for (int r = 0; r < nrect; ++r) {
        rect_t q = make_rect(seq_seed, r, W, H, t);
        if (!q.moving) continue;
        /* YOLO-style box file convention (reference: Examples/RGB-D/rgbd_my.cc:232-252):
         * id cx cy w h  ->  Rect2d(max(cx-w/2,0), max(cy-h/2,0), w, h) */
        double bw = q.w + 2.0 * margin, bh = q.h + 2.0 * margin;
        double cx = (double)(q.x0 - ox) + q.w / 2.0, cy = (double)(q.y0 - oy) + q.h / 2.0;
        if (cx < 0 || cy < 0 || cx >= W || cy >= H) continue;
        if (n < cap) {
            double bx = cx - bw / 2, by = cy - bh / 2;
            xywh[4 * n] = bx > 0 ? bx : 0;
            xywh[4 * n + 1] = by > 0 ? by : 0;
            xywh[4 * n + 2] = bw;
            xywh[4 * n + 3] = bh;
            if (ids) ids[n] = r;
        }
        ++n;
    }
    return n < cap ? n : cap;
}

uint64_t sdyn_synth_hash(uint64_t seed, uint64_t tag, int64_t a, int64_t b) { return hash4(seed, tag, a, b); }
