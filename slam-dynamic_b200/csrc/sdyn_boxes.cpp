/* Host-side edges of the dynamic-keypoint path: the YOLO detection file of a frame and the box
 * association between consecutive frames.  Tens of boxes per frame — plain host code in the C ABI, no device work.
 *   reference: Examples/RGB-D/rgbd_my.cc:232-252 (file format "id cx cy w h", one detection per line)
 *              Frame::boxTrack, src/Frame.cc:481-552 (cv::Rect2d &, |, area, contains) */
#include "../../include/sdyn.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Rect { double x, y, w, h; };

inline bool is_empty(const Rect& r) { return r.w <= 0 || r.h <= 0; }

/* cv::Rect_<double>::operator& : an empty intersection is the all-zero rectangle */
inline Rect intersect(const Rect& a, const Rect& b)
{
    const double x1 = std::max(a.x, b.x), y1 = std::max(a.y, b.y);
    Rect r{x1, y1, std::min(a.x + a.w, b.x + b.w) - x1, std::min(a.y + a.h, b.y + b.h) - y1};
    if (is_empty(r)) r = Rect{0, 0, 0, 0};
    return r;
}

/* cv::Rect_<double>::operator| : bounding rectangle; an empty operand yields the other one */
inline Rect bounding(const Rect& a, const Rect& b)
{
    if (is_empty(a)) return b;
    if (is_empty(b)) return a;
    const double x1 = std::min(a.x, b.x), y1 = std::min(a.y, b.y);
    return Rect{x1, y1, std::max(a.x + a.w, b.x + b.w) - x1, std::max(a.y + a.h, b.y + b.h) - y1};
}

}  // namespace

extern "C" {

int sdyn_boxes_parse(const char* text, size_t len, double* xywh, int cap, int* nOut)
{
    if (!nOut || (len > 0 && !text) || cap < 0 || (cap > 0 && !xywh)) return SDYN_ERR_ARG;
    int n = 0;
    size_t pos = 0;
    std::string line;
    while (pos < len) {
        size_t end = pos;
        while (end < len && text[end] != '\n') ++end;
        line.assign(text + pos, end - pos);
        pos = end + 1;
        if (line.empty()) continue;                           /* rgbd_my.cc:242 */
        double v[5];
        const char* p = line.c_str();
        int got = 0;
        for (; got < 5; ++got) {
            char* q = nullptr;
            v[got] = std::strtod(p, &q);
            if (q == p) break;
            p = q;
        }
        if (got < 5) continue;                                /* the reference reads uninitialised values here */
        if (n < cap) {
            /* cv::Rect2d(MAX(cx - w/2, 0), MAX(cy - h/2, 0), w, h), rgbd_my.cc:248 */
            const double x = v[1] - v[3] / 2, y = v[2] - v[4] / 2;
            xywh[4 * n] = x > 0 ? x : 0; xywh[4 * n + 1] = y > 0 ? y : 0; xywh[4 * n + 2] = v[3]; xywh[4 * n + 3] = v[4];
        }
        ++n;
    }
    *nOut = n;
    return n > cap ? SDYN_ERR_CAPACITY : SDYN_OK;
}

int sdyn_boxes_read(const char* path, double* xywh, int cap, int* nOut)
{
    if (!path || !nOut) return SDYN_ERR_ARG;
    *nOut = 0;
    FILE* f = std::fopen(path, "rb");
    if (!f) return SDYN_OK;                                   /* a missing file is a frame without boxes (rgbd_my.cc:235-236) */
    std::vector<char> buf;
    char chunk[4096];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
    std::fclose(f);
    return sdyn_boxes_parse(buf.data(), buf.size(), xywh, cap, nOut);
}

int sdyn_box_track(double* boxesIO, int nboxes, int cap, const double* lastObjects, const int32_t* lastBoxIdx,
                   const uint8_t* lastOmit, const double* lastVelocity, int nlast, int imgW, int imgH, int32_t* boxIdx,
                   uint8_t* omit, double* velocity, int* nOut)
{
    if (!nOut || nboxes < 0 || nlast < 0 || cap < nboxes || (cap > 0 && (!boxesIO || !boxIdx || !omit || !velocity)) ||
        (nlast > 0 && (!lastObjects || !lastBoxIdx || !lastOmit || !lastVelocity)))
        return SDYN_ERR_ARG;
    Rect* boxes = reinterpret_cast<Rect*>(boxesIO);
    const Rect* last = reinterpret_cast<const Rect*>(lastObjects);
    int n = nboxes;
    for (int j = 0; j < nboxes; ++j) { boxIdx[j] = -1; omit[j] = 0; velocity[2 * j] = velocity[2 * j + 1] = 0.0; }
    if (nlast == 0) {                                          /* re-initialised: number from zero (:546-550) */
        for (int j = 0; j < nboxes; ++j) boxIdx[j] = j;
        *nOut = n;
        return SDYN_OK;
    }
    /* every object of the last frame picks the current box of largest IoU (:488-517) */
    for (int i = 0; i < nlast; ++i) {
        double best = 1;
        int at = -1;
        for (int j = 0; j < nboxes; ++j) {
            const Rect in = intersect(last[i], boxes[j]), un = bounding(last[i], boxes[j]);
            const double cost = 1 - (in.w * in.h) / (un.w * un.h);
            if (cost < best) { best = cost; at = j; }
        }
        if (at != -1 && !lastOmit[i]) {
            boxIdx[at] = lastBoxIdx[i];
            velocity[2 * at] = boxes[at].x + boxes[at].w / 2 - last[i].x - last[i].w / 2;
            velocity[2 * at + 1] = boxes[at].y + boxes[at].h / 2 - last[i].y - last[i].h / 2;
        }
    }
    /* objects that found no box are carried over once, moved by their velocity, while their centre stays in the image (:519-536) */
    for (int i = 0; i < nlast; ++i) {
        if (lastOmit[i]) continue;
        if (std::count(boxIdx, boxIdx + n, lastBoxIdx[i])) continue;
        const float cxf = (float)(last[i].x + last[i].w / 2 + lastVelocity[2 * i]);
        const float cyf = (float)(last[i].y + last[i].h / 2 + lastVelocity[2 * i + 1]);
        const float rw = (float)imgW, rh = (float)imgH;        /* Rect2f(Point2f(0,0), Point2f(cols, rows)) */
        if (0.0f <= cxf && cxf < rw && 0.0f <= cyf && cyf < rh) {
            if (n >= cap) return SDYN_ERR_CAPACITY;
            boxes[n] = Rect{last[i].x + lastVelocity[2 * i], last[i].y + lastVelocity[2 * i + 1], last[i].w, last[i].h};
            boxIdx[n] = lastBoxIdx[i]; omit[n] = 1;
            velocity[2 * n] = lastVelocity[2 * i]; velocity[2 * n + 1] = lastVelocity[2 * i + 1];
            ++n;
        }
    }
    /* boxes of this frame that matched nothing are new objects: next free number (:540-545) */
    for (int j = 0; j < nboxes; ++j) {
        if (boxIdx[j] != -1) continue;
        boxIdx[j] = *std::max_element(boxIdx, boxIdx + n) + 1;
    }
    *nOut = n;
    return SDYN_OK;
}

}  // extern "C"
