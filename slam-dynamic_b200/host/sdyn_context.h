/* Process-wide access to a GPU context for the stateless drop-in entry points (ORBmatcher is a value type constructed at
 * every call site and used from Tracking, LocalMapping and LoopClosing concurrently: include/ORBmatcher.h:41-95).  One context
 * per calling thread, created on first use; the matcher entry points size their device arena on demand. */
#ifndef SDYN_HOST_CONTEXT_H
#define SDYN_HOST_CONTEXT_H
#include "../../include/sdyn.h"
#include <cstdio>

namespace sdyn_host {

inline int& DefaultDevice() { static int device = 0; return device; }

inline sdyn_ctx* ThreadContext()
{
    struct Holder {
        sdyn_ctx* ctx = nullptr;
        ~Holder() { if (ctx) sdyn_destroy(ctx); }
    };
    static thread_local Holder h;
    if (!h.ctx) {
        /* the ORB parameters are irrelevant to the searches: the smallest legal extractor (one level, 128 x 128) */
        const sdyn_orb_params p = {500, 1.2f, 1, 20, 7};
        if (sdyn_create(&p, 128, 128, 1, DefaultDevice(), &h.ctx) != SDYN_OK) {
            std::fprintf(stderr, "sdyn: no GPU context for the matcher: %s\n", sdyn_last_error(nullptr));
            h.ctx = nullptr;
        }
    }
    return h.ctx;
}

}  // namespace sdyn_host
#endif
