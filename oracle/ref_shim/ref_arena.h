/* ORACLE — TEST INFRASTRUCTURE ONLY.  Allocation-order pin for the reference's octree tie-break (ORBextractor.cc:684, SURVEY
 * Appendix B-1); defined in oracle/ref_api_extractor.cpp. */
#pragma once
#include <cstddef>
namespace refapi {
/* While the scope lives (and ref_set_alloc_mode(1) is in force) `operator new` of the calling thread is served from a
 * monotonic arena: later allocation = higher address.
 *   persist = false: everything allocated inside is dead when the scope ends; the pages are recycled.
 *   persist = true:  the allocations stay valid for the life of the process (a Frame built inside keeps its vectors). */
struct BumpScope {
    bool on, persist;
    explicit BumpScope(bool persist = false);
    ~BumpScope();
    void pause(bool p);      /* temporarily route allocations to the heap */
};
size_t arena_used();
}  // namespace refapi
