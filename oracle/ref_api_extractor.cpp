/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points over the REFERENCE'S OWN ORBextractor, compiled unchanged from
 * /root/reference/src/ORBextractor.cc against oracle/ref_shim (minicv) by `make -C oracle ref` into oracle/_ref/libref.so.
 * tests/test_oracle_ref.py uses it to pin the restatement (oracle/orc_extractor.cpp) to the reference itself.
 *
 * Allocator switch.  ORBextractor.cc:684 sorts pair<int, ExtractorNode*>, so equal-size octree nodes are ordered by heap
 * address.  ref_set_alloc_mode(1) serves every `operator new` issued while an extraction runs from
 * a monotonic arena (later allocation = higher address — the order the oracle pins); mode 0 is glibc malloc, whose
 * address order is whatever its free lists yield.  Tests compare both against the oracle and report the disagreement
 * rate of mode 0.  The library is linked -Bsymbolic so the replacement `operator new/delete` stays private to it.
 */
#include <opencv2/core/core.hpp>
#include "ORBextractor.h"
#include "ref_shim/ref_arena.h"
#include <cstdlib>
#include <cstring>
#include <new>
#include <sys/mman.h>

namespace {
constexpr size_t ARENA_BYTES = (size_t)64 << 30;   /* address space only (MAP_NORESERVE) */
char* g_arena = nullptr;
thread_local size_t t_off = 0;
thread_local bool t_bump = false;
int g_mode = 0;

inline bool in_arena(const void* p) { return g_arena && (const char*)p >= g_arena && (const char*)p < g_arena + ARENA_BYTES; }

void* ref_alloc(size_t n)
{
    if (t_bump) {
        const size_t a = (t_off + 15) & ~(size_t)15;
        if (a + n <= ARENA_BYTES) { t_off = a + n; return g_arena + a; }
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void ref_free(void* p) { if (p && !in_arena(p)) std::free(p); }

size_t g_persist = 0;      /* arena bytes below this offset belong to objects that outlive their scope (Frames) */
}  // namespace

namespace refapi {
BumpScope::BumpScope(bool persist_) : on(g_mode == 1), persist(persist_)
{
    if (!on) return;
    if (!g_arena) {
        void* m = mmap(nullptr, ARENA_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) { on = false; return; }
        g_arena = (char*)m;
    }
    t_off = g_persist; t_bump = true;
}
BumpScope::~BumpScope()
{
    if (!on) return;
    t_bump = false;
    const size_t end = (t_off + 4095) & ~(size_t)4095;
    if (persist) g_persist = end;
    else madvise(g_arena + g_persist, end - g_persist, MADV_DONTNEED);   /* give the pages back */
}
void BumpScope::pause(bool p) { if (on) t_bump = !p; }
size_t arena_used() { return t_off; }
}  // namespace refapi
using refapi::BumpScope;

void* operator new(size_t n) { return ref_alloc(n); }
void* operator new[](size_t n) { return ref_alloc(n); }
void operator delete(void* p) noexcept { ref_free(p); }
void operator delete[](void* p) noexcept { ref_free(p); }
void operator delete(void* p, size_t) noexcept { ref_free(p); }
void operator delete[](void* p, size_t) noexcept { ref_free(p); }

using ORB_SLAM2::ORBextractor;

extern "C" {

/* 0 = glibc malloc (the reference's real behaviour on this libc), 1 = monotonic arena during extraction */
void ref_set_alloc_mode(int mode) { g_mode = mode; }
double ref_arena_used_mb() { return refapi::arena_used() / 1048576.0; }

void* ref_extractor_create(int nf, float sf, int nl, int ini, int mn) { return new ORBextractor(nf, sf, nl, ini, mn); }
void ref_extractor_destroy(void* e) { delete (ORBextractor*)e; }

/* ORBextractor::operator() on a caller image.  Returns the keypoint count (copies at most cap). */
int ref_extractor_run(void* e, const uint8_t* img, int w, int h, int stride, cv::KeyPoint* kps, uint8_t* desc, int cap)
{
    ORBextractor* E = (ORBextractor*)e;
    int n;
    {
        std::vector<cv::KeyPoint> k;
        cv::Mat d;
        {
            BumpScope scope;                 /* outputs are copied out before the arena is recycled */
            cv::Mat image(h, w, CV_8UC1, (void*)img, (size_t)stride);
            std::vector<cv::KeyPoint> kk;
            cv::Mat dd;
            (*E)(image, cv::Mat(), kk, dd);
            n = (int)kk.size();
            const int m = std::min(n, cap);
            if (m > 0) {
                std::memcpy(kps, kk.data(), sizeof(cv::KeyPoint) * (size_t)m);
                for (int i = 0; i < m; ++i) std::memcpy(desc + 32 * (size_t)i, dd.ptr(i), 32);
            }
            if (scope.on) {
                /* the pyramid Mats were allocated from the arena: re-home them on the heap before it is recycled */
                scope.pause(true);
                for (auto& lvl : E->mvImagePyramid) {
                    if (lvl.empty()) continue;
                    cv::Mat whole(lvl.rows + 38, lvl.cols + 38, lvl.type());
                    for (int r = 0; r < whole.rows; ++r) std::memcpy(whole.ptr(r), lvl.data - 19 * (ptrdiff_t)lvl.step - 19 + r * (ptrdiff_t)lvl.step, whole.cols);
                    lvl = whole(cv::Rect(19, 19, lvl.cols, lvl.rows));
                }
                scope.pause(false);
            }
        }
    }
    return n;
}

void ref_extractor_tables(void* e, float* scale, float* inv, float* sig, float* invsig)
{
    ORBextractor* E = (ORBextractor*)e;
    const std::vector<float> a = E->GetScaleFactors(), b = E->GetInverseScaleFactors(), c = E->GetScaleSigmaSquares(), d = E->GetInverseScaleSigmaSquares();
    for (int i = 0; i < E->GetLevels(); ++i) { scale[i] = a[i]; inv[i] = b[i]; sig[i] = c[i]; invsig[i] = d[i]; }
}

int ref_extractor_level_dims(void* e, int l, int* w, int* h)
{
    ORBextractor* E = (ORBextractor*)e;
    if (l < 0 || l >= (int)E->mvImagePyramid.size() || E->mvImagePyramid[l].empty()) return -1;
    *w = E->mvImagePyramid[l].cols; *h = E->mvImagePyramid[l].rows;
    return 0;
}

/* copies the bordered level, (w+38) x (h+38): mvImagePyramid[l] is a ROI at (19,19) of its parent */
int ref_extractor_level_copy(void* e, int l, uint8_t* out)
{
    ORBextractor* E = (ORBextractor*)e;
    if (l < 0 || l >= (int)E->mvImagePyramid.size() || E->mvImagePyramid[l].empty()) return -1;
    const cv::Mat& m = E->mvImagePyramid[l];
    const int W = m.cols + 38, H = m.rows + 38;
    const uint8_t* base = m.data - 19 * (ptrdiff_t)m.step - 19;
    for (int r = 0; r < H; ++r) std::memcpy(out + (size_t)r * W, base + r * (ptrdiff_t)m.step, W);
    return 0;
}

}  // extern "C"
