/* Batched, device-resident front end: extract -> SearchByProjection(cur,last) -> SearchByProjection(F,map)
 * -> dynamic mask, enqueued on one stream without host synchronisation (include/sdyn.h, "batched front end"). */
#include "match_internal.h"
#include <algorithm>
#include <cstring>

/* device-resident MapPoint table (sdyn_map_create) */
struct sdyn_map { int device; int capacity; sdyn_map_point* d; };
static const sdyn_map_point* sdyn_map_table(const sdyn_map* m) { return m ? m->d : nullptr; }

namespace sdyn {

int api_fail(sdyn_ctx* c, int code, const std::string& msg);

struct TrackState {
    int B = 0, cap = 0, maxQ = 0, refStride = 0, poolPerJob = 0, wantPool = 0; bool distorted = false;
    uint8_t* block = nullptr; size_t bytes = 0;
    /* carved from block */
    MatchJob* dJobs; int32_t* cellOff; int32_t* sorted; int32_t* cellOf; float4* gridEntry; int32_t* assign; uint8_t* locked;
    int2* qspan; int32_t* qperm; int32_t* qAccepted; int32_t* qBin; uint32_t* pool; int32_t* poolUsed; int32_t* qNext; int32_t* result;
    uint64_t* mask; unsigned long long* has; int8_t* slotMap; int32_t* boxList; int32_t* nnQ; int32_t* nnT; int32_t* readmit;
    int32_t* staticExit; uint8_t* dynMask; int32_t* counts;
    /* the frame the searches ran on in the last step when it was an rgbd_split step (Frame.cc:297-403 + UpdateFrame) */
    sdyn_keypoint* fKp; sdyn_keypoint* fKpUn; uint8_t* fDesc; int32_t* fOrder; int32_t* fCount; int32_t* fStatic;
    bool lastWasSplit = false;
    /* resident LastFrame: the tracked keypoint list of every slot's previous step */
    sdyn_keypoint* pKp; sdyn_keypoint* pKpUn; int32_t* pCount;
    /* query records gathered from ids + the MapPoint table (resident forms) */
    sdyn_last_point* gLast; sdyn_mappoint_query* gMap;
    int32_t* hOrder = nullptr; int32_t* hFCount = nullptr;
    size_t zeroFrom = 0, zeroBytes = 0;     /* region cleared at the start of every step */
    /* device staging of host inputs (sdyn_track_batch) */
    uint8_t* inBlock = nullptr; size_t inBytes = 0;
    /* asynchronous step in flight (sdyn_track_batch_async .. sdyn_track_wait) */
    struct Pending { bool active = false; int nframes = 0, cap = 0; sdyn_keypoint* kp = nullptr; uint8_t* desc = nullptr;
                     int* nOut = nullptr; int32_t* assign = nullptr; uint8_t* locked = nullptr; uint8_t* mask = nullptr; } pending;
    /* pinned staging for sdyn_track_fetch */
    int32_t* hAssign = nullptr; uint8_t* hLocked = nullptr; uint8_t* hMask = nullptr; int32_t* hCounts = nullptr; int32_t* hResult = nullptr;
};

static void release_track_state(TrackState* t)
{
    if (!t) return;
    cudaFree(t->block); cudaFree(t->inBlock);
    cudaFreeHost(t->hAssign); cudaFreeHost(t->hLocked); cudaFreeHost(t->hMask); cudaFreeHost(t->hCounts); cudaFreeHost(t->hResult);
    cudaFreeHost(t->hOrder); cudaFreeHost(t->hFCount);
    delete t;
}

void free_track_state(sdyn_ctx* c)
{
    release_track_state(static_cast<TrackState*>(c->track));
    c->track = nullptr;
}

static int ensure_track_state(sdyn_ctx* c, int maxQ, int refStride)
{
    TrackState* t = static_cast<TrackState*>(c->track);
    const int B = c->maxBatch, cap = c->maxKp;
    if (t && t->maxQ >= maxQ && t->refStride >= refStride && t->distorted == (c->camera.enabled != 0) && t->wantPool <= t->poolPerJob) return SDYN_OK;
    const int wantPool = t ? t->wantPool : 0;
    if (t) { maxQ = std::max(maxQ, t->maxQ); refStride = std::max(refStride, t->refStride); }
    if (t && t->pending.active)
        return api_fail(c, SDYN_ERR_ARG, "track state would have to grow while an asynchronous step is in flight: call sdyn_track_wait first");
    /* the state is re-allocated larger; what it carries from step to step — every slot's resident LastFrame — moves over */
    TrackState* old = t;
    if (old) { cudaStreamSynchronize(c->stream); c->track = nullptr; }
    t = new TrackState();
    t->B = B; t->cap = cap; t->maxQ = maxQ; t->refStride = refStride; t->distorted = c->camera.enabled != 0;
    t->poolPerJob = std::max(std::max(64 * maxQ, 1 << 16), wantPool);      /* grown after an overflow (track_fetch_finish) */
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t J = 2 * (size_t)B;
    const size_t oJobs = take(J * sizeof(MatchJob));
    const size_t oCellOff = take((size_t)B * (kGridCells + 1) * 4), oSorted = take((size_t)B * cap * 4), oCellOf = take((size_t)B * cap * 4);
    const size_t oGridEntry = take((size_t)B * cap * sizeof(float4));
    const size_t oQperm = take((size_t)B * maxQ * 4);
    const size_t oQspan = take(J * maxQ * sizeof(int2)), oQAcc = take(J * maxQ * 4), oQBin = take(J * maxQ * 4);
    const size_t oPool = take(J * (size_t)t->poolPerJob * 4);
    const size_t oMask = take((size_t)B * cap * 8), oHas = take((size_t)B * 8), oSlotMap = take((size_t)B * 128);
    const size_t oBoxList = take((size_t)B * 64 * cap * 4), oNnQ = take((size_t)B * 64 * cap * 4);
    const size_t oNnT = take((size_t)B * 64 * std::max(refStride, 1) * 4);
    const size_t oDynMask = take((size_t)B * cap);
    const bool distorted = c->camera.enabled != 0;
    const size_t oFKp = take((size_t)B * cap * sizeof(sdyn_keypoint)), oFKpUn = distorted ? take((size_t)B * cap * sizeof(sdyn_keypoint)) : oFKp;
    const size_t oFDesc = take((size_t)B * cap * 32), oFOrder = take((size_t)B * cap * 4), oFCount = take((size_t)B * 4), oFStatic = take((size_t)B * 4);
    const size_t oPKp = take((size_t)B * cap * sizeof(sdyn_keypoint)), oPKpUn = distorted ? take((size_t)B * cap * sizeof(sdyn_keypoint)) : oPKp;
    const size_t oGLast = take((size_t)B * maxQ * sizeof(sdyn_last_point)), oGMap = take((size_t)B * maxQ * sizeof(sdyn_mappoint_query));
    /* cleared every step, contiguous: */
    const size_t zeroFrom = off;
    const size_t oAssign = take((size_t)B * cap * 4);      /* set to -1 separately */
    const size_t oLocked = take((size_t)B * cap);
    const size_t oPoolUsed = take(J * 4), oQNext = take(J * 4), oResult = take(J * 4 * 4);
    const size_t oReadmit = take((size_t)B * cap * 4), oStatic = take((size_t)B * 4), oCounts = take((size_t)B * 16);
    const size_t oPCount = take((size_t)B * 4);      /* outside the per-step clear: it carries over to the next step */
    t->zeroFrom = oLocked; t->zeroBytes = oPCount - oLocked;
    (void)zeroFrom;
    t->bytes = off;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&t->block), off);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->hAssign), (size_t)B * cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->hLocked), (size_t)B * cap);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->hMask), (size_t)B * cap);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->hCounts), (size_t)B * 16);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->hResult), J * 16);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->hOrder), (size_t)B * cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&t->hFCount), (size_t)B * 8);
    if (e == cudaSuccess) e = cudaMemset(t->block + oPCount, 0, (size_t)B * 4);
    if (e != cudaSuccess) {
        release_track_state(t);
        c->track = old;                     /* the smaller state stays usable */
        return api_fail(c, SDYN_ERR_NOMEM, std::string("track state: ") + cudaGetErrorString(e));
    }
    uint8_t* b = t->block;
    t->dJobs = reinterpret_cast<MatchJob*>(b + oJobs);
    t->cellOff = reinterpret_cast<int32_t*>(b + oCellOff); t->sorted = reinterpret_cast<int32_t*>(b + oSorted);
    t->cellOf = reinterpret_cast<int32_t*>(b + oCellOf);
    t->gridEntry = reinterpret_cast<float4*>(b + oGridEntry);
    t->qperm = reinterpret_cast<int32_t*>(b + oQperm);
    t->qspan = reinterpret_cast<int2*>(b + oQspan); t->qAccepted = reinterpret_cast<int32_t*>(b + oQAcc);
    t->qBin = reinterpret_cast<int32_t*>(b + oQBin); t->pool = reinterpret_cast<uint32_t*>(b + oPool);
    t->mask = reinterpret_cast<uint64_t*>(b + oMask); t->has = reinterpret_cast<unsigned long long*>(b + oHas); t->slotMap = reinterpret_cast<int8_t*>(b + oSlotMap);
    t->boxList = reinterpret_cast<int32_t*>(b + oBoxList); t->nnQ = reinterpret_cast<int32_t*>(b + oNnQ);
    t->nnT = reinterpret_cast<int32_t*>(b + oNnT); t->dynMask = b + oDynMask;
    t->assign = reinterpret_cast<int32_t*>(b + oAssign); t->locked = b + oLocked;
    t->poolUsed = reinterpret_cast<int32_t*>(b + oPoolUsed); t->qNext = reinterpret_cast<int32_t*>(b + oQNext); t->result = reinterpret_cast<int32_t*>(b + oResult);
    t->fKp = reinterpret_cast<sdyn_keypoint*>(b + oFKp); t->fKpUn = reinterpret_cast<sdyn_keypoint*>(b + oFKpUn); t->fDesc = b + oFDesc;
    t->fOrder = reinterpret_cast<int32_t*>(b + oFOrder); t->fCount = reinterpret_cast<int32_t*>(b + oFCount); t->fStatic = reinterpret_cast<int32_t*>(b + oFStatic);
    t->pKp = reinterpret_cast<sdyn_keypoint*>(b + oPKp); t->pKpUn = reinterpret_cast<sdyn_keypoint*>(b + oPKpUn); t->pCount = reinterpret_cast<int32_t*>(b + oPCount);
    t->gLast = reinterpret_cast<sdyn_last_point*>(b + oGLast); t->gMap = reinterpret_cast<sdyn_mappoint_query*>(b + oGMap);
    t->readmit = reinterpret_cast<int32_t*>(b + oReadmit); t->staticExit = reinterpret_cast<int32_t*>(b + oStatic);
    t->counts = reinterpret_cast<int32_t*>(b + oCounts);
    if (old) {
        const size_t kb = (size_t)B * cap * sizeof(sdyn_keypoint);
        e = cudaMemcpy(t->pKp, old->pKp, kb, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess && t->pKpUn != t->pKp)
            e = cudaMemcpy(t->pKpUn, old->pKpUn != old->pKp ? old->pKpUn : old->pKp, kb, cudaMemcpyDeviceToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(t->pCount, old->pCount, (size_t)B * 4, cudaMemcpyDeviceToDevice);
        t->lastWasSplit = old->lastWasSplit;
        release_track_state(old);
        if (e != cudaSuccess) {
            release_track_state(t);
            return api_fail(c, SDYN_ERR_CUDA, std::string("track state: ") + cudaGetErrorString(e));
        }
    }
    c->track = t;
    return SDYN_OK;
}

/* The 2*B job descriptors of a step differ only in per-frame base pointers, which are affine in the frame index.
 * They are built ON the device from the descriptors of frames 0 and 1 passed as kernel arguments: word-wise
 * J(f) = J(0) + f * (J(1) - J(0)) reproduces every pointer and leaves every other field untouched.  No host-to-device
 * copy of the job table is left in the step — the only per-step H2D traffic is the caller's data (a pageable
 * cudaMemcpyAsync here would also synchronise the stream with the host). */
struct JobPair { MatchJob j0, j1; };
static_assert(sizeof(MatchJob) % 8 == 0, "MatchJob is copied as 64-bit words");

__global__ void k_build_jobs(const __grid_constant__ JobPair F, const __grid_constant__ JobPair M, MatchJob* __restrict__ jobs,
                             int B, int nframes)
{
    constexpr int W = (int)(sizeof(MatchJob) / 8);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * nframes * W) return;
    const int w = i % W, jf = i / W;
    const bool isMap = jf >= nframes;
    const int f = isMap ? jf - nframes : jf;
    const JobPair& P = isMap ? M : F;
    const unsigned long long a = reinterpret_cast<const unsigned long long*>(&P.j0)[w];
    const unsigned long long b = reinterpret_cast<const unsigned long long*>(&P.j1)[w];
    reinterpret_cast<unsigned long long*>(jobs + (isMap ? B : 0) + f)[w] = a + (unsigned long long)f * (b - a);
}

/* Last kernel of a step.  CTA 0 copies the two match counts of every frame; then the step's tracked keypoint lists become the
 * resident LastFrame of the next step (one sequence per slot) — unless a search ran out of candidate pool: that step's matches
 * are incomplete, the host is told to repeat it (track_fetch_finish), and every slot must still hold the LastFrame it had. */
__global__ void __launch_bounds__(256)
k_copy_match_counts(const int32_t* __restrict__ result, int B, int nframes, int32_t* __restrict__ counts,
                    const uint32_t* __restrict__ sKp, const uint32_t* __restrict__ sKpUn, const int32_t* __restrict__ sCount,
                    uint32_t* __restrict__ pKp, uint32_t* __restrict__ pKpUn, int32_t* __restrict__ pCount, unsigned long long words)
{
    __shared__ int overflow;
    if (threadIdx.x == 0) overflow = 0;
    __syncthreads();
    for (int f = threadIdx.x; f < nframes; f += 256)
        if (result[f * 4 + 2] | result[(B + f) * 4 + 2]) overflow = 1;
    __syncthreads();
    if (blockIdx.x == 0)
        for (int f = threadIdx.x; f < nframes; f += 256) {
            counts[f * 4] = result[f * 4];                 /* SearchByProjection(cur, last) */
            counts[f * 4 + 1] = result[(B + f) * 4];       /* SearchByProjection(F, map points) */
        }
    if (overflow) return;
    const unsigned long long stride = (unsigned long long)gridDim.x * 256;
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < words; i += stride) {
        pKp[i] = sKp[i];
        if (pKpUn) pKpUn[i] = sKpUn[i];
    }
    if (blockIdx.x == 0)
        for (int f = threadIdx.x; f < nframes; f += 256) pCount[f] = sCount[f];
}

int track_fetch_enqueue(sdyn_ctx* c, int nframes, int32_t* assign, uint8_t* locked, uint8_t* dynMask, int32_t* counts, int cap,
                        cudaStream_t st)
{
    TrackState* t = static_cast<TrackState*>(c->track);
    const int kc = t->cap;
    const bool direct = cap == kc;
#define TQ(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return api_fail(c, SDYN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
    if (assign) TQ(cudaMemcpyAsync(direct ? assign : t->hAssign, t->assign, (size_t)nframes * kc * 4, cudaMemcpyDeviceToHost, st));
    if (locked) TQ(cudaMemcpyAsync(direct ? locked : t->hLocked, t->locked, (size_t)nframes * kc, cudaMemcpyDeviceToHost, st));
    if (dynMask) TQ(cudaMemcpyAsync(direct ? dynMask : t->hMask, t->dynMask, (size_t)nframes * kc, cudaMemcpyDeviceToHost, st));
    TQ(cudaMemcpyAsync(counts ? counts : t->hCounts, t->counts, (size_t)nframes * 16, cudaMemcpyDeviceToHost, st));
    TQ(cudaMemcpyAsync(t->hResult, t->result, (size_t)2 * t->B * 16, cudaMemcpyDeviceToHost, st));
#undef TQ
    return SDYN_OK;
}

int track_fetch_finish(sdyn_ctx* c, int nframes, int32_t* assign, uint8_t* locked, uint8_t* dynMask, int cap)
{
    TrackState* t = static_cast<TrackState*>(c->track);
    const int kc = t->cap;
    for (int f = 0; f < nframes; ++f)
        if (t->hResult[f * 4 + 2] || t->hResult[(t->B + f) * 4 + 2]) {
            /* k_match_candidates reserves the un-gated bound (every keypoint of every touched cell): a dense frame or a wide
             * window can exceed it.  The step's matches are incomplete; the pool is doubled for the next call, so repeating
             * the step succeeds (the single-search entry points retry internally, sdyn_match.cpp). */
            t->wantPool = 2 * t->poolPerJob;
            return api_fail(c, SDYN_ERR_CAPACITY, "matcher candidate pool exhausted in the batched front end: pool doubled, repeat the step (no slot's resident LastFrame was advanced)");
        }
    if (cap != kc) {
        const int m = std::min(cap, kc);
        for (int f = 0; f < nframes; ++f) {
            if (assign) std::memcpy(assign + (size_t)f * cap, t->hAssign + (size_t)f * kc, (size_t)m * 4);
            if (locked) std::memcpy(locked + (size_t)f * cap, t->hLocked + (size_t)f * kc, m);
            if (dynMask) std::memcpy(dynMask + (size_t)f * cap, t->hMask + (size_t)f * kc, m);
        }
    }
    return SDYN_OK;
}

}  // namespace sdyn

using namespace sdyn;

#define TCU(c, call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return api_fail((c), SDYN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

static bool bad_track_inputs(const sdyn_ctx* c, const sdyn_track_inputs* in, int nframes)
{
    if (!in || nframes < 1 || nframes > c->maxBatch || in->last_stride < 0 || in->map_stride < 0 || in->ref_stride < 0 ||
        !(in->max_x > in->min_x) || !(in->max_y > in->min_y) || !in->boxes || !in->n_boxes || !in->ref_box || !in->ref_off ||
        !in->fmat || !in->poses || c->maxKp > 65535)
        return true;
    if (in->last_stride > 0) {
        const bool explicitForm = in->last_points && in->last_keys && in->n_last;
        const bool residentForm = !in->last_points && in->last_ids && in->last_flags && in->map;
        if (!explicitForm && !residentForm) return true;
    }
    if (in->map_stride > 0) {
        const bool explicitForm = in->map_points && in->n_map;
        const bool residentForm = !in->map_points && in->map_ids && (in->map_proj || in->map_flags) && in->map && in->n_map;
        if (!explicitForm && !residentForm) return true;
    }
    return false;
}

/* Steps 2-4 of the batched front end on the keypoints / descriptors the context's last extraction left on the device.
 * uRight: mvuRight of the current frames ([B][cap], device) for the stereo gates of the two searches, or null. */
static int track_after_extract(sdyn_ctx* c, int nframes, const sdyn_track_inputs* in, const float* uRight, cudaStream_t st);

int sdyn_track_batch_device(sdyn_ctx* c, int nframes, const uint8_t* dGray, size_t frameStride, int W, int H, int stride,
                            const sdyn_track_inputs* in, void* stream)
{
    if (!c) return SDYN_ERR_ARG;
    if (bad_track_inputs(c, in, nframes) || !dGray || W < 1 || H < 1 || stride < W)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_track_batch_device: bad argument");
    TCU(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = ensure_geometry(c, W, H);
    if (rc != SDYN_OK) return rc;
    rc = ensure_track_state(c, std::max(std::max(in->last_stride, in->map_stride), 1), in->ref_stride);
    if (rc != SDYN_OK) return rc;
    rc = enqueue_extract(c, nframes, dGray, frameStride, stride, st);
    if (rc != SDYN_OK) return rc;
    return track_after_extract(c, nframes, in, nullptr, st);
}

int sdyn_track_batch_stereo_device(sdyn_ctx* c, sdyn_ctx* right, int nframes, const uint8_t* dGrayLeft, const uint8_t* dGrayRight,
                                   size_t frameStride, int W, int H, int stride, const sdyn_track_inputs* in, float mb, float mbf)
{
    if (!c) return SDYN_ERR_ARG;
    if (!right || bad_track_inputs(c, in, nframes) || nframes > right->maxBatch || !dGrayLeft || !dGrayRight || W < 1 || H < 1 ||
        stride < W)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_track_batch_stereo_device: bad argument");
    TCU(c, cudaSetDevice(c->device));
    int rc = ensure_geometry(c, W, H);
    if (rc == SDYN_OK) rc = ensure_geometry(right, W, H);
    if (rc != SDYN_OK) return rc;
    rc = ensure_track_state(c, std::max(std::max(in->last_stride, in->map_stride), 1), in->ref_stride);
    if (rc != SDYN_OK) return rc;
    /* the stereo Frame constructor (src/Frame.cc:140-175): both extractions side by side, ComputeStereoMatches, then the
     * searches gate on mvuRight */
    rc = enqueue_extract(c, nframes, dGrayLeft, frameStride, stride, c->stream);
    if (rc == SDYN_OK) rc = enqueue_extract(right, nframes, dGrayRight, frameStride, stride, right->stream);
    if (rc != SDYN_OK) return rc == SDYN_OK ? rc : api_fail(c, rc, "sdyn_track_batch_stereo_device: extraction failed");
    rc = sdyn_stereo_match_device(c, right, nframes, mb, mbf, nullptr);
    if (rc != SDYN_OK) return rc;
    sdyn_stereo_view sv;
    if (sdyn_stereo_results(c, &sv) != SDYN_OK) return api_fail(c, SDYN_ERR_ARG, "stereo results unavailable");
    return track_after_extract(c, nframes, in, sv.u_right, c->stream);
}

static int track_after_extract(sdyn_ctx* c, int nframes, const sdyn_track_inputs* in, const float* uRight, cudaStream_t st)
{
    TrackState* t = static_cast<TrackState*>(c->track);
    const int B = t->B, cap = t->cap;
    const bool residentLast = in->last_stride > 0 && !in->last_points, residentMap = in->map_stride > 0 && !in->map_points;
    const bool split = in->rgbd_split != 0;
    const sdyn_keypoint* keysUn0 = c->camera.enabled ? c->dKpUn : c->dKp;
    /* the frame the searches see: every extracted keypoint, or (rgbd_split) the static + re-admitted list */
    const sdyn_keypoint* sKp = split ? t->fKp : c->dKp;
    const sdyn_keypoint* sKpUn = split ? t->fKpUn : keysUn0;
    const uint8_t* sDesc = split ? t->fDesc : c->dDesc;
    const int32_t* sCount = split ? t->fCount : c->dCount;

    JobPair pF, pM;
    for (int f = 0; f < 2; ++f) {
        MatchJob J; std::memset(&J, 0, sizeof(J));
        J.keysUn = sKpUn + (size_t)f * cap; J.desc = sDesc + (size_t)f * cap * 32;
        J.uRight = uRight ? uRight + (size_t)f * cap : nullptr;
        J.nPtr = sCount + f; J.n = cap;
        J.minX = in->min_x; J.minY = in->min_y; J.maxX = in->max_x; J.maxY = in->max_y;
        J.gridWInv = static_cast<float>(SDYN_GRID_COLS) / static_cast<float>(in->max_x - in->min_x);
        J.gridHInv = static_cast<float>(SDYN_GRID_ROWS) / static_cast<float>(in->max_y - in->min_y);
        for (int l = 0; l < SDYN_MAX_LEVELS; ++l) J.scale[l] = l < c->scales.nlevels ? c->scales.scale[l] : 1.f;
        J.cellOff = t->cellOff + (size_t)f * (kGridCells + 1); J.sorted = t->sorted + (size_t)f * cap;
        J.cellOf = t->cellOf + (size_t)f * cap;
        J.gridEntry = t->gridEntry + (size_t)f * cap;
        J.assign = t->assign + (size_t)f * cap; J.locked = t->locked + (size_t)f * cap;
        J.poolCap = t->poolPerJob; J.distTh = SDYN_TH_HIGH;
        J.fx = in->fx; J.fy = in->fy; J.cx = in->cx; J.cy = in->cy; J.bf = in->bf;

        MatchJob F = J;                      /* SearchByProjection(CurrentFrame, LastFrame, th, bMono) */
        F.mode = MM_FRAME;
        const long long P = in->frame_pitch;
        F.pose = frame_part(in->poses, f, 96, P); F.mb = in->b; F.mono = in->mono;   /* Tcw, bForward, bBackward per frame, on the device */
        if (residentLast) {                  /* LastFrame = this slot's previous step; points gathered from the table */
            F.queries = t->gLast + (size_t)f * in->last_stride;
            F.qKeys = t->pKp + (size_t)f * cap; F.qKeysUn = t->pKpUn + (size_t)f * cap;
            F.nqPtr = t->pCount + f;
        } else {
            F.queries = frame_part(in->last_points, f, (size_t)in->last_stride * sizeof(sdyn_last_point), P);
            F.qKeys = frame_part(in->last_keys, f, (size_t)in->last_stride * sizeof(sdyn_keypoint), P);
            F.qKeysUn = frame_part(in->last_keys_un ? in->last_keys_un : in->last_keys, f, (size_t)in->last_stride * sizeof(sdyn_keypoint), P);
            F.nqPtr = frame_part(in->n_last, f, 4, P);
        }
        F.nq = in->last_stride;
        F.th = in->th_frame; F.checkOri = in->check_orientation;
        F.assignBase = 0;
        const size_t jf = (size_t)f, jm = (size_t)B + f;
        F.qspan = t->qspan + jf * t->maxQ; F.qAccepted = t->qAccepted + jf * t->maxQ; F.qBin = t->qBin + jf * t->maxQ;
        F.pool = t->pool + jf * t->poolPerJob; F.poolUsed = t->poolUsed + jf; F.qNext = t->qNext + jf; F.result = t->result + jf * 4;
        (f ? pF.j1 : pF.j0) = F;

        MatchJob M = J;                      /* SearchByProjection(Frame, vpMapPoints, th) */
        M.mode = MM_MAP;
        M.queries = residentMap ? t->gMap + (size_t)f * in->map_stride
                                : frame_part(in->map_points, f, (size_t)in->map_stride * sizeof(sdyn_mappoint_query), P);
        M.nqPtr = frame_part(in->n_map, f, 4, P); M.nq = in->map_stride;
        M.th = in->th_map; M.nnratio = in->nnratio_map; M.assignBase = in->last_stride;
        M.qperm = t->qperm + (size_t)f * t->maxQ;
        M.qspan = t->qspan + jm * t->maxQ; M.qAccepted = t->qAccepted + jm * t->maxQ; M.qBin = t->qBin + jm * t->maxQ;
        M.pool = t->pool + jm * t->poolPerJob; M.poolUsed = t->poolUsed + jm; M.qNext = t->qNext + jm; M.result = t->result + jm * 4;
        (f ? pM.j1 : pM.j0) = M;
    }
    TCU(c, cudaMemsetAsync(t->assign, 0xff, (size_t)B * cap * 4, st));
    TCU(c, cudaMemsetAsync(t->block + t->zeroFrom, 0, t->zeroBytes, st));
    if (residentLast || residentMap) {
        /* query records of the two searches from ids + the resident MapPoint table */
        FrustumParams fp = {in->fx, in->fy, in->cx, in->cy, in->bf, in->min_x, in->min_y, in->max_x, in->max_y, in->viewing_cos_limit,
                            logf((float)c->params.scale_factor), c->scales.nlevels};      /* Frame::mfLogScaleFactor = log(mfScaleFactor) */
        TCU(c, launch_gather_queries(sdyn_map_table(in->map), sdyn_map_capacity(in->map),
                                     residentLast ? in->last_ids : nullptr, in->last_flags, t->pCount, in->last_stride, t->gLast,
                                     residentMap ? in->map_ids : nullptr, in->map_proj, in->n_map, in->map_stride, t->gMap,
                                     in->map_flags, in->poses, fp, nframes, in->frame_pitch, st));
        c->launches += 1;
    }
    /* The dynamic-keypoint mask reads the extraction results only.  With stereo-constructor semantics the two searches do
     * not read the mask: it runs on the second stream beside them and both join in front of the step's last kernel.  With
     * rgbd_split the searched frame is a product of the mask stage, so the stages run in order. */
    const bool fork = !c->profiling && !split;      /* per-stage profiling keeps one stream so stage times stay unmixed */
    cudaStream_t ds = fork ? c->aux : st;
    if (fork) {
        TCU(c, cudaEventRecord(c->evFork2, st));
        TCU(c, cudaStreamWaitEvent(c->aux, c->evFork2, 0));
    }
    {
        StageTimer tm(c, ds, SDYN_STAGE_DYNAMIC);
        TCU(c, launch_dyn_stage(*in, c->dKp, keysUn0, c->dDesc, c->dCount, cap, t->mask, t->has, t->slotMap, t->boxList, t->nnQ, t->nnT,
                                std::max(t->refStride, 1), t->readmit, t->staticExit, t->dynMask, t->counts, nframes, ds));
        c->launches += 4;
        if (split) {
            TCU(c, launch_frame_compact(c->dKp, keysUn0, c->dDesc, c->dCount, cap, t->mask, t->readmit, t->staticExit, t->fKp, t->fKpUn,
                                        t->fDesc, t->fOrder, t->fCount, t->fStatic, nframes, ds));
            c->launches += 1;
        }
    }
    t->lastWasSplit = split;
    if (fork) TCU(c, cudaEventRecord(c->evJoin2, c->aux));
    {
        StageTimer tm(c, st, SDYN_STAGE_MATCH);
        {
            const int words = 2 * nframes * (int)(sizeof(MatchJob) / 8);
            k_build_jobs<<<(words + 255) / 256, 256, 0, st>>>(pF, pM, t->dJobs, B, nframes);
            TCU(c, cudaGetLastError());
        }
        /* grids of the B frames + visiting order of the B map searches: one launch */
        TCU(c, launch_grid_build(t->dJobs, nframes, cap, st, in->map_stride > 0 ? t->dJobs + B : nullptr, nframes));
        /* candidate generation only applies static gates, so both searches of every frame go out in ONE launch
         * when their jobs are contiguous (full batch); the claims are then resolved frame search first */
        const bool both = in->last_stride > 0 && in->map_stride > 0 && nframes == B;
        if (both) {
            StageTimer tc(c, st, SDYN_STAGE_CANDIDATES);
            TCU(c, launch_match_candidates(t->dJobs, 2 * B, std::max(in->last_stride, in->map_stride), cap, st));
        }
        const bool two = in->last_stride > 0 && in->map_stride > 0;
        if (!both && in->last_stride > 0) TCU(c, launch_match_candidates(t->dJobs, nframes, in->last_stride, cap, st));
        if (!both && in->map_stride > 0) TCU(c, launch_match_candidates(t->dJobs + B, nframes, in->map_stride, cap, st));
        /* the claims: frame search first, then the map search on the mvpMapPoints it left — per frame, so one CTA does both */
        if (two) TCU(c, launch_match_resolve(t->dJobs, nframes, MM_FRAME, cap, std::max(in->last_stride, in->map_stride), st, B));
        else if (in->last_stride > 0) TCU(c, launch_match_resolve(t->dJobs, nframes, MM_FRAME, cap, in->last_stride, st));
        else if (in->map_stride > 0) TCU(c, launch_match_resolve(t->dJobs + B, nframes, MM_MAP, cap, in->map_stride, st));
        /* kernels only: job table, grid, [query order], candidates (one launch for both searches of a full batch), resolves */
        c->launches += 2 + (both ? 1 : (in->last_stride > 0) + (in->map_stride > 0)) + ((in->last_stride > 0 || in->map_stride > 0) ? 1 : 0);
    }
    if (fork) TCU(c, cudaStreamWaitEvent(st, c->evJoin2, 0));
    {
        static_assert(sizeof(sdyn_keypoint) % 4 == 0, "keypoint lists are committed as 32-bit words");
        const unsigned long long words = (unsigned long long)nframes * cap * (sizeof(sdyn_keypoint) / 4);
        const int ctas = (int)std::min<unsigned long long>((words + 255) / 256, 4 * 148);
        k_copy_match_counts<<<std::max(ctas, 1), 256, 0, st>>>(
            t->result, B, nframes, t->counts, reinterpret_cast<const uint32_t*>(sKp), reinterpret_cast<const uint32_t*>(sKpUn), sCount,
            reinterpret_cast<uint32_t*>(t->pKp), t->pKpUn != t->pKp ? reinterpret_cast<uint32_t*>(t->pKpUn) : nullptr, t->pCount, words);
        TCU(c, cudaGetLastError());
        c->launches += 1;
    }
    return SDYN_OK;
}

/* The per-frame input arrays of a step, in upload order, and their offsets in the staging block for `n` frames (each
 * array starts on a 256-byte boundary).  Returns the block size. */
struct TrackItem { const void* src; size_t bytesPerFrame; size_t off; };

static size_t track_items(const sdyn_track_inputs* in, size_t n, TrackItem* items)
{
    const bool residentLast = !in->last_points && in->last_ids, residentMap = !in->map_points && in->map_ids;
    const bool sepUn = !residentLast && in->last_keys_un && in->last_keys_un != in->last_keys;
    const size_t ls = (size_t)in->last_stride, ms = (size_t)in->map_stride;
    const TrackItem init[SDYN_TRACK_INPUT_ARRAYS] = {
        {residentLast ? nullptr : in->last_points, residentLast ? 0 : ls * sizeof(sdyn_last_point), 0},
        {residentLast ? nullptr : in->last_keys, residentLast ? 0 : ls * sizeof(sdyn_keypoint), 0},
        /* mvKeysUn == mvKeys for an undistorted camera (src/Frame.cc:814-818): one upload serves both */
        {sepUn ? in->last_keys_un : nullptr, sepUn ? ls * sizeof(sdyn_keypoint) : 0, 0},
        {residentLast ? nullptr : in->n_last, residentLast ? 0 : (size_t)4, 0},
        {residentMap ? nullptr : in->map_points, residentMap ? 0 : ms * sizeof(sdyn_mappoint_query), 0},
        {in->n_map, 4, 0},
        {in->boxes, 64 * 4 * sizeof(double), 0},
        {in->n_boxes, 4, 0},
        {in->ref_box, 64 * 4, 0},
        {in->ref_desc, (size_t)in->ref_stride * 32, 0},
        {in->ref_xy, (size_t)in->ref_stride * 8, 0},
        {in->ref_off, 65 * 4, 0},
        {in->fmat, 9 * 4, 0},
        {in->poses, 24 * 4, 0},
        {residentLast ? in->last_ids : nullptr, residentLast ? ls * 4 : 0, 0},
        {residentLast ? in->last_flags : nullptr, residentLast ? ls : 0, 0},
        {residentMap ? in->map_ids : nullptr, residentMap ? ms * 4 : 0, 0},
        {residentMap ? in->map_proj : nullptr, (residentMap && in->map_proj) ? ms * sizeof(sdyn_map_proj) : 0, 0},
        {(residentMap && !in->map_proj) ? in->map_flags : nullptr, (residentMap && !in->map_proj) ? ms : 0, 0},
    };
    size_t total = 0;
    for (int i = 0; i < SDYN_TRACK_INPUT_ARRAYS; ++i) {
        items[i] = init[i];
        items[i].off = total;
        total += (items[i].bytesPerFrame * n + 255) / 256 * 256;
    }
    return total;
}

int sdyn_track_input_layout(int nframes, int last_stride, int map_stride, int ref_stride, int forms,
                            size_t offsets[SDYN_TRACK_INPUT_ARRAYS], size_t* total)
{
    if (nframes < 1 || last_stride < 0 || map_stride < 0 || ref_stride < 0 || !offsets || !total) return SDYN_ERR_ARG;
    sdyn_track_inputs in;
    memset(&in, 0, sizeof in);
    in.last_stride = last_stride; in.map_stride = map_stride; in.ref_stride = ref_stride;
    /* only the NULL-ness of the pointers matters to the layout */
    static const sdyn_keypoint kA = {}, kB = {};
    static const sdyn_last_point lp = {};
    static const sdyn_mappoint_query mq = {};
    static const int32_t ids = 0; static const uint8_t fl = 0; static const sdyn_map_proj pr = {};
    if (forms & SDYN_FORM_RESIDENT_LAST) { in.last_ids = &ids; in.last_flags = &fl; }
    else { in.last_points = &lp; in.last_keys = &kA; in.last_keys_un = (forms & SDYN_FORM_SEPARATE_KEYS_UN) ? &kB : &kA; }
    if (forms & SDYN_FORM_RESIDENT_MAP) { in.map_ids = &ids; if (forms & SDYN_FORM_DEVICE_FRUSTUM) in.map_flags = &fl; else in.map_proj = &pr; }
    else in.map_points = &mq;
    TrackItem items[SDYN_TRACK_INPUT_ARRAYS];
    *total = track_items(&in, (size_t)nframes, items);
    for (int i = 0; i < SDYN_TRACK_INPUT_ARRAYS; ++i) offsets[i] = items[i].off;
    return SDYN_OK;
}

int sdyn_track_record_layout(int last_stride, int map_stride, int ref_stride, int forms, size_t offsets[SDYN_TRACK_INPUT_ARRAYS],
                             size_t* pitch)
{
    /* one frame's worth of every array, each on a 256-byte boundary: exactly the array-major layout of a 1-frame batch */
    return sdyn_track_input_layout(1, last_stride, map_stride, ref_stride, forms, offsets, pitch);
}

int sdyn_track_batch(sdyn_ctx* c, int nframes, const uint8_t* gray, size_t frameStride, int W, int H, int stride,
                     const sdyn_track_inputs* in, sdyn_keypoint* kpOut, uint8_t* descOut, int* nOut, int32_t* assign,
                     uint8_t* locked, uint8_t* dynMask, int32_t* counts, int cap)
{
    int rc = sdyn_track_batch_async(c, nframes, gray, frameStride, W, H, stride, in, kpOut, descOut, nOut, assign, locked,
                                    dynMask, counts, cap);
    return rc == SDYN_OK ? sdyn_track_wait(c) : rc;
}

int sdyn_track_batch_async(sdyn_ctx* c, int nframes, const uint8_t* gray, size_t frameStride, int W, int H, int stride,
                           const sdyn_track_inputs* in, sdyn_keypoint* kpOut, uint8_t* descOut, int* nOut, int32_t* assign,
                           uint8_t* locked, uint8_t* dynMask, int32_t* counts, int cap)
{
    if (!c) return SDYN_ERR_ARG;
    if (!in || !gray || nframes < 1 || nframes > c->maxBatch || W < 1 || H < 1 || stride < W || W > c->maxW || H > c->maxH)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_track_batch: bad argument");
    TCU(c, cudaSetDevice(c->device));
    if (c->track && static_cast<TrackState*>(c->track)->pending.active)      /* before anything is enqueued or regrown */
        return api_fail(c, SDYN_ERR_ARG, "sdyn_track_batch_async: previous step not waited for");
    int rc = ensure_track_state(c, std::max(std::max(in->last_stride, in->map_stride), 1), in->ref_stride);
    if (rc != SDYN_OK) return rc;
    TrackState* t = static_cast<TrackState*>(c->track);
    /* device staging of the per-frame input arrays, laid out as sdyn_track_input_layout() describes */
    const size_t n = (size_t)nframes;
    TrackItem items[SDYN_TRACK_INPUT_ARRAYS];
    const bool records = in->frame_pitch > 0;            /* frame-major: one record per frame, one copy for the batch */
    const size_t total = records ? (size_t)in->frame_pitch * n : track_items(in, n, items);
    if (records) {
        const size_t one = track_items(in, 1, items);
        if (one > (size_t)in->frame_pitch || (in->frame_pitch & 255))
            return api_fail(c, SDYN_ERR_ARG, "sdyn_track_batch: frame_pitch is smaller than a frame record (sdyn_track_record_layout) or not a multiple of 256");
    }
    if (total > t->inBytes) {
        const size_t want = records ? (size_t)in->frame_pitch * c->maxBatch : std::max(total, track_items(in, (size_t)c->maxBatch, items));
        track_items(in, records ? 1 : n, items);
        TCU(c, cudaStreamSynchronize(c->stream));
        cudaFree(t->inBlock); t->inBlock = nullptr; t->inBytes = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&t->inBlock), want);
        if (e != cudaSuccess) return api_fail(c, SDYN_ERR_NOMEM, std::string("track input staging: ") + cudaGetErrorString(e));
        t->inBytes = want;
    }
    /* A caller that keeps its arrays in ONE host block at the layout's offsets gets ONE copy: a dozen small copies cost
     * more PCIe time than their bytes (each pays the copy engine's set-up latency). */
    const uint8_t* base = nullptr;
    for (const auto& it : items)
        if (!base && it.src && it.bytesPerFrame) base = static_cast<const uint8_t*>(it.src) - it.off;   /* first array in use */
    bool packed = base != nullptr;
    for (const auto& it : items)
        if (it.src && it.bytesPerFrame && static_cast<const uint8_t*>(it.src) != base + it.off) packed = false;
    if (records && !packed)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_track_batch: with frame_pitch every array must sit at its sdyn_track_record_layout offset of one record pool");
    if (packed) {
        TCU(c, cudaMemcpyAsync(t->inBlock, base, total, cudaMemcpyHostToDevice, c->stream));
    } else {
        for (const auto& it : items)
            if (it.src && it.bytesPerFrame)
                TCU(c, cudaMemcpyAsync(t->inBlock + it.off, it.src, it.bytesPerFrame * n, cudaMemcpyHostToDevice, c->stream));
    }
    TCU(c, upload_frames(c, nframes, gray, frameStride, W, H, stride, c->stream));
    sdyn_track_inputs d = *in;
    auto dev = [&](int i) -> const uint8_t* { return items[i].src ? t->inBlock + items[i].off : nullptr; };
    d.last_points = reinterpret_cast<const sdyn_last_point*>(dev(0));
    d.last_keys = reinterpret_cast<const sdyn_keypoint*>(dev(1));
    d.last_keys_un = items[2].src ? reinterpret_cast<const sdyn_keypoint*>(dev(2)) : d.last_keys;
    d.n_last = reinterpret_cast<const int32_t*>(dev(3));
    d.map_points = reinterpret_cast<const sdyn_mappoint_query*>(dev(4));
    d.n_map = reinterpret_cast<const int32_t*>(dev(5));
    d.poses = reinterpret_cast<const float*>(dev(13));
    d.last_ids = reinterpret_cast<const int32_t*>(dev(14)); d.last_flags = dev(15);
    d.map_ids = reinterpret_cast<const int32_t*>(dev(16)); d.map_proj = reinterpret_cast<const sdyn_map_proj*>(dev(17));
    d.map_flags = dev(18);
    d.boxes = reinterpret_cast<const double*>(t->inBlock + items[6].off);
    d.n_boxes = reinterpret_cast<const int32_t*>(t->inBlock + items[7].off);
    d.ref_box = reinterpret_cast<const int32_t*>(t->inBlock + items[8].off);
    d.ref_desc = t->inBlock + items[9].off;
    d.ref_xy = reinterpret_cast<const float*>(t->inBlock + items[10].off);
    d.ref_off = reinterpret_cast<const int32_t*>(t->inBlock + items[11].off);
    d.fmat = reinterpret_cast<const float*>(t->inBlock + items[12].off);
    rc = sdyn_track_batch_device(c, nframes, c->dIn, (size_t)W * H, W, H, W, &d, nullptr);
    if (rc != SDYN_OK) return rc;
    t = static_cast<TrackState*>(c->track);
    if (nOut) {
        rc = fetch_enqueue(c, nframes, kpOut, descOut, (kpOut && descOut) ? cap : 0, c->stream);
        if (rc != SDYN_OK) return rc;
    }
    rc = track_fetch_enqueue(c, nframes, assign, locked, dynMask, counts, cap, c->stream);
    if (rc != SDYN_OK) return rc;
    t->pending.active = true; t->pending.nframes = nframes; t->pending.cap = cap; t->pending.kp = kpOut; t->pending.desc = descOut;
    t->pending.nOut = nOut; t->pending.assign = assign; t->pending.locked = locked; t->pending.mask = dynMask;
    return SDYN_OK;
}

int sdyn_track_frame_order(sdyn_ctx* c, int nframes, int32_t* order, int32_t* n, int32_t* nStatic, int cap)
{
    if (!c) return SDYN_ERR_ARG;
    TrackState* t = static_cast<TrackState*>(c->track);
    if (!t || nframes < 1 || nframes > t->B || cap < 0 || !t->lastWasSplit)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_track_frame_order: the last step was not an rgbd_split step");
    TCU(c, cudaSetDevice(c->device));
    TCU(c, cudaMemcpyAsync(t->hOrder, t->fOrder, (size_t)nframes * t->cap * 4, cudaMemcpyDeviceToHost, c->stream));
    TCU(c, cudaMemcpyAsync(t->hFCount, t->fCount, (size_t)nframes * 4, cudaMemcpyDeviceToHost, c->stream));
    TCU(c, cudaMemcpyAsync(t->hFCount + t->B, t->fStatic, (size_t)nframes * 4, cudaMemcpyDeviceToHost, c->stream));
    TCU(c, cudaStreamSynchronize(c->stream));
    for (int f = 0; f < nframes; ++f) {
        if (n) n[f] = t->hFCount[f];
        if (nStatic) nStatic[f] = t->hFCount[t->B + f];
        if (order) std::memcpy(order + (size_t)f * cap, t->hOrder + (size_t)f * t->cap, (size_t)std::min(cap, t->cap) * 4);
    }
    return SDYN_OK;
}

/* ---- device-resident MapPoint table ---- */
void* sdyn_stream(sdyn_ctx* c) { return c ? (void*)c->stream : nullptr; }

int sdyn_map_create(int device, int capacity, sdyn_map** out)
{
    if (!out || capacity < 1) return SDYN_ERR_ARG;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return SDYN_ERR_CUDA;
    sdyn_map* m = new sdyn_map{device, capacity, nullptr};
    if (cudaMalloc(reinterpret_cast<void**>(&m->d), (size_t)capacity * sizeof(sdyn_map_point)) != cudaSuccess) { delete m; return SDYN_ERR_NOMEM; }
    cudaMemset(m->d, 0, (size_t)capacity * sizeof(sdyn_map_point));
    *out = m;
    return SDYN_OK;
}

int sdyn_map_destroy(sdyn_map* m)
{
    if (!m) return SDYN_ERR_ARG;
    cudaSetDevice(m->device);
    cudaFree(m->d);
    delete m;
    return SDYN_OK;
}

int sdyn_map_update(sdyn_map* m, int first, int count, const sdyn_map_point* pts, void* stream)
{
    if (!m || !pts || first < 0 || count < 0 || first + count > m->capacity) return SDYN_ERR_ARG;
    if (cudaSetDevice(m->device) != cudaSuccess) return SDYN_ERR_CUDA;
    if (count == 0) return SDYN_OK;
    return cudaMemcpyAsync(m->d + first, pts, (size_t)count * sizeof(sdyn_map_point), cudaMemcpyHostToDevice, (cudaStream_t)stream) == cudaSuccess
               ? SDYN_OK : SDYN_ERR_CUDA;
}

int sdyn_map_capacity(const sdyn_map* m) { return m ? m->capacity : 0; }

int sdyn_track_stats(const sdyn_ctx* c, int nframes, long long evals[2])
{
    if (!c || !c->track || !evals) return SDYN_ERR_ARG;
    const TrackState* t = static_cast<const TrackState*>(c->track);
    if (nframes < 1 || nframes > t->B) return SDYN_ERR_ARG;
    evals[0] = evals[1] = 0;
    for (int f = 0; f < nframes; ++f) { evals[0] += t->hResult[f * 4 + 3]; evals[1] += t->hResult[(t->B + f) * 4 + 3]; }
    return SDYN_OK;
}

int sdyn_track_results(const sdyn_ctx* c, sdyn_track_view* out)
{
    if (!c || !out || !c->track) return SDYN_ERR_ARG;
    const TrackState* t = static_cast<const TrackState*>(c->track);
    out->assign = t->assign; out->locked = t->locked; out->dyn_mask = t->dynMask; out->counts = t->counts;
    return SDYN_OK;
}

int sdyn_track_fetch(sdyn_ctx* c, int nframes, int32_t* assign, uint8_t* locked, uint8_t* dynMask, int32_t* counts, int cap,
                     void* stream)
{
    if (!c) return SDYN_ERR_ARG;
    TrackState* t = static_cast<TrackState*>(c->track);
    if (!t || nframes < 1 || nframes > t->B || cap < 0) return api_fail(c, SDYN_ERR_ARG, "sdyn_track_fetch: bad argument");
    TCU(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    int rc = track_fetch_enqueue(c, nframes, assign, locked, dynMask, counts, cap, st);
    if (rc != SDYN_OK) return rc;
    TCU(c, cudaStreamSynchronize(st));
    return track_fetch_finish(c, nframes, assign, locked, dynMask, cap);
}

int sdyn_track_wait(sdyn_ctx* c)
{
    if (!c) return SDYN_ERR_ARG;
    TrackState* t = static_cast<TrackState*>(c->track);
    if (!t || !t->pending.active) return api_fail(c, SDYN_ERR_ARG, "sdyn_track_wait: no asynchronous step in flight");
    TCU(c, cudaSetDevice(c->device));
    TCU(c, cudaStreamSynchronize(c->stream));
    const TrackState::Pending p = t->pending;
    t->pending.active = false;
    int rc = SDYN_OK;
    if (p.nOut) {
        rc = fetch_finish(c, p.nframes, p.kp, p.desc, (p.kp && p.desc) ? p.cap : 0, p.nOut);
        if (rc != SDYN_OK && !(rc == SDYN_ERR_CAPACITY && !(p.kp && p.desc))) return rc;
    }
    return track_fetch_finish(c, p.nframes, p.assign, p.locked, p.mask, p.cap);
}

}  // extern "C"
