/* C-ABI of libsdyn: context management and the ORBextractor::operator() replacement.
 * See include/sdyn.h for the contract; kernels live in k_*.cu. */
#include "sdyn_internal.h"
#include "tma.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>

using namespace sdyn;

namespace {

thread_local std::string g_createError = "";

int fail(sdyn_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg; else g_createError = msg;
    return code;
}

int cuda_fail(sdyn_ctx* c, cudaError_t e, const char* what)
{
    return fail(c, SDYN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(c, call)                                                   \
    do {                                                              \
        cudaError_t e_ = (call);                                      \
        if (e_ != cudaSuccess) return cuda_fail((c), e_, #call);      \
    } while (0)

template <class T>
cudaError_t dalloc(T** p, size_t count) { return cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(count, 1) * sizeof(T)); }

void build_tiles(const Geom& g, std::vector<TileRef>& fast, std::vector<TileRef>& blur)
{
    fast.clear(); blur.clear();
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom& L = g.L[l];
        for (int ty = 0; ty * kFastTileH < L.fh; ++ty)
            for (int tx = 0; tx * kFastTileW - kFastLead < L.fw; ++tx) fast.push_back({(int16_t)l, (int16_t)tx, (int16_t)ty, 0});
        for (int ty = 0; ty * kBlurTileH < L.h; ++ty)
            for (int tx = 0; tx * kBlurTileW < L.w; ++tx) blur.push_back({(int16_t)l, (int16_t)tx, (int16_t)ty, 0});
    }
}

}  // namespace

namespace sdyn {
/* (Re)computes geometry when the image size changes; device buffers were sized for (maxW, maxH). */
int ensure_geometry(sdyn_ctx* c, int W, int H)
{
    if (c->geomValid && c->geom.W == W && c->geom.H == H) return SDYN_OK;
    if (W > c->maxW || H > c->maxH) return fail(c, SDYN_ERR_ARG, "image larger than the size given to sdyn_create");
    Geom g; std::vector<uint8_t> tables;
    int rc = compute_geometry(c->params, c->scales, W, H, g, tables);
    if (rc != SDYN_OK)
        return fail(c, rc, "image shape unsupported: a pyramid level is too small for the 30-px FAST grid / octree "
                           "(the reference divides by zero here)");
    if ((size_t)g.frameBytes > c->pyrFrameCap || g.cellsPerFrame > c->cellCap || g.candPerFrame > c->candCap ||
        g.kpPerFrame > c->levelKpCap || tables.size() > c->tablesCap)
        return fail(c, SDYN_ERR_ARG, "internal: geometry exceeds the buffers sized at sdyn_create");
    std::vector<TileRef> ft, bt;
    build_tiles(g, ft, bt);
    /* the previous geometry may still be in use by enqueued kernels */
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(c->dTables, tables.data(), tables.size(), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->dFastTiles, ft.data(), ft.size() * sizeof(TileRef), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->dBlurTiles, bt.data(), bt.size() * sizeof(TileRef), cudaMemcpyHostToDevice));
    {
        const char* why = "";
        TmaMaps* tm = static_cast<TmaMaps*>(c->tma);
        cudaError_t e = encode_level_maps(g, c->dPyr, c->maxBatch, kBlurStageW, kBlurStageH, &tm->blurTile, &why);
        if (e == cudaSuccess) e = encode_level_maps(g, c->dPyr, c->maxBatch, kFastStageW, kFastStageH, &tm->fastTile, &why);
        if (e == cudaSuccess) e = encode_level_maps(g, c->dPyr, c->maxBatch, kPatchPitch, kOrientRows, &tm->orientPatch, &why);
        if (e == cudaSuccess) e = encode_level_maps(g, c->dBlur, c->maxBatch, kPatchPitch, kDescRows, &tm->descPatch, &why);
        if (e == cudaSuccess) e = encode_resize_maps(g, c->dPyr, c->maxBatch, &tm->resizeSrc, &why);
        if (e != cudaSuccess) return fail(c, SDYN_ERR_CUDA, std::string("tensor maps: ") + why + ": " + cudaGetErrorString(e));
    }
    c->geom = g; c->nFastTiles = (int)ft.size(); c->nBlurTiles = (int)bt.size();
    c->geomValid = true;
    if (c->graph1) { cudaGraphExecDestroy(c->graph1); c->graph1 = nullptr; }      /* captured for the previous geometry */
    return SDYN_OK;
}

}  // namespace sdyn

namespace {
cudaEvent_t take_event(sdyn_ctx* c)
{
    cudaEvent_t e = nullptr;
    if (!c->evPool.empty()) { e = c->evPool.back(); c->evPool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

/* Collects finished spans into c->acc (synchronises on the events). */
void drain_spans(sdyn_ctx* c)
{
    for (auto& s : c->spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            c->acc.ms[s.stage] += ms;
            c->acc.calls[s.stage] += 1;
        }
        c->evPool.push_back(s.a); c->evPool.push_back(s.b);
    }
    c->spans.clear();
}

}  // namespace

namespace sdyn {
StageTimer::StageTimer(sdyn_ctx* c_, cudaStream_t st_, int stage_) : c(c_), st(st_), stage(stage_), a(nullptr)
{
    if (c->profiling) { a = take_event(c); cudaEventRecord(a, st); }
}
StageTimer::~StageTimer()
{
    if (a) { cudaEvent_t b = take_event(c); cudaEventRecord(b, st); c->spans.push_back({a, b, stage}); }
}

/* H2D of the input frames; rows are packed to W on the device.  Contiguous inputs go out as ONE linear copy. */
cudaError_t upload_frames(sdyn_ctx* c, int nframes, const uint8_t* gray, size_t frameStride, int W, int H, int stride,
                          cudaStream_t st)
{
    if (stride == W && (nframes == 1 || frameStride == (size_t)W * H))
        return cudaMemcpyAsync(c->dIn, gray, (size_t)nframes * W * H, cudaMemcpyHostToDevice, st);
    for (int f = 0; f < nframes; ++f) {
        cudaError_t e = stride == W
            ? cudaMemcpyAsync(c->dIn + (size_t)f * W * H, gray + (size_t)f * frameStride, (size_t)W * H, cudaMemcpyHostToDevice, st)
            : cudaMemcpy2DAsync(c->dIn + (size_t)f * W * H, W, gray + (size_t)f * frameStride, stride, W, H,
                                cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

/* D2H of the extraction results: enqueue (asynchronous) and finish (after the stream has been synchronised). */
int fetch_enqueue(sdyn_ctx* c, int nframes, sdyn_keypoint* kpOut, uint8_t* descOut, int cap, cudaStream_t st)
{
    CU(c, cudaMemcpyAsync(c->hCount, c->dCount, sizeof(int32_t) * nframes, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(c->hStatus, c->dStatus, sizeof(int32_t) * nframes, cudaMemcpyDeviceToHost, st));
    /* same per-frame capacity as the device arrays: copy straight into the caller's buffers (asynchronous when
     * they are pinned); otherwise stage in pinned memory and scatter in fetch_finish */
    const bool direct = cap == c->maxKp;
    if (cap > 0) {
        CU(c, cudaMemcpyAsync(direct ? kpOut : c->hKp, c->dKp, sizeof(sdyn_keypoint) * (size_t)c->maxKp * nframes, cudaMemcpyDeviceToHost, st));
        CU(c, cudaMemcpyAsync(direct ? descOut : c->hDesc, c->dDesc, (size_t)32 * c->maxKp * nframes, cudaMemcpyDeviceToHost, st));
    }
    return SDYN_OK;
}

int fetch_finish(sdyn_ctx* c, int nframes, sdyn_keypoint* kpOut, uint8_t* descOut, int cap, int* nOut)
{
    const bool direct = cap == c->maxKp;
    int rc = SDYN_OK;
    for (int f = 0; f < nframes; ++f) {
        if (c->hStatus[f]) return fail(c, SDYN_ERR_CAPACITY, "internal candidate/node capacity exceeded");
        const int n = c->hCount[f];
        nOut[f] = n;
        const int m = std::min(n, cap);
        if (n > cap) rc = SDYN_ERR_CAPACITY;
        if (m > 0 && !direct) {
            std::memcpy(kpOut + (size_t)f * cap, c->hKp + (size_t)f * c->maxKp, sizeof(sdyn_keypoint) * m);
            std::memcpy(descOut + (size_t)f * cap * 32, c->hDesc + (size_t)f * c->maxKp * 32, (size_t)32 * m);
        }
    }
    if (rc != SDYN_OK) fail(c, rc, "output capacity too small (see sdyn_max_keypoints)");
    return rc;
}

/* Enqueues the whole extraction pipeline for nframes frames already resident in device memory. */
int enqueue_extract(sdyn_ctx* c, int nframes, const uint8_t* dGray, size_t frameStride, int rowStride, cudaStream_t st)
{
    const Geom& g = c->geom;
    const sdyn_orb_params& p = c->params;
    if (c->spans.size() > 4096) drain_spans(c);
    {
        StageTimer t(c, st, SDYN_STAGE_LEVEL0);
        CU(c, cudaMemsetAsync(c->dCandCount, 0, sizeof(int32_t) * SDYN_MAX_LEVELS * nframes, st));
        CU(c, cudaMemsetAsync(c->dCellFlag, 0, (size_t)g.cellsPerFrame * nframes, st));
        CU(c, launch_level0(g, dGray, frameStride, rowStride, c->dPyr, nframes, st));
    }
    {
        StageTimer t(c, st, SDYN_STAGE_PYRAMID);
        for (int l = 1; l < g.nlevels; ++l) CU(c, launch_resize(g, l, c->dTables, c->tma, c->dPyr, nframes, st));
    }
    /* The blur only needs the pyramid, FAST + octree only need the pyramid: two branches on two streams, joined in
     * front of the descriptor kernel (which reads the blurred levels and the selected keypoints). */
    /* (while per-stage profiling is on, everything stays on one stream so that stage times are not inflated by overlap) */
    const bool fork = !c->profiling;
    cudaStream_t bs = fork ? c->aux : st;
    if (fork) {
        CU(c, cudaEventRecord(c->evFork, st));
        CU(c, cudaStreamWaitEvent(c->aux, c->evFork, 0));
    }
    {
        StageTimer t(c, bs, SDYN_STAGE_BLUR);
        CU(c, launch_blur(g, c->dBlurTiles, c->nBlurTiles, c->tma, c->dBlur, nframes, bs));
    }
    if (fork) CU(c, cudaEventRecord(c->evJoin, c->aux));
    {
        StageTimer t(c, st, SDYN_STAGE_FAST);
        CU(c, launch_fast(g, c->dFastTiles, c->nFastTiles, c->tma, p.ini_th_fast, p.min_th_fast, c->dCellFlag,
                          c->dCand, c->dCandCount, nframes, st));
    }
    {
        StageTimer t(c, st, SDYN_STAGE_OCTREE);
        CU(c, launch_octree(g, p.ini_th_fast, p.min_th_fast, c->dCellFlag, c->dCand, c->dCandCount, c->dCandNode,
                            c->dSelCount, c->dLevelKp, c->dLevelCount, c->dStatus, nframes, st));
    }
    if (fork) CU(c, cudaStreamWaitEvent(st, c->evJoin, 0));
    {
        StageTimer t(c, st, SDYN_STAGE_DESCRIBE);
        CU(c, launch_orient_describe(g, c->tma, c->dLevelKp, c->dLevelCount, c->dKp, c->dDesc, c->dCount,
                                     c->maxKp, nframes, st));
    }
    c->launches += 1 + (g.nlevels - 1) + 4;      /* kernels only: level 0, resizes, FAST, octree, blur, describe (memsets are not counted) */
    if (c->camera.enabled) {          /* Frame::UndistortKeyPoints follows ExtractORB in every Frame constructor */
        StageTimer t(c, st, SDYN_STAGE_DESCRIBE);
        CU(c, launch_undistort(c->camera, c->dKp, c->dCount, c->maxKp, c->dKpUn, nframes, st));
        c->launches += 1;
    }
    return SDYN_OK;
}

}  // namespace sdyn

namespace {
void free_all(sdyn_ctx* c)
{
    sdyn::free_track_state(c);
    sdyn::free_stereo_state(c);
    sdyn::free_bow_state(c);
    cudaFree(c->dFastTiles); cudaFree(c->dBlurTiles); cudaFree(c->dTables); cudaFree(c->dIn); cudaFree(c->dPyr); delete static_cast<TmaMaps*>(c->tma); c->tma = nullptr;
    cudaFree(c->dBlur); cudaFree(c->dCellFlag); cudaFree(c->dCand); cudaFree(c->dCandNode); cudaFree(c->dCandCount); cudaFree(c->dSelCount);
    cudaFree(c->dLevelKp); cudaFree(c->dLevelCount); cudaFree(c->dCount); cudaFree(c->dStatus); cudaFree(c->dKp);
    cudaFree(c->dDesc); cudaFree(c->dArena); cudaFree(c->dKpUn);
    cudaFreeHost(c->hKp); cudaFreeHost(c->hDesc); cudaFreeHost(c->hCount); cudaFreeHost(c->hStatus); cudaFreeHost(c->hIn);
    if (c->graph1) cudaGraphExecDestroy(c->graph1);
    for (auto& sp : c->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : c->evPool) cudaEventDestroy(e);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    if (c->evFork2) cudaEventDestroy(c->evFork2);
    if (c->evJoin2) cudaEventDestroy(c->evJoin2);
    if (c->aux) cudaStreamDestroy(c->aux);
    if (c->stream) cudaStreamDestroy(c->stream);
}

}  // namespace

extern "C" {

int sdyn_create(const sdyn_orb_params* params, int maxW, int maxH, int maxBatch, int device, sdyn_ctx** out)
{
    if (!params || !out) return fail(nullptr, SDYN_ERR_ARG, "null argument");
    *out = nullptr;
    if (params->nlevels < 1 || params->nlevels > SDYN_MAX_LEVELS || params->nfeatures < 1 ||
        !(params->scale_factor > 1.0f) || params->ini_th_fast < 1 || params->min_th_fast < 1 ||
        params->ini_th_fast > 255 || params->min_th_fast > 255 || maxBatch < 1 || maxW < 1 || maxH < 1)
        return fail(nullptr, SDYN_ERR_ARG, "bad ORB parameters or limits");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, SDYN_ERR_CUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, SDYN_ERR_ARG, "device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");

    sdyn_ctx* c = new (std::nothrow) sdyn_ctx();
    if (!c) return fail(nullptr, SDYN_ERR_NOMEM, "out of host memory");
    c->params = *params; c->maxW = maxW; c->maxH = maxH; c->maxBatch = maxBatch; c->device = device;
    compute_scale_info(c->params, c->scales, c->umax);
    c->maxKp = max_keypoints_per_frame(c->params, c->scales, maxW, maxH);

    /* size every buffer with the geometry of the largest image */
    Geom g; std::vector<uint8_t> tables;
    int rc = compute_geometry(c->params, c->scales, maxW, maxH, g, tables);
    if (rc != SDYN_OK) { delete c; return fail(nullptr, rc, "max_width x max_height is too small for this pyramid"); }
    std::vector<TileRef> ft, bt;
    build_tiles(g, ft, bt);
    const size_t B = (size_t)maxBatch;
    /* smaller images of another aspect ratio can need slightly different per-level splits: 12% headroom */
    auto room = [](size_t v) { return v + v / 8 + 4096; };
    c->pyrFrameCap = room((size_t)g.frameBytes);
    c->cellCap = (int)room(g.cellsPerFrame); c->candCap = (int)room(g.candPerFrame);
    c->levelKpCap = (int)room(g.kpPerFrame); c->tablesCap = room(tables.size());
    c->inFrameCap = (size_t)maxW * maxH;
    const size_t tileCap = room(std::max(ft.size(), bt.size()));

    cudaError_t ce = cudaSuccess;
    auto A = [&](cudaError_t r) { if (ce == cudaSuccess) ce = r; };
    A(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    A(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming)); A(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&c->evFork2, cudaEventDisableTiming)); A(cudaEventCreateWithFlags(&c->evJoin2, cudaEventDisableTiming));
    A(dalloc(&c->dFastTiles, tileCap)); A(dalloc(&c->dBlurTiles, tileCap));
    A(dalloc(&c->dTables, c->tablesCap));
    A(dalloc(&c->dIn, c->inFrameCap * B));
    A(dalloc(&c->dPyr, c->pyrFrameCap * B)); A(dalloc(&c->dBlur, c->pyrFrameCap * B));
    c->tma = new (std::nothrow) TmaMaps();
    if (!c->tma) A(cudaErrorMemoryAllocation);
    A(dalloc(&c->dCellFlag, (size_t)c->cellCap * B));
    A(dalloc(&c->dCand, (size_t)c->candCap * B)); A(dalloc(&c->dCandNode, (size_t)c->candCap * B));
    A(dalloc(&c->dCandCount, SDYN_MAX_LEVELS * B)); A(dalloc(&c->dSelCount, SDYN_MAX_LEVELS * B));
    A(dalloc(&c->dLevelKp, (size_t)c->levelKpCap * B));
    A(dalloc(&c->dLevelCount, SDYN_MAX_LEVELS * B));
    A(dalloc(&c->dCount, B)); A(dalloc(&c->dStatus, B));
    A(dalloc(&c->dKp, (size_t)c->maxKp * B)); A(dalloc(&c->dDesc, (size_t)c->maxKp * 32 * B));
    A(cudaMallocHost(reinterpret_cast<void**>(&c->hKp), sizeof(sdyn_keypoint) * c->maxKp * B));
    A(cudaMallocHost(reinterpret_cast<void**>(&c->hDesc), (size_t)32 * c->maxKp * B));
    A(cudaMallocHost(reinterpret_cast<void**>(&c->hCount), sizeof(int32_t) * B));
    A(cudaMallocHost(reinterpret_cast<void**>(&c->hStatus), sizeof(int32_t) * B));
    if (ce == cudaSuccess) ce = cudaMemset(c->dStatus, 0, sizeof(int32_t) * B);
    if (ce == cudaSuccess) ce = cudaMemset(c->dBlur, 0, c->pyrFrameCap * B);
    if (ce == cudaSuccess) ce = cudaMemset(c->dPyr, 0, c->pyrFrameCap * B);
    if (ce != cudaSuccess) {
        std::string msg = std::string("allocation failed: ") + cudaGetErrorString(ce);
        free_all(c); delete c;
        return fail(nullptr, ce == cudaErrorMemoryAllocation ? SDYN_ERR_NOMEM : SDYN_ERR_CUDA, msg);
    }
    c->geomValid = false; c->profiling = false; std::memset(&c->acc, 0, sizeof(c->acc));
    *out = c;
    return SDYN_OK;
}

int sdyn_destroy(sdyn_ctx* c)
{
    if (!c) return SDYN_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->aux) cudaStreamSynchronize(c->aux);
    free_all(c);
    delete c;
    return SDYN_OK;
}

const char* sdyn_last_error(const sdyn_ctx* c) { return c ? c->err.c_str() : g_createError.c_str(); }

int sdyn_scale_info_get(const sdyn_ctx* c, sdyn_scale_info* out)
{
    if (!c || !out) return SDYN_ERR_ARG;
    *out = c->scales;
    return SDYN_OK;
}

int sdyn_orb_tables(const sdyn_orb_params* p, sdyn_scale_info* out, int32_t umax[16])
{
    if (!p || !out || !umax || p->nlevels < 1 || p->nlevels > SDYN_MAX_LEVELS || !(p->scale_factor > 1.0f)) return SDYN_ERR_ARG;
    int um[16];
    compute_scale_info(*p, *out, um);
    for (int i = 0; i < 16; ++i) umax[i] = um[i];
    return SDYN_OK;
}

int sdyn_set_camera(sdyn_ctx* c, float fx, float fy, float cx, float cy, const float* dist, int ncoef)
{
    if (!c) return SDYN_ERR_ARG;
    if (!(fx > 0.0f) || !(fy > 0.0f) || ncoef < 0 || ncoef > 5 || (ncoef > 0 && !dist))
        return fail(c, SDYN_ERR_ARG, "sdyn_set_camera: bad argument");
    CU(c, cudaSetDevice(c->device));
    CameraModel m; std::memset(&m, 0, sizeof(m));
    m.fx = fx; m.fy = fy; m.cx = cx; m.cy = cy;
    for (int i = 0; i < ncoef; ++i) m.k[i] = dist[i];
    m.enabled = ncoef > 0 && dist[0] != 0.0f;
    if (m.enabled && !c->dKpUn) {
        cudaError_t e = dalloc(&c->dKpUn, (size_t)c->maxKp * c->maxBatch);
        if (e != cudaSuccess) return fail(c, SDYN_ERR_NOMEM, std::string("undistorted keypoints: ") + cudaGetErrorString(e));
    }
    CU(c, cudaStreamSynchronize(c->stream));
    c->camera = m;
    return SDYN_OK;
}

int sdyn_fetch_keypoints_un(sdyn_ctx* c, int nframes, sdyn_keypoint* kpOut, int cap, void* stream)
{
    if (!c) return SDYN_ERR_ARG;
    if (nframes < 1 || nframes > c->maxBatch || !kpOut || cap < 0) return fail(c, SDYN_ERR_ARG, "sdyn_fetch_keypoints_un: bad argument");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const sdyn_keypoint* src = c->camera.enabled ? c->dKpUn : c->dKp;      /* mvKeysUn = mvKeys without distortion */
    const int m = std::min(cap, c->maxKp);
    CU(c, cudaMemcpy2DAsync(kpOut, (size_t)cap * sizeof(sdyn_keypoint), src, (size_t)c->maxKp * sizeof(sdyn_keypoint),
                            (size_t)m * sizeof(sdyn_keypoint), nframes, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    return SDYN_OK;
}

int sdyn_undistort_points(sdyn_ctx* c, const float* xy, int n, float* out)
{
    if (!c) return SDYN_ERR_ARG;
    if (n < 0 || (n > 0 && (!xy || !out))) return fail(c, SDYN_ERR_ARG, "sdyn_undistort_points: bad argument");
    if (n == 0) return SDYN_OK;
    if (!c->camera.enabled) { std::memcpy(out, xy, (size_t)n * 8); return SDYN_OK; }
    CU(c, cudaSetDevice(c->device));
    float* d = nullptr;
    cudaError_t e = dalloc(&d, (size_t)n * 4);
    if (e != cudaSuccess) return fail(c, SDYN_ERR_NOMEM, std::string("sdyn_undistort_points: ") + cudaGetErrorString(e));
    e = cudaMemcpyAsync(d, xy, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = launch_undistort_xy(c->camera, d, n, d + (size_t)2 * n, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + (size_t)2 * n, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(c, e, "sdyn_undistort_points");
    c->launches += 1;
    return SDYN_OK;
}

int sdyn_image_bounds(sdyn_ctx* c, int width, int height, float bounds[4])
{
    if (!c) return SDYN_ERR_ARG;
    if (!bounds || width < 1 || height < 1) return fail(c, SDYN_ERR_ARG, "sdyn_image_bounds: bad argument");
    if (!c->camera.enabled) { bounds[0] = 0.0f; bounds[1] = 0.0f; bounds[2] = (float)width; bounds[3] = (float)height; return SDYN_OK; }
    const float corners[8] = {0.0f, 0.0f, (float)width, 0.0f, 0.0f, (float)height, (float)width, (float)height};
    float u[8];
    int rc = sdyn_undistort_points(c, corners, 4, u);
    if (rc != SDYN_OK) return rc;
    bounds[0] = std::min(u[0], u[4]);      /* mnMinX = min(mat(0,0), mat(2,0)) */
    bounds[2] = std::max(u[2], u[6]);      /* mnMaxX = max(mat(1,0), mat(3,0)) */
    bounds[1] = std::min(u[1], u[3]);      /* mnMinY = min(mat(0,1), mat(1,1)) */
    bounds[3] = std::max(u[5], u[7]);      /* mnMaxY = max(mat(2,1), mat(3,1)) */
    return SDYN_OK;
}

int sdyn_max_keypoints(const sdyn_ctx* c) { return c ? c->maxKp : SDYN_ERR_ARG; }

int sdyn_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return SDYN_ERR_ARG;
    return cudaMallocHost(ptr, bytes) == cudaSuccess ? SDYN_OK : SDYN_ERR_NOMEM;
}

int sdyn_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? SDYN_OK : SDYN_ERR_CUDA; }

int sdyn_extract_batch_device(sdyn_ctx* c, int nframes, const uint8_t* dGray, size_t frameStride,
                              int W, int H, int stride, void* stream)
{
    if (!c) return SDYN_ERR_ARG;
    if (nframes < 1 || nframes > c->maxBatch || !dGray || W < 1 || H < 1 || stride < W)
        return fail(c, SDYN_ERR_ARG, "sdyn_extract_batch_device: bad argument");
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_geometry(c, W, H);
    if (rc != SDYN_OK) return rc;
    return enqueue_extract(c, nframes, dGray, frameStride, stride, stream ? (cudaStream_t)stream : c->stream);
}

int sdyn_fetch_results(sdyn_ctx* c, int nframes, sdyn_keypoint* kpOut, uint8_t* descOut, int cap, int* nOut, void* stream)
{
    if (!c) return SDYN_ERR_ARG;
    if (nframes < 1 || nframes > c->maxBatch || !nOut || cap < 0 || (cap > 0 && (!kpOut || !descOut)))
        return fail(c, SDYN_ERR_ARG, "sdyn_fetch_results: bad argument");
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CU(c, cudaSetDevice(c->device));
    int rc = sdyn::fetch_enqueue(c, nframes, kpOut, descOut, cap, st);
    if (rc != SDYN_OK) return rc;
    CU(c, cudaStreamSynchronize(st));
    return sdyn::fetch_finish(c, nframes, kpOut, descOut, cap, nOut);
}

/* Latency mode (one frame per call, the reference's own execution model): the whole call — H2D of the frame from pinned
 * staging, the 14 kernels on two streams, D2H of counts / keypoints / descriptors into pinned staging — is captured once per
 * image size as a CUDA graph and replayed with one launch; the call then costs one memcpy into the staging buffer, one graph
 * launch, one synchronisation and the copy-out.  Same kernels, same order: outputs are bit-identical to the stream path. */
static int extract_one_graph(sdyn_ctx* c, const uint8_t* gray, int W, int H, int stride, sdyn_keypoint* kpOut, uint8_t* descOut,
                             int cap, int* nOut)
{
    const size_t bytes = (size_t)W * H;
    if (bytes > c->hInCap) {
        CU(c, cudaStreamSynchronize(c->stream));
        cudaFreeHost(c->hIn); c->hIn = nullptr; c->hInCap = 0;
        if (cudaMallocHost(reinterpret_cast<void**>(&c->hIn), (size_t)c->maxW * c->maxH) != cudaSuccess)
            return fail(c, SDYN_ERR_NOMEM, "pinned frame staging");
        c->hInCap = (size_t)c->maxW * c->maxH;
        if (c->graph1) { cudaGraphExecDestroy(c->graph1); c->graph1 = nullptr; }
    }
    if (c->graph1 && (c->graphW != W || c->graphH != H || c->graphCam != c->camera.enabled)) {
        cudaGraphExecDestroy(c->graph1); c->graph1 = nullptr;
    }
    if (!c->graph1) {
        cudaGraph_t g = nullptr;
        CU(c, cudaStreamSynchronize(c->stream));
        CU(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = cudaMemcpyAsync(c->dIn, c->hIn, bytes, cudaMemcpyHostToDevice, c->stream);
        int rc = e == cudaSuccess ? enqueue_extract(c, 1, c->dIn, bytes, W, c->stream) : SDYN_ERR_CUDA;
        if (rc == SDYN_OK) rc = sdyn::fetch_enqueue(c, 1, c->hKp, c->hDesc, c->maxKp > 0 ? c->maxKp - 1 : 0, c->stream);   /* cap != maxKp: stage in pinned memory */
        cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
        if (rc != SDYN_OK || e2 != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            return fail(c, SDYN_ERR_CUDA, std::string("graph capture of the one-frame extraction failed: ") + cudaGetErrorString(e2 != cudaSuccess ? e2 : e));
        }
        e = cudaGraphInstantiate(&c->graph1, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { c->graph1 = nullptr; return cuda_fail(c, e, "cudaGraphInstantiate"); }
        c->graphW = W; c->graphH = H; c->graphCam = c->camera.enabled;
    } else {
        c->launches += 1 + (c->geom.nlevels - 1) + 4 + (c->camera.enabled ? 1 : 0);       /* the kernels one replay launches */
    }
    if (stride == W) std::memcpy(c->hIn, gray, bytes);
    else for (int y = 0; y < H; ++y) std::memcpy(c->hIn + (size_t)y * W, gray + (size_t)y * stride, W);
    CU(c, cudaGraphLaunch(c->graph1, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    /* results are in the pinned staging (hKp / hDesc / hCount / hStatus): scatter like fetch_finish's staged path */
    if (c->hStatus[0]) return fail(c, SDYN_ERR_CAPACITY, "internal candidate/node capacity exceeded");
    const int n = c->hCount[0];
    nOut[0] = n;
    const int m = std::min(n, cap);
    if (m > 0) {
        std::memcpy(kpOut, c->hKp, sizeof(sdyn_keypoint) * (size_t)m);
        std::memcpy(descOut, c->hDesc, (size_t)32 * m);
    }
    if (n > cap) return fail(c, SDYN_ERR_CAPACITY, "output capacity too small (see sdyn_max_keypoints)");
    return SDYN_OK;
}

int sdyn_set_latency_mode(sdyn_ctx* c, int on)
{
    if (!c) return SDYN_ERR_ARG;
    c->latencyMode = on ? 1 : -1;
    return SDYN_OK;
}

int sdyn_extract_batch(sdyn_ctx* c, int nframes, const uint8_t* gray, size_t frameStride, int W, int H, int stride,
                       sdyn_keypoint* kpOut, uint8_t* descOut, int cap, int* nOut)
{
    if (!c) return SDYN_ERR_ARG;
    if (!nOut) return fail(c, SDYN_ERR_ARG, "n_out is null");
    if (!gray || W <= 0 || H <= 0) {               /* empty image: the reference returns silently */
        for (int f = 0; f < std::max(nframes, 0); ++f) nOut[f] = 0;
        return SDYN_OK;
    }
    if (nframes < 1 || nframes > c->maxBatch || stride < W)
        return fail(c, SDYN_ERR_ARG, "sdyn_extract_batch: bad argument");
    CU(c, cudaSetDevice(c->device));
    int rc = ensure_geometry(c, W, H);
    if (rc != SDYN_OK) return rc;
    /* one frame per call on a one-frame context: the graph-captured path (default; sdyn_set_latency_mode(ctx, 0) disables) */
    if (nframes == 1 && c->maxBatch == 1 && c->latencyMode >= 0 && !c->profiling && cap > 0 && kpOut && descOut)
        return extract_one_graph(c, gray, W, H, stride, kpOut, descOut, cap, nOut);
    CU(c, upload_frames(c, nframes, gray, frameStride, W, H, stride, c->stream));
    rc = enqueue_extract(c, nframes, c->dIn, (size_t)W * H, W, c->stream);
    if (rc != SDYN_OK) return rc;
    return sdyn_fetch_results(c, nframes, kpOut, descOut, cap, nOut, nullptr);
}

int sdyn_fetch_level(sdyn_ctx* c, int frame, int level, uint8_t* out, int* width, int* height)
{
    if (!c || !c->geomValid) return SDYN_ERR_ARG;
    if (frame < 0 || frame >= c->maxBatch || level < 0 || level >= c->geom.nlevels || !out)
        return fail(c, SDYN_ERR_ARG, "sdyn_fetch_level: bad argument");
    const LevelGeom& L = c->geom.L[level];
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    const uint8_t* src = c->dPyr + (size_t)frame * c->geom.frameBytes + L.off - (long long)kEdge * L.pitch - kEdge;
    CU(c, cudaMemcpy2D(out, L.w + 2 * kEdge, src, L.pitch, L.w + 2 * kEdge, L.h + 2 * kEdge, cudaMemcpyDeviceToHost));
    if (width) *width = L.w;
    if (height) *height = L.h;
    return SDYN_OK;
}

int sdyn_extract(sdyn_ctx* c, const uint8_t* gray, int W, int H, int stride, sdyn_keypoint* kpOut, uint8_t* descOut,
                 int cap, int* nOut, uint8_t* const* pyrOut)
{
    int rc = sdyn_extract_batch(c, 1, gray, 0, W, H, stride, kpOut, descOut, cap, nOut);
    if ((rc == SDYN_OK || rc == SDYN_ERR_CAPACITY) && pyrOut && gray && W > 0 && H > 0) {
        for (int l = 0; l < c->geom.nlevels; ++l) {
            if (!pyrOut[l]) continue;
            int r2 = sdyn_fetch_level(c, 0, l, pyrOut[l], nullptr, nullptr);
            if (r2 != SDYN_OK) return r2;
        }
    }
    return rc;
}

int sdyn_device_results(const sdyn_ctx* c, sdyn_device_view* out)
{
    if (!c || !out || !c->geomValid) return SDYN_ERR_ARG;
    std::memset(out, 0, sizeof(*out));
    out->kp = c->dKp; out->desc = c->dDesc; out->count = c->dCount; out->level_count = c->dLevelCount;
    out->pyramid = c->dPyr; out->blurred = c->dBlur;
    out->pyramid_frame_bytes = (size_t)c->geom.frameBytes;
    out->cap = c->maxKp; out->nlevels = c->geom.nlevels;
    for (int l = 0; l < c->geom.nlevels; ++l) {
        const LevelGeom& L = c->geom.L[l];
        out->level[l].width = L.w; out->level[l].height = L.h; out->level[l].pitch = L.pitch;
        out->level[l].offset = (size_t)(L.off - (long long)kEdge * L.pitch - kEdge);
    }
    return SDYN_OK;
}

int sdyn_fetch_candidates(sdyn_ctx* c, int frame, int level, int32_t* xyv, int cap, int* nOut)
{
    if (!c || !c->geomValid || !nOut) return SDYN_ERR_ARG;
    if (frame < 0 || frame >= c->maxBatch || level < 0 || level >= c->geom.nlevels)
        return fail(c, SDYN_ERR_ARG, "sdyn_fetch_candidates: bad argument");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    /* the octree kernel applies the per-cell threshold fallback and compacts the survivors to the front of
     * the level's candidate slice; their number is left in dSelCount */
    const LevelGeom& L = c->geom.L[level];
    int32_t n = 0;
    CU(c, cudaMemcpy(&n, c->dSelCount + frame * SDYN_MAX_LEVELS + level, sizeof(int32_t), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> k((size_t)std::max(n, 1));
    CU(c, cudaMemcpy(k.data(), c->dCand + (size_t)frame * c->geom.candPerFrame + L.candOff, sizeof(uint32_t) * n,
                     cudaMemcpyDeviceToHost));
    *nOut = n;
    for (int i = 0; i < std::min(n, cap); ++i) {
        xyv[3 * i] = (int32_t)(k[i] & 4095);
        xyv[3 * i + 1] = (int32_t)((k[i] >> 12) & 4095);
        xyv[3 * i + 2] = (int32_t)(k[i] >> 24);
    }
    return n > cap ? SDYN_ERR_CAPACITY : SDYN_OK;
}

int sdyn_sync(sdyn_ctx* c)
{
    if (!c) return SDYN_ERR_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    return SDYN_OK;
}

long long sdyn_launch_count(const sdyn_ctx* c) { return c ? c->launches : 0; }

int sdyn_profile_enable(sdyn_ctx* c, int on)
{
    if (!c) return SDYN_ERR_ARG;
    cudaSetDevice(c->device);
    drain_spans(c);
    c->profiling = on != 0;
    std::memset(&c->acc, 0, sizeof(c->acc));
    return SDYN_OK;
}

int sdyn_profile_read(sdyn_ctx* c, sdyn_stage_times* out)
{
    if (!c || !out) return SDYN_ERR_ARG;
    cudaSetDevice(c->device);
    drain_spans(c);
    *out = c->acc;
    std::memset(&c->acc, 0, sizeof(c->acc));
    return SDYN_OK;
}

}  // extern "C"
