/* C-ABI of the stereo association: Frame::ComputeStereoMatches (src/Frame.cc:874-1048) on the device-resident
 * results of two extraction contexts (include/sdyn.h, "Frame::ComputeStereoMatches"). */
#include "stereo_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace sdyn {

int api_fail(sdyn_ctx* c, int code, const std::string& msg);

struct StereoState {
    int B = 0, cap = 0, capR = 0, maxRows = 0, listCap = 0;
    uint8_t* block = nullptr;
    int32_t* rowOff; int32_t* rowList; int2* rightKey; int32_t* sad; int32_t* status; float* uRight; float* depth; int32_t* kept;
    float* hU = nullptr; float* hZ = nullptr; int32_t* hKept = nullptr; int32_t* hStatus = nullptr;
    cudaEvent_t evRight = nullptr, evLeft = nullptr;
};

void free_stereo_state(sdyn_ctx* c)
{
    StereoState* s = static_cast<StereoState*>(c->stereo);
    if (!s) return;
    cudaFree(s->block);
    cudaFreeHost(s->hU); cudaFreeHost(s->hZ); cudaFreeHost(s->hKept); cudaFreeHost(s->hStatus);
    if (s->evRight) cudaEventDestroy(s->evRight);
    if (s->evLeft) cudaEventDestroy(s->evLeft);
    delete s;
    c->stereo = nullptr;
}

static int ensure_stereo_state(sdyn_ctx* c, int capR)
{
    StereoState* s = static_cast<StereoState*>(c->stereo);
    if (s && s->capR >= capR) return SDYN_OK;
    if (s) { cudaStreamSynchronize(c->stream); free_stereo_state(c); }
    s = new StereoState();
    s->B = c->maxBatch; s->cap = c->maxKp; s->capR = capR; s->maxRows = c->maxH;
    /* a right keypoint occupies rows floor(y - r) .. ceil(y + r), r = 2 * scale[octave]: at most 2r + 3 rows */
    const float rmax = 2.0f * c->scales.scale[c->scales.nlevels - 1];
    s->listCap = capR * ((int)std::ceil(2.0f * rmax) + 3);
    const size_t B = (size_t)s->B;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t oRowOff = take(B * (s->maxRows + 1) * 4), oRowList = take(B * s->listCap * 4), oRK = take(B * capR * sizeof(int2));
    const size_t oSad = take(B * s->cap * 4), oStatus = take(B * 4), oU = take(B * s->cap * 4), oZ = take(B * s->cap * 4), oKept = take(B * 4);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->block), off);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&s->hU), B * s->cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&s->hZ), B * s->cap * 4);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&s->hKept), B * 4);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&s->hStatus), B * 4);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->evRight, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->evLeft, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        c->stereo = s; free_stereo_state(c);
        return api_fail(c, SDYN_ERR_NOMEM, std::string("stereo state: ") + cudaGetErrorString(e));
    }
    uint8_t* b = s->block;
    s->rowOff = reinterpret_cast<int32_t*>(b + oRowOff); s->rowList = reinterpret_cast<int32_t*>(b + oRowList);
    s->rightKey = reinterpret_cast<int2*>(b + oRK); s->sad = reinterpret_cast<int32_t*>(b + oSad);
    s->status = reinterpret_cast<int32_t*>(b + oStatus); s->uRight = reinterpret_cast<float*>(b + oU);
    s->depth = reinterpret_cast<float*>(b + oZ); s->kept = reinterpret_cast<int32_t*>(b + oKept);
    c->stereo = s;
    return SDYN_OK;
}

}  // namespace sdyn

using namespace sdyn;

#define SCU(c, call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return api_fail((c), SDYN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

int sdyn_stereo_match_device(sdyn_ctx* left, sdyn_ctx* right, int nframes, float mb, float mbf, void* stream)
{
    if (!left) return SDYN_ERR_ARG;
    if (!right || left == right || nframes < 1 || nframes > left->maxBatch || nframes > right->maxBatch || !(mb > 0.0f) ||
        left->device != right->device)
        return api_fail(left, SDYN_ERR_ARG, "sdyn_stereo_match: bad argument");
    if (!left->geomValid || !right->geomValid || left->geom.W != right->geom.W || left->geom.H != right->geom.H ||
        left->geom.nlevels != right->geom.nlevels || left->params.scale_factor != right->params.scale_factor ||
        left->geom.frameBytes != right->geom.frameBytes)
        return api_fail(left, SDYN_ERR_ARG, "sdyn_stereo_match: the two contexts have not extracted frames of the same geometry");
    if (left->maxKp > 65535 || right->maxKp > 65535 || left->geom.H > 4095)
        return api_fail(left, SDYN_ERR_ARG, "sdyn_stereo_match: more than 65535 keypoints or 4095 rows per frame");
    SCU(left, cudaSetDevice(left->device));
    int rc = ensure_stereo_state(left, right->maxKp);
    if (rc != SDYN_OK) return rc;
    StereoState* s = static_cast<StereoState*>(left->stereo);
    cudaStream_t st = stream ? (cudaStream_t)stream : left->stream;
    /* order the step behind both extractions */
    SCU(left, cudaEventRecord(s->evRight, right->stream));
    SCU(left, cudaStreamWaitEvent(st, s->evRight, 0));
    if (st != left->stream) {
        SCU(left, cudaEventRecord(s->evLeft, left->stream));
        SCU(left, cudaStreamWaitEvent(st, s->evLeft, 0));
    }
    StereoArgs a; std::memset(&a, 0, sizeof(a));
    a.keysL = left->dKp; a.descL = left->dDesc; a.countL = left->dCount; a.capL = left->maxKp;
    a.keysR = right->dKp; a.descR = right->dDesc; a.countR = right->dCount; a.capR = right->maxKp;
    a.pyrL = left->dPyr; a.pyrR = right->dPyr;
    a.frameBytesL = (size_t)left->geom.frameBytes; a.frameBytesR = (size_t)right->geom.frameBytes;
    a.nRows = left->geom.H; a.maxRows = s->maxRows;
    for (int l = 0; l < SDYN_MAX_LEVELS; ++l) {
        a.scale[l] = l < left->scales.nlevels ? left->scales.scale[l] : 1.f;
        a.invScale[l] = l < left->scales.nlevels ? left->scales.inv_scale[l] : 1.f;
    }
    a.mb = mb; a.mbf = mbf;
    a.rowOff = s->rowOff; a.rowList = s->rowList; a.listCap = s->listCap; a.rightKey = s->rightKey; a.sad = s->sad;
    a.status = s->status; a.uRight = s->uRight; a.depth = s->depth; a.kept = s->kept;
    {
        StageTimer tm(left, st, SDYN_STAGE_STEREO);
        SCU(left, cudaMemsetAsync(s->status, 0, (size_t)nframes * 4, st));
        SCU(left, launch_stereo(left->geom, a, nframes, left->maxKp, st));
        left->launches += 3;
    }
    /* the right context's next extraction must not overwrite its pyramid while this step reads it */
    SCU(left, cudaEventRecord(s->evLeft, st));
    SCU(left, cudaStreamWaitEvent(right->stream, s->evLeft, 0));
    if (st != left->stream) SCU(left, cudaStreamWaitEvent(left->stream, s->evLeft, 0));
    return SDYN_OK;
}

int sdyn_stereo_results(const sdyn_ctx* left, sdyn_stereo_view* out)
{
    if (!left || !out || !left->stereo) return SDYN_ERR_ARG;
    const StereoState* s = static_cast<const StereoState*>(left->stereo);
    out->u_right = s->uRight; out->depth = s->depth; out->kept = s->kept; out->cap = s->cap;
    return SDYN_OK;
}

int sdyn_stereo_fetch(sdyn_ctx* left, int nframes, float* uRight, float* depth, int cap, int32_t* kept, void* stream)
{
    if (!left) return SDYN_ERR_ARG;
    StereoState* s = static_cast<StereoState*>(left->stereo);
    if (!s || nframes < 1 || nframes > s->B || cap < 0) return api_fail(left, SDYN_ERR_ARG, "sdyn_stereo_fetch: bad argument");
    SCU(left, cudaSetDevice(left->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : left->stream;
    const size_t n = (size_t)nframes * s->cap;
    if (uRight) SCU(left, cudaMemcpyAsync(s->hU, s->uRight, n * 4, cudaMemcpyDeviceToHost, st));
    if (depth) SCU(left, cudaMemcpyAsync(s->hZ, s->depth, n * 4, cudaMemcpyDeviceToHost, st));
    SCU(left, cudaMemcpyAsync(s->hKept, s->kept, (size_t)nframes * 4, cudaMemcpyDeviceToHost, st));
    SCU(left, cudaMemcpyAsync(s->hStatus, s->status, (size_t)nframes * 4, cudaMemcpyDeviceToHost, st));
    SCU(left, cudaStreamSynchronize(st));
    const int m = std::min(cap, s->cap);
    for (int f = 0; f < nframes; ++f) {
        if (uRight) std::memcpy(uRight + (size_t)f * cap, s->hU + (size_t)f * s->cap, (size_t)m * 4);
        if (depth) std::memcpy(depth + (size_t)f * cap, s->hZ + (size_t)f * s->cap, (size_t)m * 4);
        if (kept) kept[f] = s->hKept[f];
    }
    for (int f = 0; f < nframes; ++f)
        if (s->hStatus[f])
            return api_fail(left, SDYN_ERR_GEOMETRY, "sdyn_stereo_match: a right keypoint's row band leaves the image");
    return SDYN_OK;
}

int sdyn_stereo_match(sdyn_ctx* left, sdyn_ctx* right, int nframes, float mb, float mbf, float* uRight, float* depth,
                      int cap, int32_t* kept)
{
    int rc = sdyn_stereo_match_device(left, right, nframes, mb, mbf, nullptr);
    return rc == SDYN_OK ? sdyn_stereo_fetch(left, nframes, uRight, depth, cap, kept, nullptr) : rc;
}

}  // extern "C"
