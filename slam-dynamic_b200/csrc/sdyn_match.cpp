/* C-ABI glue of the ORBmatcher searches and the dynamic-keypoint rejection: stages the caller's host
 * arrays into a per-context device arena, fills the job descriptors, launches k_match.cu / k_dynamic.cu and
 * copies the results back.  No CPU computation of distances or matches happens here. */
#include "match_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>

using namespace sdyn;

namespace sdyn {

int api_fail(sdyn_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }

#define MCU(c, call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return api_fail((c), SDYN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

/* Bump allocator over one growable device block (reset per call). */
struct Arena {
    sdyn_ctx* c; size_t used = 0; bool failed = false;
    explicit Arena(sdyn_ctx* c_) : c(c_) {}
    template <class T> T* take(size_t count)
    {
        const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) / 256 * 256;
        if (used + bytes > c->arenaCap) { failed = true; used += bytes; return nullptr; }
        T* p = reinterpret_cast<T*>(c->dArena + used);
        used += bytes;
        return p;
    }
};

int ensure_arena(sdyn_ctx* c, size_t bytes)
{
    if (bytes <= c->arenaCap) return SDYN_OK;
    MCU(c, cudaStreamSynchronize(c->stream));
    if (c->dArena) cudaFree(c->dArena);
    c->dArena = nullptr; c->arenaCap = 0;
    const size_t cap = bytes + bytes / 2 + (1 << 20);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->dArena), cap);
    if (e != cudaSuccess) return api_fail(c, SDYN_ERR_NOMEM, std::string("matcher arena: ") + cudaGetErrorString(e));
    c->arenaCap = cap;
    return SDYN_OK;
}

template <class T>
cudaError_t up(T* d, const T* h, size_t n, cudaStream_t st)
{ return n ? cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, st) : cudaSuccess; }
template <class T>
cudaError_t down(T* h, const T* d, size_t n, cudaStream_t st)
{ return n ? cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, st) : cudaSuccess; }

/* The searched frame's arrays + grid scratch. */
void stage_frame(Arena& A, const sdyn_frame_view* f, MatchJob& J, bool needGrid)
{
    J.keysUn = A.take<sdyn_keypoint>(f->n);
    J.desc = A.take<uint8_t>((size_t)32 * f->n);
    J.uRight = f->u_right ? A.take<float>(f->n) : nullptr;
    J.n = f->n; J.nPtr = nullptr;
    J.minX = f->min_x; J.minY = f->min_y; J.maxX = f->max_x; J.maxY = f->max_y;
    J.gridWInv = static_cast<float>(SDYN_GRID_COLS) / static_cast<float>(f->max_x - f->min_x);   /* Frame.cc:383-385 */
    J.gridHInv = static_cast<float>(SDYN_GRID_ROWS) / static_cast<float>(f->max_y - f->min_y);
    for (int l = 0; l < SDYN_MAX_LEVELS; ++l) J.scale[l] = (f->scale_factors && l < f->nlevels) ? f->scale_factors[l] : 1.f;
    if (needGrid) {
        J.cellOff = A.take<int32_t>(kGridCells + 1);
        J.sorted = A.take<int32_t>(f->n);
        J.cellOf = A.take<int32_t>(f->n);
        J.gridEntry = A.take<float4>(f->n);
    }
}

cudaError_t upload_frame(const sdyn_frame_view* f, const MatchJob& J, cudaStream_t st)
{
    cudaError_t e = up(const_cast<sdyn_keypoint*>(J.keysUn), f->keys_un, f->n, st);
    if (e == cudaSuccess) e = up(const_cast<uint8_t*>(J.desc), f->desc, (size_t)32 * f->n, st);
    if (e == cudaSuccess && f->u_right) e = up(const_cast<float*>(J.uRight), f->u_right, f->n, st);
    return e;
}

bool bad_view(const sdyn_frame_view* f)
{
    return !f || f->n < 0 || f->n > 65535 || (f->n > 0 && (!f->keys_un || !f->desc)) || f->nlevels < 1 ||
           f->nlevels > SDYN_MAX_LEVELS || !(f->max_x > f->min_x) || !(f->max_y > f->min_y);
}

/* Runs one job with pool-overflow retry.  `fill` (re)builds the job for a given pool capacity. */
template <class Fill, class Fetch>
int run_job(sdyn_ctx* c, size_t fixedBytes, int nq, int poolGuess, int poolMax, Fill fill, Fetch fetch)
{
    MCU(c, cudaSetDevice(c->device));
    int pool = std::max(poolGuess, 4096);
    for (int attempt = 0; attempt < 6; ++attempt) {
        const size_t need = fixedBytes + (size_t)pool * 4 + (size_t)nq * 16 + sizeof(MatchJob) + (64 << 10);
        int rc = ensure_arena(c, need);
        if (rc != SDYN_OK) return rc;
        Arena A(c);
        MatchJob J; std::memset(&J, 0, sizeof(J));
        J.pool = A.take<uint32_t>(pool); J.poolCap = pool;
        J.poolUsed = A.take<int32_t>(2); J.qNext = J.poolUsed + 1;
        J.result = A.take<int32_t>(4);
        J.qspan = A.take<int2>(nq); J.qAccepted = A.take<int32_t>(nq); J.qBin = A.take<int32_t>(nq);
        MatchJob* dJob = A.take<MatchJob>(1);
        rc = fill(A, J);
        if (rc != SDYN_OK) return rc;
        if (A.failed) { rc = ensure_arena(c, A.used + (1 << 20)); if (rc != SDYN_OK) return rc; continue; }
        MCU(c, cudaMemsetAsync(J.poolUsed, 0, 2 * sizeof(int32_t), c->stream));
        MCU(c, cudaMemsetAsync(J.result, 0, 4 * sizeof(int32_t), c->stream));
        MCU(c, cudaMemcpyAsync(dJob, &J, sizeof(J), cudaMemcpyHostToDevice, c->stream));
        MCU(c, launch_grid_build(dJob, 1, std::max(J.n, 1), c->stream));
        MCU(c, launch_match_candidates(dJob, 1, nq, std::max(J.n, 1), c->stream));
        MCU(c, launch_match_resolve(dJob, 1, J.mode, std::max(J.n, 1), std::max(nq, 1), c->stream));
        c->launches += 3;
        int32_t res[4];
        MCU(c, cudaMemcpyAsync(res, J.result, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
        MCU(c, cudaStreamSynchronize(c->stream));
        if (res[2] && pool < poolMax) { pool = (int)std::min<long long>((long long)pool * 4, poolMax); continue; }
        if (res[2]) return api_fail(c, SDYN_ERR_CAPACITY, "matcher candidate pool exhausted");
        c->lastEvals = res[3];
        return fetch(J, res);
    }
    return api_fail(c, SDYN_ERR_CAPACITY, "matcher candidate pool exhausted");
}

}  // namespace sdyn

extern "C" {

long long sdyn_match_last_evals(const sdyn_ctx* c) { return c ? c->lastEvals : 0; }

int sdyn_hamming(const uint8_t* a, const uint8_t* b)
{
    int d = 0;
    for (int i = 0; i < 32; i += 8) {
        uint64_t x, y;
        std::memcpy(&x, a + i, 8); std::memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

int sdyn_match_projection_map(sdyn_ctx* c, const sdyn_frame_view* f, const sdyn_mappoint_query* mps, int nmp, float th,
                              float nnratio, int32_t* assign, uint8_t* locked, int* nmatches)
{
    if (!c) return SDYN_ERR_ARG;
    if (bad_view(f) || nmp < 0 || (nmp > 0 && !mps) || !assign || !locked || !nmatches || !f->scale_factors)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_projection_map: bad argument");
    *nmatches = 0;
    if (nmp == 0 || f->n == 0) return SDYN_OK;
    for (int i = 0; i < nmp; ++i)
        if (mps[i].track_in_view && !mps[i].bad && (mps[i].level < 0 || mps[i].level >= f->nlevels))
            return api_fail(c, SDYN_ERR_ARG, "map point with predicted level outside the pyramid");
    const size_t fixed = (size_t)f->n * 100 + (size_t)nmp * sizeof(sdyn_mappoint_query) + (kGridCells + 1) * 4;
    return run_job(c, fixed, nmp, 64 * nmp, (int)std::min<long long>((long long)nmp * f->n, 1 << 30),
        [&](Arena& A, MatchJob& J) {
            J.mode = MM_MAP; J.distTh = SDYN_TH_HIGH;
            stage_frame(A, f, J, true);
            sdyn_mappoint_query* dq = A.take<sdyn_mappoint_query>(nmp);
            J.queries = dq; J.nq = nmp; J.th = th; J.nnratio = nnratio;
            J.assign = A.take<int32_t>(f->n); J.locked = A.take<uint8_t>(f->n);
            if (A.failed) return (int)SDYN_OK;
            MCU(c, upload_frame(f, J, c->stream));
            MCU(c, up(dq, mps, nmp, c->stream));
            MCU(c, up(J.assign, assign, f->n, c->stream));
            MCU(c, up(J.locked, locked, f->n, c->stream));
            return (int)SDYN_OK;
        },
        [&](const MatchJob& J, const int32_t* res) {
            MCU(c, down(assign, J.assign, f->n, c->stream));
            MCU(c, down(locked, J.locked, f->n, c->stream));
            MCU(c, cudaStreamSynchronize(c->stream));
            *nmatches = res[0];
            return (int)SDYN_OK;
        });
}

int sdyn_match_projection_frame(sdyn_ctx* c, const sdyn_frame_view* cur, const sdyn_frame_view* last,
                                const sdyn_last_point* lp, float th, int mono, int checkOri, int32_t* assign,
                                uint8_t* locked, int* nmatches, float* pairs, int* npairs)
{
    if (!c) return SDYN_ERR_ARG;
    if (bad_view(cur) || !last || last->n < 0 || (last->n > 0 && (!lp || !last->keys || !last->keys_un)) || !assign ||
        !locked || !nmatches || !cur->scale_factors || (pairs && !npairs))
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_projection_frame: bad argument");
    *nmatches = 0;
    if (npairs) *npairs = 0;
    if (last->n == 0 || cur->n == 0) return SDYN_OK;
    for (int i = 0; i < last->n; ++i)
        if (last->keys[i].octave < 0 || last->keys[i].octave >= cur->nlevels)
            return api_fail(c, SDYN_ERR_ARG, "last-frame keypoint octave outside the pyramid");
    /* bForward / bBackward (ORBmatcher.cc:1497-1506): twc = -Rcw^T tcw, tlc = Rlw twc + tlw */
    float twc[3], tlc[3];
    for (int r = 0; r < 3; ++r) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (double)cur->tcw[4 * k + r] * (double)cur->tcw[4 * k + 3];
        twc[r] = (float)(-1.0 * s);
    }
    for (int r = 0; r < 3; ++r) {
        const float* T = last->tcw + 4 * r;
        const float s = T[0] * twc[0] + T[1] * twc[1] + T[2] * twc[2];
        tlc[r] = s + T[3];
    }
    const int forward = (tlc[2] > cur->b) && !mono, backward = (-tlc[2] > cur->b) && !mono;
    const int nq = last->n;
    const size_t fixed = (size_t)cur->n * 100 + (size_t)nq * (sizeof(sdyn_last_point) + 2 * sizeof(sdyn_keypoint) + 16) +
                         (kGridCells + 1) * 4;
    return run_job(c, fixed, nq, 64 * nq, (int)std::min<long long>((long long)nq * cur->n, 1 << 30),
        [&](Arena& A, MatchJob& J) {
            J.mode = MM_FRAME; J.distTh = SDYN_TH_HIGH;
            stage_frame(A, cur, J, true);
            sdyn_last_point* dq = A.take<sdyn_last_point>(nq);
            sdyn_keypoint* dk = A.take<sdyn_keypoint>(nq);
            sdyn_keypoint* dku = A.take<sdyn_keypoint>(nq);
            J.queries = dq; J.qKeys = dk; J.qKeysUn = dku; J.nq = nq; J.th = th; J.checkOri = checkOri;
            J.forward = forward; J.backward = backward;
            std::memcpy(J.Tcw, cur->tcw, sizeof(J.Tcw));
            J.fx = cur->fx; J.fy = cur->fy; J.cx = cur->cx; J.cy = cur->cy; J.bf = cur->bf;
            J.assign = A.take<int32_t>(cur->n); J.locked = A.take<uint8_t>(cur->n);
            J.pairs = pairs ? A.take<float>((size_t)4 * nq) : nullptr;
            if (A.failed) return (int)SDYN_OK;
            MCU(c, upload_frame(cur, J, c->stream));
            MCU(c, up(dq, lp, nq, c->stream));
            MCU(c, up(dk, last->keys, nq, c->stream));
            MCU(c, up(dku, last->keys_un, nq, c->stream));
            MCU(c, up(J.assign, assign, cur->n, c->stream));
            MCU(c, up(J.locked, locked, cur->n, c->stream));
            return (int)SDYN_OK;
        },
        [&](const MatchJob& J, const int32_t* res) {
            MCU(c, down(assign, J.assign, cur->n, c->stream));
            MCU(c, down(locked, J.locked, cur->n, c->stream));
            if (pairs) MCU(c, down(pairs, J.pairs, (size_t)4 * res[1], c->stream));
            MCU(c, cudaStreamSynchronize(c->stream));
            *nmatches = res[0];
            if (npairs) *npairs = res[1];
            return (int)SDYN_OK;
        });
}

int sdyn_match_projection_pose(sdyn_ctx* c, const sdyn_frame_view* target, const sdyn_proj_point* pts, int npts,
                               const sdyn_proj_params* prm, int32_t* assign, int* nmatches)
{
    if (!c) return SDYN_ERR_ARG;
    if (bad_view(target) || npts < 0 || (npts > 0 && !pts) || !prm || !assign || !nmatches || !target->scale_factors ||
        (prm->variant != SDYN_PROJ_FRAME_KEYFRAME && prm->variant != SDYN_PROJ_KEYFRAME_SIM3) || prm->nlevels < 1 ||
        prm->nlevels > target->nlevels || !(prm->log_scale_factor > 0.0f) || prm->max_descriptor_distance < 0)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_projection_pose: bad argument");
    *nmatches = 0;
    if (npts == 0 || target->n == 0) return SDYN_OK;
    std::vector<uint8_t> locked((size_t)target->n);
    for (int i = 0; i < target->n; ++i) locked[i] = assign[i] != -1;      /* every occupant blocks (:362, :1701) */
    const size_t fixed = (size_t)target->n * 100 + (size_t)npts * (sizeof(sdyn_proj_point) + 16) + (kGridCells + 1) * 4;
    return run_job(c, fixed, npts, 64 * npts, (int)std::min<long long>((long long)npts * target->n, 1 << 30),
        [&](Arena& A, MatchJob& J) {
            J.mode = MM_POSE; J.poseVariant = prm->variant; J.distTh = prm->max_descriptor_distance;
            stage_frame(A, target, J, true);
            sdyn_proj_point* dq = A.take<sdyn_proj_point>(npts);
            J.queries = dq; J.nq = npts; J.th = prm->th;
            J.checkOri = prm->variant == SDYN_PROJ_FRAME_KEYFRAME && prm->check_orientation;
            for (int r = 0; r < 3; ++r) {
                for (int k = 0; k < 3; ++k) J.Tcw[4 * r + k] = prm->rcw[3 * r + k];
                J.Tcw[4 * r + 3] = prm->tcw[r];
                J.Ow[r] = prm->ow[r];
            }
            J.fx = target->fx; J.fy = target->fy; J.cx = target->cx; J.cy = target->cy;
            J.logScaleFactor = prm->log_scale_factor; J.predLevels = prm->nlevels;
            J.assign = A.take<int32_t>(target->n); J.locked = A.take<uint8_t>(target->n);
            if (A.failed) return (int)SDYN_OK;
            MCU(c, upload_frame(target, J, c->stream));
            MCU(c, up(dq, pts, npts, c->stream));
            MCU(c, up(J.assign, assign, target->n, c->stream));
            MCU(c, up(J.locked, locked.data(), target->n, c->stream));
            return (int)SDYN_OK;
        },
        [&](const MatchJob& J, const int32_t* res) {
            MCU(c, down(assign, J.assign, target->n, c->stream));
            MCU(c, cudaStreamSynchronize(c->stream));
            *nmatches = res[0];
            return (int)SDYN_OK;
        });
}

int sdyn_match_projection_best(sdyn_ctx* c, const sdyn_frame_view* target, const sdyn_proj_point* pts, int npts,
                               const sdyn_best_params* prm, int32_t* bestIdx, int32_t* bestDist)
{
    if (!c) return SDYN_ERR_ARG;
    if (bad_view(target) || npts < 0 || (npts > 0 && (!pts || !bestIdx || !bestDist)) || !prm || !target->scale_factors ||
        prm->nlevels < 1 || prm->nlevels > target->nlevels || !(prm->log_scale_factor > 0.0f))
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_projection_best: bad argument");
    for (int i = 0; i < npts; ++i) { bestIdx[i] = -1; bestDist[i] = 256; }
    if (npts == 0 || target->n == 0) return SDYN_OK;
    const size_t fixed = (size_t)target->n * 100 + (size_t)npts * (sizeof(sdyn_proj_point) + 16) + (kGridCells + 1) * 4;
    MCU(c, cudaSetDevice(c->device));
    int pool = 64 * npts;
    const int poolMax = (int)std::min<long long>((long long)npts * target->n, 1 << 30);
    for (int attempt = 0; attempt < 8; ++attempt) {
        int rc = ensure_arena(c, fixed + (size_t)pool * 4 + (size_t)npts * 16 + sizeof(MatchJob) + (64 << 10));
        if (rc != SDYN_OK) return rc;
        Arena A(c);
        MatchJob J; std::memset(&J, 0, sizeof(J));
        J.mode = MM_BEST;
        J.pool = A.take<uint32_t>(pool); J.poolCap = pool;
        J.poolUsed = A.take<int32_t>(2); J.qNext = J.poolUsed + 1;
        J.result = A.take<int32_t>(4);
        J.qspan = A.take<int2>(npts); J.qAccepted = A.take<int32_t>(npts); J.qBin = A.take<int32_t>(npts);
        MatchJob* dJob = A.take<MatchJob>(1);
        stage_frame(A, target, J, true);
        sdyn_proj_point* dq = A.take<sdyn_proj_point>(npts);
        J.queries = dq; J.nq = npts; J.th = prm->th;
        std::memcpy(J.Tcw, prm->t1, sizeof(J.Tcw)); std::memcpy(J.T2, prm->t2, sizeof(J.T2));
        J.useT2 = prm->use_t2; J.invzDouble = prm->invz_double; J.distFromCamera = prm->dist_from_camera;
        J.checkNormal = prm->check_normal; J.chi2Gate = prm->chi2_gate;
        for (int k = 0; k < 3; ++k) J.Ow[k] = prm->ow[k];
        for (int l = 0; l < SDYN_MAX_LEVELS; ++l) J.invSigma2[l] = prm->inv_level_sigma2[l];
        J.fx = target->fx; J.fy = target->fy; J.cx = target->cx; J.cy = target->cy; J.bf = prm->bf;
        J.logScaleFactor = prm->log_scale_factor; J.predLevels = prm->nlevels;
        if (A.failed) { rc = ensure_arena(c, A.used + (1 << 20)); if (rc != SDYN_OK) return rc; continue; }
        MCU(c, cudaMemsetAsync(J.poolUsed, 0, 2 * sizeof(int32_t), c->stream));
        MCU(c, cudaMemsetAsync(J.result, 0, 4 * sizeof(int32_t), c->stream));
        MCU(c, upload_frame(target, J, c->stream));
        MCU(c, up(dq, pts, npts, c->stream));
        MCU(c, cudaMemcpyAsync(dJob, &J, sizeof(J), cudaMemcpyHostToDevice, c->stream));
        MCU(c, launch_grid_build(dJob, 1, std::max(J.n, 1), c->stream));
        MCU(c, launch_match_candidates(dJob, 1, npts, std::max(J.n, 1), c->stream));
        c->launches += 2;
        int32_t res[4];
        MCU(c, cudaMemcpyAsync(res, J.result, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
        MCU(c, down(bestIdx, J.qAccepted, npts, c->stream));
        MCU(c, down(bestDist, J.qBin, npts, c->stream));
        MCU(c, cudaStreamSynchronize(c->stream));
        if (res[2] && pool < poolMax) { pool = (int)std::min<long long>((long long)pool * 4, poolMax); continue; }
        if (res[2]) return api_fail(c, SDYN_ERR_CAPACITY, "matcher candidate pool exhausted");
        return SDYN_OK;
    }
    return api_fail(c, SDYN_ERR_CAPACITY, "matcher candidate pool exhausted");
}

int sdyn_match_init(sdyn_ctx* c, const sdyn_frame_view* f1, const sdyn_frame_view* f2, float* prevMatched,
                    int32_t* matches12, int window, float nnratio, int checkOri, int* nmatches)
{
    if (!c) return SDYN_ERR_ARG;
    if (bad_view(f2) || !f1 || f1->n < 0 || (f1->n > 0 && (!f1->keys_un || !f1->desc || !prevMatched || !matches12)) || !nmatches)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_init: bad argument");
    *nmatches = 0;
    for (int i = 0; i < f1->n; ++i) matches12[i] = -1;
    if (f1->n == 0 || f2->n == 0) return SDYN_OK;
    const int nq = f1->n;
    const size_t fixed = (size_t)f2->n * 110 + (size_t)nq * 80 + (kGridCells + 1) * 4;
    return run_job(c, fixed, nq, 512 * nq, (int)std::min<long long>((long long)nq * f2->n, 1 << 30),
        [&](Arena& A, MatchJob& J) {
            J.mode = MM_INIT;
            stage_frame(A, f2, J, true);
            sdyn_keypoint* dk = A.take<sdyn_keypoint>(nq);
            uint8_t* dd = A.take<uint8_t>((size_t)32 * nq);
            J.qKeys = dk; J.qDesc = dd; J.nq = nq; J.window = window; J.nnratio = nnratio; J.checkOri = checkOri;
            J.prevMatched = A.take<float>((size_t)2 * nq);
            J.assign = A.take<int32_t>(nq);
            J.matchedDist = A.take<int32_t>(f2->n); J.m21 = A.take<int32_t>(f2->n);
            if (A.failed) return (int)SDYN_OK;
            MCU(c, upload_frame(f2, J, c->stream));
            MCU(c, up(dk, f1->keys_un, nq, c->stream));
            MCU(c, up(dd, f1->desc, (size_t)32 * nq, c->stream));
            MCU(c, up(J.prevMatched, prevMatched, (size_t)2 * nq, c->stream));
            MCU(c, cudaMemsetAsync(J.assign, 0xff, sizeof(int32_t) * nq, c->stream));            /* -1 */
            MCU(c, cudaMemsetAsync(J.m21, 0xff, sizeof(int32_t) * f2->n, c->stream));
            MCU(c, cudaMemsetAsync(J.matchedDist, 0x7f, sizeof(int32_t) * f2->n, c->stream));    /* 0x7f7f7f7f > 256 */
            return (int)SDYN_OK;
        },
        [&](const MatchJob& J, const int32_t* res) {
            MCU(c, down(matches12, J.assign, nq, c->stream));
            MCU(c, down(prevMatched, J.prevMatched, (size_t)2 * nq, c->stream));
            MCU(c, cudaStreamSynchronize(c->stream));
            *nmatches = res[0];
            return (int)SDYN_OK;
        });
}

/* Shared body of the two BoW searches.  fValid (nullable): searched features that may be matched at all; strictLow:
 * the KeyFrame-KeyFrame overload accepts bestDist < TH_LOW, the KeyFrame-Frame one bestDist <= TH_LOW. */
static int bow_search(sdyn_ctx* c, const sdyn_frame_view* kf, const uint8_t* kfValid, const sdyn_feature_vector* a,
                      const sdyn_frame_view* f, const uint8_t* fValid, const sdyn_feature_vector* b, float nnratio, int checkOri,
                      int strictLow, int32_t* assign, int* nmatches)
{
    if (!c) return SDYN_ERR_ARG;
    if (!kf || !f || !a || !b || !assign || !nmatches || kf->n < 0 || f->n < 0 || f->n > 65535 || (kf->n > 0 && (!kfValid || !kf->desc || !kf->keys_un)) ||
        (f->n > 0 && (!f->desc || !f->keys_un)))
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_bow: bad argument");
    *nmatches = 0;
    for (int i = 0; i < f->n; ++i) assign[i] = (!fValid || fValid[i]) ? -1 : -3;      /* -3: never a candidate */
    /* merge-join of the two feature vectors on node id (ORBmatcher.cc:182-253) -> one query per valid
     * keyframe feature of every shared node, in reference order */
    std::vector<BowQuery> qs;
    long long pool = 0;
    int ia = 0, ib = 0;
    while (ia < a->nnodes && ib < b->nnodes) {
        if (a->node_id[ia] == b->node_id[ib]) {
            const int fo = b->offset[ib], fc = b->offset[ib + 1] - fo;
            for (int p = a->offset[ia]; p < a->offset[ia + 1]; ++p) {
                const uint32_t k = a->index[p];
                if ((int)k >= kf->n) return api_fail(c, SDYN_ERR_ARG, "feature vector index out of range");
                if (!kfValid[k]) continue;
                qs.push_back({(int32_t)k, fo, fc});
                pool += fc;
            }
            ++ia; ++ib;
        } else if (a->node_id[ia] < b->node_id[ib]) {
            ia = (int)(std::lower_bound(a->node_id, a->node_id + a->nnodes, b->node_id[ib]) - a->node_id);
        } else {
            ib = (int)(std::lower_bound(b->node_id, b->node_id + b->nnodes, a->node_id[ia]) - b->node_id);
        }
    }
    const int nq = (int)qs.size();
    if (nq == 0 || f->n == 0) return SDYN_OK;
    const int nIndex = b->offset[b->nnodes];
    for (int i = 0; i < nIndex; ++i)
        if ((int)b->index[i] >= f->n) return api_fail(c, SDYN_ERR_ARG, "feature vector index out of range");
    if (pool > (1ll << 30)) return api_fail(c, SDYN_ERR_CAPACITY, "BoW search too large");
    const size_t fixed = (size_t)f->n * 100 + (size_t)kf->n * 64 + (size_t)nq * 12 + (size_t)nIndex * 4;
    return run_job(c, fixed, nq, (int)pool + 64, (int)pool + 64,
        [&](Arena& A, MatchJob& J) {
            J.mode = MM_BOW;
            sdyn_frame_view fv = *f;
            if (!(fv.max_x > fv.min_x)) { fv.min_x = 0; fv.max_x = 1; }     /* the BoW search does not use the grid */
            if (!(fv.max_y > fv.min_y)) { fv.min_y = 0; fv.max_y = 1; }
            stage_frame(A, &fv, J, false);
            BowQuery* dq = A.take<BowQuery>(nq);
            sdyn_keypoint* dk = A.take<sdyn_keypoint>(kf->n);
            uint8_t* dd = A.take<uint8_t>((size_t)32 * kf->n);
            uint32_t* di = A.take<uint32_t>(nIndex);
            J.queries = dq; J.qKeys = dk; J.qDesc = dd; J.fIndex = di; J.nq = nq; J.nnratio = nnratio; J.checkOri = checkOri; J.strictLow = strictLow;
            J.assign = A.take<int32_t>(f->n);
            if (A.failed) return (int)SDYN_OK;
            MCU(c, upload_frame(&fv, J, c->stream));
            MCU(c, up(dq, qs.data(), nq, c->stream));
            MCU(c, up(dk, kf->keys_un, kf->n, c->stream));
            MCU(c, up(dd, kf->desc, (size_t)32 * kf->n, c->stream));
            MCU(c, up(di, b->index, nIndex, c->stream));
            MCU(c, up(J.assign, assign, f->n, c->stream));
            return (int)SDYN_OK;
        },
        [&](const MatchJob& J, const int32_t* res) {
            MCU(c, down(assign, J.assign, f->n, c->stream));
            MCU(c, cudaStreamSynchronize(c->stream));
            *nmatches = res[0];
            return (int)SDYN_OK;
        });
}

int sdyn_match_bow(sdyn_ctx* c, const sdyn_frame_view* kf, const uint8_t* kfValid, const sdyn_feature_vector* a,
                   const sdyn_frame_view* f, const sdyn_feature_vector* b, float nnratio, int checkOri, int32_t* assign,
                   int* nmatches)
{
    return bow_search(c, kf, kfValid, a, f, nullptr, b, nnratio, checkOri, 0, assign, nmatches);
}

int sdyn_match_bow_kf(sdyn_ctx* c, const sdyn_frame_view* kf1, const uint8_t* valid1, const sdyn_feature_vector* fv1,
                      const sdyn_frame_view* kf2, const uint8_t* valid2, const sdyn_feature_vector* fv2, float nnratio,
                      int checkOri, int32_t* matches12, int* nmatches)
{
    if (!c) return SDYN_ERR_ARG;
    if (!kf1 || !kf2 || !matches12 || kf1->n < 0 || kf2->n < 0 || (kf2->n > 0 && !valid2))
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_bow_kf: bad argument");
    for (int i = 0; i < kf1->n; ++i) matches12[i] = -1;
    std::vector<int32_t> a2((size_t)std::max(kf2->n, 1));
    int rc = bow_search(c, kf1, valid1, fv1, kf2, valid2, fv2, nnratio, checkOri, 1, a2.data(), nmatches);
    if (rc != SDYN_OK) return rc;
    /* the search records, per KeyFrame-2 feature, the KeyFrame-1 feature that took it: invert into vpMatches12 */
    for (int i2 = 0; i2 < kf2->n; ++i2)
        if (a2[i2] >= 0) matches12[a2[i2]] = i2;
    return SDYN_OK;
}

int sdyn_match_triangulation(sdyn_ctx* c, const sdyn_frame_view* kf1, const uint8_t* hasMp1, const sdyn_feature_vector* a,
                             const sdyn_frame_view* kf2, const uint8_t* hasMp2, const sdyn_feature_vector* b,
                             const sdyn_tri_params* prm, int32_t* matches12, int* nmatches)
{
    if (!c) return SDYN_ERR_ARG;
    if (!kf1 || !kf2 || !a || !b || !prm || !matches12 || !nmatches || kf1->n < 0 || kf2->n < 0 || kf2->n > 65535 ||
        (kf1->n > 0 && (!hasMp1 || !kf1->desc || !kf1->keys_un)) || (kf2->n > 0 && (!hasMp2 || !kf2->desc || !kf2->keys_un)) ||
        !kf2->scale_factors || kf2->nlevels < 1 || kf2->nlevels > SDYN_MAX_LEVELS)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_match_triangulation: bad argument");
    *nmatches = 0;
    for (int i = 0; i < kf1->n; ++i) matches12[i] = -1;
    /* merge-join on node id (:850-946): one query per KeyFrame-1 feature without a MapPoint (and stereo if bOnlyStereo) */
    std::vector<BowQuery> qs;
    int ia = 0, ib = 0;
    while (ia < a->nnodes && ib < b->nnodes) {
        if (a->node_id[ia] == b->node_id[ib]) {
            const int fo = b->offset[ib], fc = b->offset[ib + 1] - fo;
            for (int p = a->offset[ia]; p < a->offset[ia + 1]; ++p) {
                const uint32_t k = a->index[p];
                if ((int)k >= kf1->n) return api_fail(c, SDYN_ERR_ARG, "feature vector index out of range");
                if (hasMp1[k]) continue;
                if (prm->only_stereo && !(kf1->u_right && kf1->u_right[k] >= 0)) continue;
                qs.push_back({(int32_t)k, fo, fc});
            }
            ++ia; ++ib;
        } else if (a->node_id[ia] < b->node_id[ib]) {
            ia = (int)(std::lower_bound(a->node_id, a->node_id + a->nnodes, b->node_id[ib]) - a->node_id);
        } else {
            ib = (int)(std::lower_bound(b->node_id, b->node_id + b->nnodes, a->node_id[ia]) - b->node_id);
        }
    }
    const int nq = (int)qs.size();
    if (nq == 0 || kf2->n == 0) return SDYN_OK;
    const int nIndex = b->offset[b->nnodes];
    for (int i = 0; i < nIndex; ++i)
        if ((int)b->index[i] >= kf2->n) return api_fail(c, SDYN_ERR_ARG, "feature vector index out of range");
    std::vector<uint8_t> fValid((size_t)kf2->n);
    for (int i = 0; i < kf2->n; ++i)
        fValid[i] = !hasMp2[i] && (!prm->only_stereo || (kf2->u_right && kf2->u_right[i] >= 0));
    MCU(c, cudaSetDevice(c->device));
    const size_t need = (size_t)kf2->n * 110 + (size_t)kf1->n * 70 + (size_t)nq * 24 + (size_t)nIndex * 4 + sizeof(MatchJob) + (128 << 10);
    std::vector<int32_t> acc((size_t)nq), bin((size_t)nq);
    for (int attempt = 0; attempt < 4; ++attempt) {
        int rc = ensure_arena(c, need << attempt);
        if (rc != SDYN_OK) return rc;
        Arena A(c);
        MatchJob J; std::memset(&J, 0, sizeof(J));
        J.mode = MM_TRI;
        J.pool = A.take<uint32_t>(64); J.poolCap = 64;
        J.poolUsed = A.take<int32_t>(2); J.qNext = J.poolUsed + 1;
        J.result = A.take<int32_t>(4);
        J.qspan = A.take<int2>(nq); J.qAccepted = A.take<int32_t>(nq); J.qBin = A.take<int32_t>(nq);
        MatchJob* dJob = A.take<MatchJob>(1);
        sdyn_frame_view fv = *kf2;
        if (!(fv.max_x > fv.min_x)) { fv.min_x = 0; fv.max_x = 1; }
        if (!(fv.max_y > fv.min_y)) { fv.min_y = 0; fv.max_y = 1; }
        stage_frame(A, &fv, J, false);
        BowQuery* dq = A.take<BowQuery>(nq);
        sdyn_keypoint* dk = A.take<sdyn_keypoint>(kf1->n);
        uint8_t* dd = A.take<uint8_t>((size_t)32 * kf1->n);
        uint32_t* di = A.take<uint32_t>(nIndex);
        uint8_t* dv = A.take<uint8_t>(kf2->n);
        float* dur = kf1->u_right ? A.take<float>(kf1->n) : nullptr;
        if (A.failed) continue;
        J.queries = dq; J.qKeys = dk; J.qDesc = dd; J.fIndex = di; J.nq = nq; J.fValid = dv; J.qURight = dur;
        std::memcpy(J.F12, prm->f12, sizeof(J.F12)); J.epiX = prm->epipole_x; J.epiY = prm->epipole_y;
        for (int l = 0; l < SDYN_MAX_LEVELS; ++l) J.sigma2[l] = prm->level_sigma2[l];
        MCU(c, cudaMemsetAsync(J.poolUsed, 0, 2 * sizeof(int32_t), c->stream));
        MCU(c, cudaMemsetAsync(J.result, 0, 4 * sizeof(int32_t), c->stream));
        MCU(c, upload_frame(&fv, J, c->stream));
        MCU(c, up(dq, qs.data(), nq, c->stream));
        MCU(c, up(dk, kf1->keys_un, kf1->n, c->stream));
        MCU(c, up(dd, kf1->desc, (size_t)32 * kf1->n, c->stream));
        MCU(c, up(di, b->index, nIndex, c->stream));
        MCU(c, up(dv, fValid.data(), kf2->n, c->stream));
        if (dur) MCU(c, up(dur, kf1->u_right, kf1->n, c->stream));
        MCU(c, cudaMemcpyAsync(dJob, &J, sizeof(J), cudaMemcpyHostToDevice, c->stream));
        MCU(c, launch_match_candidates(dJob, 1, nq, std::max(J.n, 1), c->stream));
        c->launches += 1;
        MCU(c, down(acc.data(), J.qAccepted, nq, c->stream));
        MCU(c, down(bin.data(), J.qBin, nq, c->stream));
        MCU(c, cudaStreamSynchronize(c->stream));
        /* rotation histogram and its three-maxima cull (:920-965) over at most N matches: host code */
        int n = 0;
        int hist[SDYN_HISTO_LENGTH] = {0};
        for (int q = 0; q < nq; ++q)
            if (acc[q] >= 0) { matches12[qs[q].kfIdx] = acc[q]; ++n; ++hist[bin[q]]; }
        if (prm->check_orientation) {
            int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < SDYN_HISTO_LENGTH; ++i) {
                const int sz = hist[i];
                if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; i3 = i2; i2 = i1; i1 = i; }
                else if (sz > max2) { max3 = max2; max2 = sz; i3 = i2; i2 = i; }
                else if (sz > max3) { max3 = sz; i3 = i; }
            }
            if ((float)max2 < 0.1f * (float)max1) { i2 = -1; i3 = -1; }
            else if ((float)max3 < 0.1f * (float)max1) i3 = -1;
            for (int q = 0; q < nq; ++q)
                if (acc[q] >= 0 && bin[q] != i1 && bin[q] != i2 && bin[q] != i3) { matches12[qs[q].kfIdx] = -1; --n; }
        }
        *nmatches = n;
        return SDYN_OK;
    }
    return api_fail(c, SDYN_ERR_NOMEM, "sdyn_match_triangulation: arena");
}

int sdyn_dyn_box_mask(sdyn_ctx* c, const sdyn_keypoint* keys, int n, const double* boxes, int nboxes, uint64_t* mask)
{
    if (!c) return SDYN_ERR_ARG;
    if (n < 0 || nboxes < 0 || nboxes > 64 || (n > 0 && (!keys || !mask)) || (nboxes > 0 && !boxes))
        return api_fail(c, SDYN_ERR_ARG, "sdyn_dyn_box_mask: bad argument (at most 64 boxes)");
    if (n == 0) return SDYN_OK;
    MCU(c, cudaSetDevice(c->device));
    int rc = ensure_arena(c, (size_t)n * 40 + 64 * 32 + 4096);
    if (rc != SDYN_OK) return rc;
    Arena A(c);
    sdyn_keypoint* dk = A.take<sdyn_keypoint>(n);
    double* db = A.take<double>(64 * 4);
    uint64_t* dm = A.take<uint64_t>(n);
    MCU(c, up(dk, keys, n, c->stream));
    MCU(c, up(db, boxes, (size_t)4 * nboxes, c->stream));
    MCU(c, launch_box_mask(dk, nullptr, n, n, db, nullptr, nboxes, 64, dm, 1, c->stream));
    c->launches += 1;
    MCU(c, down(mask, dm, n, c->stream));
    MCU(c, cudaStreamSynchronize(c->stream));
    return SDYN_OK;
}

int sdyn_dyn_separate(sdyn_ctx* c, sdyn_box_pair* pairs, int npairs, const float* M, int mode)
{
    if (!c) return SDYN_ERR_ARG;
    if (npairs < 0 || (npairs > 0 && (!pairs || !M)) || (mode != 0 && mode != 1))
        return api_fail(c, SDYN_ERR_ARG, "sdyn_dyn_separate: bad argument");
    if (npairs == 0) return SDYN_OK;
    size_t need = 4096 + (size_t)npairs * sizeof(BoxPairJob);
    for (int p = 0; p < npairs; ++p) {
        const sdyn_box_pair& P = pairs[p];
        if (P.nq < 0 || P.nt < 0 || (P.nq > 0 && (!P.q_desc || !P.q_xy || !P.match_query || !P.match_train || !P.match_dist || !P.false_dyn)) ||
            (P.nt > 0 && (!P.t_desc || !P.t_xy)))
            return api_fail(c, SDYN_ERR_ARG, "sdyn_dyn_separate: bad pair");
        need += (size_t)(P.nq + P.nt) * (32 + 8 + 12) + (size_t)P.nq * 16 + 4096;
    }
    MCU(c, cudaSetDevice(c->device));
    int rc = ensure_arena(c, need);
    if (rc != SDYN_OK) return rc;
    Arena A(c);
    float Minv[9] = {0};
    if (mode == 1) {       /* H12 = H21.inv(): cv::invert's closed form for 3x3 CV_32F (double arithmetic) */
        const double a = M[0], b = M[1], cc = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
        double det = a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g);
        if (det != 0) {
            det = 1. / det;
            Minv[0] = (float)((e * i - f * h) * det); Minv[1] = (float)((cc * h - b * i) * det); Minv[2] = (float)((b * f - cc * e) * det);
            Minv[3] = (float)((f * g - d * i) * det); Minv[4] = (float)((a * i - cc * g) * det); Minv[5] = (float)((cc * d - a * f) * det);
            Minv[6] = (float)((d * h - e * g) * det); Minv[7] = (float)((b * g - a * h) * det); Minv[8] = (float)((a * e - b * d) * det);
        }
    }
    float* dM = A.take<float>(9); float* dMi = A.take<float>(9);
    BoxPairJob* dJobs = A.take<BoxPairJob>(npairs);
    std::vector<BoxPairJob> jobs(npairs);
    for (int p = 0; p < npairs; ++p) {
        const sdyn_box_pair& P = pairs[p];
        BoxPairJob& J = jobs[p];
        J.nq = P.nq; J.nt = P.nt;
        uint8_t* qd = A.take<uint8_t>((size_t)32 * P.nq); uint8_t* td = A.take<uint8_t>((size_t)32 * P.nt);
        float* qx = A.take<float>((size_t)2 * P.nq); float* tx = A.take<float>((size_t)2 * P.nt);
        J.qDesc = qd; J.tDesc = td; J.qXY = qx; J.tXY = tx;
        J.nnQ = A.take<int32_t>(P.nq); J.dQ = A.take<int32_t>(P.nq); J.nnT = A.take<int32_t>(P.nt);
        J.outQuery = A.take<int32_t>(P.nq); J.outTrain = A.take<int32_t>(P.nq); J.outDist = A.take<int32_t>(P.nq);
        J.outFalseDyn = A.take<int32_t>(P.nq); J.outCount = A.take<int32_t>(1);
        if (A.failed) return api_fail(c, SDYN_ERR_NOMEM, "internal: arena sizing");
        MCU(c, up(qd, P.q_desc, (size_t)32 * P.nq, c->stream)); MCU(c, up(td, P.t_desc, (size_t)32 * P.nt, c->stream));
        MCU(c, up(qx, P.q_xy, (size_t)2 * P.nq, c->stream)); MCU(c, up(tx, P.t_xy, (size_t)2 * P.nt, c->stream));
    }
    MCU(c, up(dM, M, 9, c->stream)); MCU(c, up(dMi, Minv, 9, c->stream));
    MCU(c, up(dJobs, jobs.data(), npairs, c->stream));
    MCU(c, launch_box_pairs(dJobs, npairs, dM, dMi, mode, c->stream));
    c->launches += 1;
    std::vector<int32_t> counts(npairs);
    for (int p = 0; p < npairs; ++p) MCU(c, down(&counts[p], jobs[p].outCount, 1, c->stream));
    MCU(c, cudaStreamSynchronize(c->stream));
    for (int p = 0; p < npairs; ++p) {
        sdyn_box_pair& P = pairs[p];
        P.nmatches = counts[p];
        MCU(c, down(P.match_query, jobs[p].outQuery, counts[p], c->stream));
        MCU(c, down(P.match_train, jobs[p].outTrain, counts[p], c->stream));
        MCU(c, down(P.match_dist, jobs[p].outDist, counts[p], c->stream));
        MCU(c, down(P.false_dyn, jobs[p].outFalseDyn, counts[p], c->stream));
    }
    MCU(c, cudaStreamSynchronize(c->stream));
    return SDYN_OK;
}

}  // extern "C"
