/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * Flat C entry points so tests/ and bench.py's cpu_baseline leg can drive the oracle through ctypes. */
#include "orc_prims.h"
#include "orc_extractor.h"
#include "orc_matcher.h"
#include "orc_dynamic.h"
#include "orc_stereo.h"
#include "../include/sdyn.h"   /* POD layouts of the C ABI only (no product code is linked) */
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

void orc_undistort_points(const float* src, int n, float fx, float fy, float cx, float cy, const float* dist, int ndist, float* dst)
{ undistort_points(src, n, fx, fy, cx, cy, dist, ndist, dst); }

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

/* Frame::ComputeStereoMatches on the pyramids the two extractors hold after run() (src/Frame.cc:874-1048) */
int orc_stereo_matches(void* eL, void* eR, const KeyPoint* keysL, int N, const uint8_t* descL, const KeyPoint* keysR, int Nr,
                       const uint8_t* descR, float mb, float mbf, float* uRight, float* depth)
{
    Extractor* L = (Extractor*)eL; Extractor* R = (Extractor*)eR;
    if ((int)L->pyr.size() != L->nlevels || (int)R->pyr.size() != R->nlevels || L->nlevels != R->nlevels) return -2;
    std::vector<LevelRef> pl(L->nlevels), pr(R->nlevels);
    for (int l = 0; l < L->nlevels; ++l) {
        pl[l] = LevelRef{L->pyr[l].roi(), L->pyr[l].w, L->pyr[l].h, L->pyr[l].stride};
        pr[l] = LevelRef{R->pyr[l].roi(), R->pyr[l].w, R->pyr[l].h, R->pyr[l].stride};
    }
    return compute_stereo_matches(keysL, N, descL, keysR, Nr, descR, pl.data(), pr.data(), L->nlevels, L->scale.data(),
                                  L->invScale.data(), mb, mbf, uRight, depth);
}

float orc_ic_angle(const uint8_t* center, int stride)
{ static Extractor E(1000, 1.2f, 8, 20, 7); return ic_angle(center, stride, E.umax); }

void orc_orb_descriptor(float angle, const uint8_t* center, int stride, uint8_t* out) { orb_descriptor(angle, center, stride, out); }

/* ---- matchers ------------------------------------------------------------------------------------ */
static_assert(sizeof(MapPointQuery) == sizeof(sdyn_mappoint_query), "layout");
static_assert(sizeof(LastFramePoint) == sizeof(sdyn_last_point), "layout");
static_assert(sizeof(KeyPoint) == sizeof(sdyn_keypoint), "layout");

static FrameView view_of(const sdyn_frame_view* v)
{
    FrameView f;
    f.N = v->n; f.nlevels = v->nlevels;
    f.keys = reinterpret_cast<const KeyPoint*>(v->keys);
    f.keysUn = reinterpret_cast<const KeyPoint*>(v->keys_un);
    f.desc = v->desc; f.uRight = v->u_right; f.scaleFactors = v->scale_factors;
    f.minX = v->min_x; f.minY = v->min_y; f.maxX = v->max_x; f.maxY = v->max_y;
    f.fx = v->fx; f.fy = v->fy; f.cx = v->cx; f.cy = v->cy; f.bf = v->bf; f.b = v->b;
    for (int i = 0; i < 12; ++i) f.Tcw[i] = v->tcw[i];
    grid_bounds(f);
    return f;
}

int orc_hamming(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

/* Frame::GetFeaturesInArea on a freshly built grid (stage-level test of the grid semantics) */
int orc_features_in_area(const sdyn_frame_view* v, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    FrameView f = view_of(v); Grid g; assign_features_to_grid(f, g);
    std::vector<int> r_ = features_in_area(f, g, x, y, r, minLevel, maxLevel);
    for (int i = 0; i < (int)r_.size() && i < cap; ++i) out[i] = r_[i];
    return (int)r_.size();
}

int orc_match_projection_map(const sdyn_frame_view* v, const sdyn_mappoint_query* mps, int nmp, float th, float nnratio,
                             int32_t* assign, uint8_t* locked, int assignBase)
{
    FrameView f = view_of(v); Grid g; assign_features_to_grid(f, g);
    return search_by_projection_map(f, g, reinterpret_cast<const MapPointQuery*>(mps), nmp, th, nnratio, assign, locked, assignBase);
}

int orc_match_projection_frame(const sdyn_frame_view* cur, const sdyn_frame_view* last, const sdyn_last_point* lp, float th,
                               int mono, int checkOri, int32_t* assign, uint8_t* locked, float* pairs, int* npairs)
{
    FrameView c = view_of(cur), l = view_of(last); Grid g; assign_features_to_grid(c, g);
    std::vector<float> pr;
    int n = search_by_projection_frame(c, g, l, reinterpret_cast<const LastFramePoint*>(lp), th, mono != 0, checkOri != 0,
                                       assign, locked, pairs ? &pr : nullptr);
    if (pairs) { std::copy(pr.begin(), pr.end(), pairs); if (npairs) *npairs = (int)pr.size() / 4; }
    return n;
}

int orc_match_init(const sdyn_frame_view* f1, const sdyn_frame_view* f2, float* prev, int32_t* m12, int window, float nnratio,
                   int checkOri)
{
    FrameView a = view_of(f1), b = view_of(f2); Grid g; assign_features_to_grid(b, g);
    return search_for_initialization(a, b, g, prev, m12, window, nnratio, checkOri != 0);
}

int orc_match_bow(const sdyn_frame_view* kf, const uint8_t* kfValid, const sdyn_feature_vector* fa, const sdyn_frame_view* f,
                  const sdyn_feature_vector* fb, float nnratio, int checkOri, int32_t* assign)
{
    FrameView a = view_of(kf), b = view_of(f);
    FeatureVec va{fa->nnodes, fa->node_id, fa->offset, fa->index}, vb{fb->nnodes, fb->node_id, fb->offset, fb->index};
    return search_by_bow(a, kfValid, va, b, vb, nnratio, checkOri != 0, assign);
}

int orc_match_bow_kf(const sdyn_frame_view* kf1, const uint8_t* valid1, const sdyn_feature_vector* fa, const sdyn_frame_view* kf2,
                     const uint8_t* valid2, const sdyn_feature_vector* fb, float nnratio, int checkOri, int32_t* matches12)
{
    FrameView a = view_of(kf1), b = view_of(kf2);
    FeatureVec va{fa->nnodes, fa->node_id, fa->offset, fa->index}, vb{fb->nnodes, fb->node_id, fb->offset, fb->index};
    return search_by_bow_kf(a, valid1, va, b, valid2, vb, nnratio, checkOri != 0, matches12);
}

static_assert(sizeof(ProjPoint) == sizeof(sdyn_proj_point), "layout");
int orc_match_projection_pose(const sdyn_frame_view* target, const sdyn_proj_point* pts, int npts, const sdyn_proj_params* prm,
                              int32_t* assign)
{
    FrameView t = view_of(target); Grid g; assign_features_to_grid(t, g);
    const ProjPoint* pp = reinterpret_cast<const ProjPoint*>(pts);
    if (prm->variant == SDYN_PROJ_FRAME_KEYFRAME)
        return search_by_projection_reloc(t, g, pp, npts, prm->rcw, prm->tcw, prm->ow, prm->th, prm->max_descriptor_distance,
                                          prm->check_orientation != 0, prm->log_scale_factor, prm->nlevels, assign);
    return search_by_projection_sim3(t, g, pp, npts, prm->rcw, prm->tcw, prm->ow, (int)prm->th, prm->log_scale_factor,
                                     prm->nlevels, assign);
}

/* which: 0 = Fuse(pKF, vpMapPoints, th), 1 = Fuse(pKF, Scw, ...) — search parts */
void orc_fuse_search(int which, const sdyn_frame_view* kf, const float* invSigma2, const sdyn_proj_point* pts, int npts,
                     const float* Rcw, const float* tcw, const float* Ow, float th, float logSf, int nlevels, int32_t* bestIdx, int32_t* bestDist)
{
    FrameView t = view_of(kf); Grid g; assign_features_to_grid(t, g);
    const ProjPoint* pp = reinterpret_cast<const ProjPoint*>(pts);
    if (which == 0) fuse_search(t, g, invSigma2, pp, npts, Rcw, tcw, Ow, th, logSf, nlevels, bestIdx, bestDist);
    else fuse_sim3_search(t, g, pp, npts, Rcw, tcw, Ow, th, logSf, nlevels, bestIdx, bestDist);
}

int orc_search_by_sim3(const sdyn_frame_view* kf1, const sdyn_frame_view* kf2, const sdyn_proj_point* pts1, const sdyn_proj_point* pts2,
                       const float* T1w, const float* T2w, const float* S12, const float* S21, float th, float logSf, int nlevels,
                       int32_t* matches12)
{
    FrameView a = view_of(kf1), b = view_of(kf2); Grid g1, g2; assign_features_to_grid(a, g1); assign_features_to_grid(b, g2);
    return search_by_sim3(a, g1, b, g2, reinterpret_cast<const ProjPoint*>(pts1), reinterpret_cast<const ProjPoint*>(pts2), T1w, T2w,
                          S12, S21, th, logSf, nlevels, matches12);
}

int orc_match_triangulation(const sdyn_frame_view* kf1, const uint8_t* hasMp1, const sdyn_feature_vector* fa, const sdyn_frame_view* kf2,
                            const uint8_t* hasMp2, const sdyn_feature_vector* fb, const sdyn_tri_params* prm, int32_t* matches12)
{
    FrameView a = view_of(kf1), b = view_of(kf2);
    FeatureVec va{fa->nnodes, fa->node_id, fa->offset, fa->index}, vb{fb->nnodes, fb->node_id, fb->offset, fb->index};
    return search_for_triangulation(a, hasMp1, va, b, hasMp2, vb, prm->f12, prm->epipole_x, prm->epipole_y, prm->level_sigma2,
                                    prm->only_stereo != 0, prm->check_orientation != 0, matches12);
}

/* ---- dynamic ----------------------------------------------------------------------------------- */
void orc_box_mask(const sdyn_keypoint* keys, int n, const double* boxes, int nboxes, uint64_t* mask)
{
    for (int i = 0; i < n; ++i) {
        uint64_t m = 0;
        for (int b = 0; b < nboxes; ++b) {
            const double x = boxes[4 * b], y = boxes[4 * b + 1], w = boxes[4 * b + 2], h = boxes[4 * b + 3];
            const double px = keys[i].x, py = keys[i].y;
            if (x <= px && px < x + w && y <= py && py < y + h) m |= 1ull << b;
        }
        mask[i] = m;
    }
}

void orc_separate_pairs(sdyn_box_pair* pairs, int npairs, const float* M, int mode)
{
    for (int p = 0; p < npairs; ++p) {
        sdyn_box_pair& P = pairs[p];
        std::vector<Match> m = bf_match_crosscheck(P.q_desc, P.nq, P.t_desc, P.nt);
        std::vector<int> fd(m.size(), -1);
        if (mode == 1) classify_h(M, P.q_xy, P.t_xy, m, fd); else classify_f(M, P.q_xy, P.t_xy, m, fd);
        P.nmatches = (int)m.size();
        for (size_t i = 0; i < m.size(); ++i) {
            P.match_query[i] = m[i].queryIdx; P.match_train[i] = m[i].trainIdx; P.match_dist[i] = m[i].dist; P.false_dyn[i] = fd[i];
        }
    }
}

void orc_invert3x3(const float* m, float* out) { invert3x3(m, out); }

/* Frame::boxTrack on flat arrays.  boxes: in/out (capacity cap), returns the new box count. */
int orc_box_track(double* boxes, int nboxes, int cap, const double* lastObjects, const int* lastBoxIdx, const uint8_t* lastOmit,
                  const double* lastVel, int nlast, int imgW, int imgH, int* boxIdx, uint8_t* omit, double* vel)
{
    std::vector<Rect> b(nboxes);
    for (int i = 0; i < nboxes; ++i) b[i] = {boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]};
    BoxState last, cur;
    for (int i = 0; i < nlast; ++i) {
        last.objects.push_back({lastObjects[4 * i], lastObjects[4 * i + 1], lastObjects[4 * i + 2], lastObjects[4 * i + 3]});
        last.box_idx.push_back(lastBoxIdx[i]); last.omit.push_back(lastOmit[i]);
        last.vel.push_back(lastVel[2 * i]); last.vel.push_back(lastVel[2 * i + 1]);
    }
    box_track(b, last, imgW, imgH, cur);
    const int n = std::min((int)b.size(), cap);
    for (int i = 0; i < n; ++i) {
        boxes[4 * i] = b[i].x; boxes[4 * i + 1] = b[i].y; boxes[4 * i + 2] = b[i].w; boxes[4 * i + 3] = b[i].h;
        boxIdx[i] = cur.box_idx[i]; omit[i] = cur.omit[i]; vel[2 * i] = cur.vel[2 * i]; vel[2 * i + 1] = cur.vel[2 * i + 1];
    }
    return (int)b.size();
}

/* Frame::firstSeparate + tail split on flat arrays.
 *   boxes/boxIdx: in/out, compacted exactly as the reference erases them; returns the remaining box count.
 *   order[N]: new keypoint order; classId[N] (per ORIGINAL index); *nDyn = N_d;
 *   dynBox/dynKey: flattened (box, original keypoint index) pairs of the per-box arrays, *nDynPairs of them,
 *   in push order per box (box-major). */
int orc_first_separate(const sdyn_keypoint* keys, int N, double* boxes, int* boxIdx, int nboxes, int* order, int* classId,
                       int* nDyn, int* dynBox, int* dynKey, int cap, int* nDynPairs)
{
    std::vector<Rect> b(nboxes);
    BoxState cur;
    for (int i = 0; i < nboxes; ++i) {
        b[i] = {boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]};
        cur.box_idx.push_back(boxIdx[i]); cur.omit.push_back(0); cur.vel.push_back(0); cur.vel.push_back(0);
    }
    SplitResult R = first_separate(reinterpret_cast<const KeyPoint*>(keys), N, b, cur);
    for (int i = 0; i < N; ++i) { order[i] = R.order[i]; classId[i] = R.class_id[i]; }
    *nDyn = R.N_d;
    int k = 0;
    for (size_t bx = 0; bx < R.dynKeys.size(); ++bx)
        for (int id : R.dynKeys[bx]) { if (k < cap) { dynBox[k] = (int)bx; dynKey[k] = id; } ++k; }
    *nDynPairs = k;
    for (size_t i = 0; i < b.size(); ++i) {
        boxes[4 * i] = b[i].x; boxes[4 * i + 1] = b[i].y; boxes[4 * i + 2] = b[i].w; boxes[4 * i + 3] = b[i].h;
        boxIdx[i] = cur.box_idx[i];
    }
    return (int)b.size();
}

/* Tracking::classifyF (flag 2) / classifyH (flag 1) on explicit matches */
void orc_classify(int flag, const float* M, const float* curXY, const float* refXY, const int* query, const int* train, int nm, int* falseDyn)
{
    std::vector<Match> m(nm);
    for (int i = 0; i < nm; ++i) m[i] = {query[i], train[i], 0};
    std::vector<int> fd(nm, -1);
    if (flag == 1) classify_h(M, curXY, refXY, m, fd); else classify_f(M, curXY, refXY, m, fd);
    for (int i = 0; i < nm; ++i) falseDyn[i] = fd[i];
}

static void boxes_from_csr(int nb, const int* off, const float* xy, const uint8_t* desc, std::vector<BoxKeys>& out)
{
    out.resize(nb);
    for (int b = 0; b < nb; ++b) {
        out[b].xy.assign(xy + 2 * (size_t)off[b], xy + 2 * (size_t)off[b + 1]);
        out[b].desc.assign(desc + 32 * (size_t)off[b], desc + 32 * (size_t)off[b + 1]);
    }
}

/* Tracking::Separate (Tracking.cc:1093-1239) on flat arrays: per-box keypoints of the current / reference frame as CSR
 * (off[nboxes+1], xy = mvdynKeysUn, desc = mdynDescriptors).  dynStatus comes back as CSR; curBoxStatus is updated. */
int orc_separate(int ncur, const int* curOff, const float* curXY, const uint8_t* curDesc, const int* curBoxIdx, int* curBoxStatus,
                 int nref, const int* refOff, const float* refXY, const uint8_t* refDesc, const int* refBoxIdx,
                 int nlast, const int* lastBoxIdx, const int* lastBoxStatus, const float* HorF, int flag, int* dsOff, int* dsVal, int cap)
{
    std::vector<BoxKeys> cur, ref;
    boxes_from_csr(ncur, curOff, curXY, curDesc, cur);
    boxes_from_csr(nref, refOff, refXY, refDesc, ref);
    std::vector<int> cbi(curBoxIdx, curBoxIdx + ncur), cbs(curBoxStatus, curBoxStatus + ncur), rbi(refBoxIdx, refBoxIdx + nref);
    std::vector<int> lbi(lastBoxIdx, lastBoxIdx + nlast), lbs(lastBoxStatus, lastBoxStatus + nlast);
    std::vector<std::vector<int>> ds;
    std::vector<std::vector<Match>> matches;
    const int r = separate(cur, cbi, cbs, ref, rbi, lbi, lbs, HorF, flag, ds, matches);
    int n = 0;
    for (int b = 0; b < ncur; ++b) {
        dsOff[b] = n;
        for (int v : ds[b]) { if (n < cap) dsVal[n] = v; ++n; }
        curBoxStatus[b] = cbs[b];
    }
    dsOff[ncur] = n;
    return r;
}

/* Frame::UpdateFrame (Frame.cc:607-641): which (box, k) entries are appended, in push order */
int orc_update_frame(int nboxes, const int* dsOff, const int* dsVal, const int* cidOff, const int* cid, int* outBox, int* outK, int cap)
{
    std::vector<std::vector<int>> ds(nboxes), ci(nboxes);
    for (int b = 0; b < nboxes; ++b) { ds[b].assign(dsVal + dsOff[b], dsVal + dsOff[b + 1]); ci[b].assign(cid + cidOff[b], cid + cidOff[b + 1]); }
    const auto pushed = update_frame(ds, ci);
    for (size_t i = 0; i < pushed.size() && (int)i < cap; ++i) { outBox[i] = pushed[i].first; outK[i] = pushed[i].second; }
    return (int)pushed.size();
}

/* Frame::isInFrustum for n points: world / normal 3 floats each; outputs per point */
void orc_in_frustum(const sdyn_frame_view* v, float logSf, int n, const float* world, const float* normal, const float* minDist,
                    const float* maxDist, float cosLimit, uint8_t* inView, float* projX, float* projY, float* projXR, int* level, float* viewCos)
{
    FrameView f = view_of(v);
    for (int i = 0; i < n; ++i) {
        projX[i] = projY[i] = projXR[i] = viewCos[i] = 0.f; level[i] = 0;
        inView[i] = is_in_frustum(f, logSf, world + 3 * i, normal + 3 * i, minDist[i], maxDist[i], cosLimit, projX + i, projY + i,
                                  projXR + i, level + i, viewCos + i);
    }
}

}  // extern "C"
