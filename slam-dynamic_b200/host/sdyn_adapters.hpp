/* Glue between the reference's object graph (Frame, KeyFrame, MapPoint — include/Frame.h, KeyFrame.h,
 * MapPoint.h of the reference) and the index-based sdyn C ABI.  Header-only templates: they compile against
 * the reference's own classes when dropped into its tree (INTEGRATION.md shows the one-line bodies that
 * replace the loops in src/ORBmatcher.cc, src/Frame.cc and src/Tracking.cc) and against the small stand-ins of
 * tests/cpp/ref_stubs.h in this repository.
 *
 * Pointer <-> index mapping: Frame::mvpMapPoints[i] == NULL <-> assign[i] == -1; a pre-existing occupant is
 * passed as -2 with locked[i] = (Observations() > 0); a keypoint claimed by query q comes back as q.
 */
#ifndef SDYN_HOST_ADAPTERS_HPP
#define SDYN_HOST_ADAPTERS_HPP

#include "../../include/sdyn.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <type_traits>
#include <utility>
#include <vector>

namespace sdyn_host {

/* ---- error reporting of the adapters: a failed device call is not "no matches" --------------------------------------------
 * Every adapter returns what the reference function returns; a device failure is additionally reported on stderr with
 * sdyn_last_error (the reference has no error channel of its own: its functions return a match count). */
inline void report(sdyn_ctx* ctx, const char* where)
{
    std::fprintf(stderr, "sdyn: %s failed: %s\n", where, sdyn_last_error(ctx));
}


/* The Frame members the searches read, gathered into the ABI struct.  FrameT: the reference's Frame. */
template <class FrameT>
inline sdyn_frame_view frame_view(const FrameT& F)
{
    sdyn_frame_view v;
    std::memset(&v, 0, sizeof(v));
    v.n = F.N;
    v.nlevels = F.mnScaleLevels;
    v.keys = reinterpret_cast<const sdyn_keypoint*>(F.mvKeys.data());
    v.keys_un = reinterpret_cast<const sdyn_keypoint*>(F.mvKeysUn.data());
    v.desc = F.mDescriptors.data;                       /* N x 32 CV_8U, continuous */
    v.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
    v.scale_factors = F.mvScaleFactors.data();
    v.min_x = FrameT::mnMinX; v.min_y = FrameT::mnMinY; v.max_x = FrameT::mnMaxX; v.max_y = FrameT::mnMaxY;
    v.fx = FrameT::fx; v.fy = FrameT::fy; v.cx = FrameT::cx; v.cy = FrameT::cy; v.bf = F.mbf; v.b = F.mb;
    if (!F.mTcw.empty())
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) v.tcw[4 * r + c] = F.mTcw.template at<float>(r, c);
    return v;
}

template <class FrameT>
inline void occupancy(const FrameT& F, std::vector<int32_t>& assign, std::vector<uint8_t>& locked)
{
    assign.assign(F.N, -1);
    locked.assign(F.N, 0);
    for (int i = 0; i < F.N; ++i)
        if (F.mvpMapPoints[i]) { assign[i] = -2; locked[i] = F.mvpMapPoints[i]->Observations() > 0; }
}

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)
 * reference: src/ORBmatcher.cc:45-129 */
template <class FrameT, class MapPointT>
inline int SearchByProjection(sdyn_ctx* ctx, FrameT& F, const std::vector<MapPointT*>& vpMapPoints, float th, float nnratio)
{
    std::vector<sdyn_mappoint_query> q(vpMapPoints.size());
    for (size_t i = 0; i < vpMapPoints.size(); ++i) {
        MapPointT* p = vpMapPoints[i];
        sdyn_mappoint_query& m = q[i];
        std::memset(&m, 0, sizeof(m));
        m.track_in_view = p->mbTrackInView;
        m.bad = p->isBad();
        if (!m.track_in_view || m.bad) continue;
        m.proj_x = p->mTrackProjX; m.proj_y = p->mTrackProjY; m.proj_xr = p->mTrackProjXR;
        m.view_cos = p->mTrackViewCos; m.level = p->mnTrackScaleLevel;
        m.obs_positive = p->Observations() > 0;
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(m.desc, d.data, 32);
    }
    std::vector<int32_t> assign; std::vector<uint8_t> locked;
    occupancy(F, assign, locked);
    sdyn_frame_view v = frame_view(F);
    int n = 0;
    if (sdyn_match_projection_map(ctx, &v, q.data(), (int)q.size(), th, nnratio, assign.data(), locked.data(), &n) != SDYN_OK) {
        report(ctx, __func__);
        return 0;
    }
    for (int i = 0; i < F.N; ++i)
        if (assign[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[assign[i]];
    return n;
}

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
 * and the fork's overload with point-pair outputs — reference: src/ORBmatcher.cc:1485-1627, 407-559 */
template <class FrameT, class PointT>
inline int SearchByProjection(sdyn_ctx* ctx, FrameT& Cur, const FrameT& Last, float th, bool bMono, bool checkOrientation,
                              std::vector<PointT>* pointsLast, std::vector<PointT>* pointsCurrent)
{
    std::vector<sdyn_last_point> lp(Last.N);
    for (int i = 0; i < Last.N; ++i) {
        sdyn_last_point& p = lp[i];
        std::memset(&p, 0, sizeof(p));
        auto* mp = Last.mvpMapPoints[i];
        if (!mp) continue;
        p.has_mp = 1; p.outlier = Last.mvbOutlier[i]; p.obs_positive = mp->Observations() > 0;
        const cv::Mat w = mp->GetWorldPos();
        for (int k = 0; k < 3; ++k) p.world[k] = w.template at<float>(k, 0);
        const cv::Mat d = mp->GetDescriptor();
        std::memcpy(p.desc, d.data, 32);
    }
    std::vector<int32_t> assign; std::vector<uint8_t> locked;
    occupancy(Cur, assign, locked);
    sdyn_frame_view c = frame_view(Cur), l = frame_view(Last);
    std::vector<float> pairs(pointsLast ? (size_t)4 * std::max(Last.N, 1) : 0);
    int n = 0, np = 0;
    if (sdyn_match_projection_frame(ctx, &c, &l, lp.data(), th, bMono, checkOrientation, assign.data(), locked.data(), &n,
                                    pointsLast ? pairs.data() : nullptr, pointsLast ? &np : nullptr) != SDYN_OK) {
        report(ctx, __func__);
        return 0;
    }
    for (int i = 0; i < Cur.N; ++i) {
        if (assign[i] >= 0) Cur.mvpMapPoints[i] = Last.mvpMapPoints[assign[i]];
        else if (assign[i] == -1) Cur.mvpMapPoints[i] = nullptr;      /* nulled by the rotation cull */
    }
    if (pointsLast && pointsCurrent)
        for (int k = 0; k < np; ++k) {
            pointsLast->push_back(PointT(pairs[4 * k], pairs[4 * k + 1]));
            pointsCurrent->push_back(PointT(pairs[4 * k + 2], pairs[4 * k + 3]));
        }
    return n;
}

/* ORBmatcher::SearchForInitialization — reference: src/ORBmatcher.cc:562-677 */
template <class FrameT, class PointT>
inline int SearchForInitialization(sdyn_ctx* ctx, FrameT& F1, FrameT& F2, std::vector<PointT>& vbPrevMatched,
                                   std::vector<int>& vnMatches12, int windowSize, float nnratio, bool checkOrientation)
{
    const int n1 = (int)F1.mvKeysUn.size();
    vnMatches12.assign(n1, -1);
    std::vector<float> prev((size_t)2 * n1);
    for (int i = 0; i < n1; ++i) { prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    sdyn_frame_view a = frame_view(F1), b = frame_view(F2);
    int n = 0;
    static_assert(sizeof(int) == sizeof(int32_t), "int32 matches");
    if (sdyn_match_init(ctx, &a, &b, prev.data(), vnMatches12.data(), windowSize, nnratio, checkOrientation, &n) != SDYN_OK) {
        report(ctx, __func__);
        return 0;
    }
    for (int i = 0; i < n1; ++i) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
    return n;
}

/* DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) -> CSR */
struct FeatureVectorCSR {
    std::vector<uint32_t> node, index; std::vector<int32_t> offset;
    sdyn_feature_vector view() const { return {(int32_t)node.size(), node.data(), offset.data(), index.data()}; }
    template <class MapT> explicit FeatureVectorCSR(const MapT& fv)
    {
        offset.push_back(0);
        for (auto it = fv.begin(); it != fv.end(); ++it) {
            node.push_back((uint32_t)it->first);
            for (unsigned v : it->second) index.push_back(v);
            offset.push_back((int32_t)index.size());
        }
    }
};

/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches)
 * reference: src/ORBmatcher.cc:159-288 */
template <class KeyFrameT, class FrameT, class MapPointT>
inline int SearchByBoW(sdyn_ctx* ctx, KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches, float nnratio,
                       bool checkOrientation)
{
    const std::vector<MapPointT*> kfPoints = pKF->GetMapPointMatches();
    vpMapPointMatches.assign(F.N, static_cast<MapPointT*>(nullptr));
    std::vector<uint8_t> valid(kfPoints.size(), 0);
    for (size_t i = 0; i < kfPoints.size(); ++i) valid[i] = kfPoints[i] && !kfPoints[i]->isBad();
    sdyn_frame_view kv;
    std::memset(&kv, 0, sizeof(kv));
    kv.n = (int)kfPoints.size(); kv.nlevels = F.mnScaleLevels;
    kv.keys = kv.keys_un = reinterpret_cast<const sdyn_keypoint*>(pKF->mvKeysUn.data());
    kv.desc = pKF->mDescriptors.data;
    kv.max_x = kv.max_y = 1.f;
    sdyn_frame_view fv = frame_view(F);
    FeatureVectorCSR a(pKF->mFeatVec), b(F.mFeatVec);
    sdyn_feature_vector av = a.view(), bv = b.view();
    std::vector<int32_t> assign(F.N, -1);
    int n = 0;
    if (sdyn_match_bow(ctx, &kv, valid.data(), &av, &fv, &bv, nnratio, checkOrientation, assign.data(), &n) != SDYN_OK) {
        report(ctx, __func__);
        return 0;
    }
    for (int i = 0; i < F.N; ++i)
        if (assign[i] >= 0) vpMapPointMatches[i] = kfPoints[assign[i]];
    return n;
}

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12)
 * reference: src/ORBmatcher.cc:679-812 (LoopClosing.cc:266) */
template <class KeyFrameT, class MapPointT>
inline int SearchByBoW(sdyn_ctx* ctx, KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, float nnratio,
                       bool checkOrientation)
{
    const std::vector<MapPointT*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
    vpMatches12.assign(mp1.size(), static_cast<MapPointT*>(nullptr));
    std::vector<uint8_t> v1(mp1.size()), v2(mp2.size());
    for (size_t i = 0; i < mp1.size(); ++i) v1[i] = mp1[i] && !mp1[i]->isBad();
    for (size_t i = 0; i < mp2.size(); ++i) v2[i] = mp2[i] && !mp2[i]->isBad();
    sdyn_frame_view a, b;
    std::memset(&a, 0, sizeof(a)); std::memset(&b, 0, sizeof(b));
    a.n = (int)mp1.size(); b.n = (int)mp2.size(); a.nlevels = b.nlevels = pKF1->mnScaleLevels;
    a.keys = a.keys_un = reinterpret_cast<const sdyn_keypoint*>(pKF1->mvKeysUn.data()); a.desc = pKF1->mDescriptors.data;
    b.keys = b.keys_un = reinterpret_cast<const sdyn_keypoint*>(pKF2->mvKeysUn.data()); b.desc = pKF2->mDescriptors.data;
    a.max_x = a.max_y = b.max_x = b.max_y = 1.f;
    FeatureVectorCSR fa(pKF1->mFeatVec), fb(pKF2->mFeatVec);
    sdyn_feature_vector av = fa.view(), bv = fb.view();
    std::vector<int32_t> m12(mp1.size(), -1);
    int n = 0;
    if (sdyn_match_bow_kf(ctx, &a, v1.data(), &av, &b, v2.data(), &bv, nnratio, checkOrientation, m12.data(), &n) != SDYN_OK) {
        report(ctx, __func__);
        return 0;
    }
    for (size_t i = 0; i < mp1.size(); ++i)
        if (m12[i] >= 0) vpMatches12[i] = mp2[m12[i]];
    return n;
}

#ifdef SDYN_HAVE_OPENCV
/* MapPoint::mfMaxDistance (PredictScale divides it by the current distance, src/MapPoint.cc:385-418) is a protected member
 * of the reference's MapPoint and only 1.2f * mfMaxDistance is published.  No header of the reference is edited for it: a
 * pointer to the member, named through a derived class, is legal C++ and reads the field of any MapPoint (the value is
 * written once at creation / UpdateNormalAndDepth; the float read is as atomic as the reference's own locked read). */
template <class MapPointT>
struct MaxDistancePeek : MapPointT {
    static float get(MapPointT* p) { return p->*(&MaxDistancePeek::mfMaxDistance); }
};

/* One candidate MapPoint of the pose-projection searches. */
template <class MapPointT>
inline void proj_point(MapPointT* pMP, bool valid, float angle, sdyn_proj_point& q)
{
    std::memset(&q, 0, sizeof(q));
    q.valid = valid;
    if (!valid) return;
    const cv::Mat w = pMP->GetWorldPos(), nrm = pMP->GetNormal(), d = pMP->GetDescriptor();
    for (int k = 0; k < 3; ++k) { q.world[k] = w.at<float>(k); q.normal[k] = nrm.at<float>(k); }
    q.min_distance = pMP->GetMinDistanceInvariance(); q.max_distance = pMP->GetMaxDistanceInvariance();
    q.max_distance_raw = MaxDistancePeek<MapPointT>::get(pMP);
    q.angle = angle;
    std::memcpy(q.desc, d.data, 32);
}

inline void pose_params(const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, sdyn_proj_params& p)
{
    for (int r = 0; r < 3; ++r) {
        for (int k = 0; k < 3; ++k) p.rcw[3 * r + k] = Rcw.at<float>(r, k);
        p.tcw[r] = tcw.at<float>(r); p.ow[r] = Ow.at<float>(r);
    }
}

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist)
 * reference: src/ORBmatcher.cc:1629-1756 (Tracking::Relocalization, Tracking.cc:2323,2337) */
template <class FrameT, class KeyFrameT, class SetT>
inline int SearchByProjection(sdyn_ctx* ctx, FrameT& CurrentFrame, KeyFrameT* pKF, const SetT& sAlreadyFound, float th,
                              int ORBdist, bool checkOrientation)
{
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);      /* :1633-1635, evaluated by OpenCV itself */
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    const auto vpMPs = pKF->GetMapPointMatches();
    std::vector<sdyn_proj_point> q(vpMPs.size());
    for (size_t i = 0; i < vpMPs.size(); ++i)
        proj_point(vpMPs[i], vpMPs[i] && !vpMPs[i]->isBad() && !sAlreadyFound.count(vpMPs[i]), pKF->mvKeysUn[i].angle, q[i]);
    sdyn_proj_params p;
    std::memset(&p, 0, sizeof(p));
    pose_params(Rcw, tcw, Ow, p);
    p.th = th; p.max_descriptor_distance = ORBdist; p.variant = SDYN_PROJ_FRAME_KEYFRAME; p.check_orientation = checkOrientation;
    p.log_scale_factor = CurrentFrame.mfLogScaleFactor; p.nlevels = CurrentFrame.mnScaleLevels;
    std::vector<int32_t> assign(CurrentFrame.N);
    for (int i = 0; i < CurrentFrame.N; ++i) assign[i] = CurrentFrame.mvpMapPoints[i] ? -2 : -1;
    sdyn_frame_view v = frame_view(CurrentFrame);
    int n = 0;
    if (sdyn_match_projection_pose(ctx, &v, q.data(), (int)q.size(), &p, assign.data(), &n) != SDYN_OK) { report(ctx, __func__); return 0; }
    for (int i = 0; i < CurrentFrame.N; ++i)
        if (assign[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[assign[i]];
    return n;
}

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
 * reference: src/ORBmatcher.cc:290-403 (LoopClosing::ComputeSim3, LoopClosing.cc:376) */
template <class KeyFrameT, class MapPointT>
inline int SearchByProjection(sdyn_ctx* ctx, KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints,
                              std::vector<MapPointT*>& vpMatched, int th)
{
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);                          /* :300-304 */
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    std::set<MapPointT*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPointT*>(nullptr));
    std::vector<sdyn_proj_point> q(vpPoints.size());
    for (size_t i = 0; i < vpPoints.size(); ++i)
        proj_point(vpPoints[i], !vpPoints[i]->isBad() && !spAlreadyFound.count(vpPoints[i]), 0.f, q[i]);
    sdyn_proj_params p;
    std::memset(&p, 0, sizeof(p));
    pose_params(Rcw, tcw, Ow, p);
    p.th = (float)th; p.max_descriptor_distance = SDYN_TH_LOW; p.variant = SDYN_PROJ_KEYFRAME_SIM3;
    p.log_scale_factor = pKF->mfLogScaleFactor; p.nlevels = pKF->mnScaleLevels;
    sdyn_frame_view v;
    std::memset(&v, 0, sizeof(v));
    v.n = pKF->N; v.nlevels = pKF->mnScaleLevels;
    v.keys = v.keys_un = reinterpret_cast<const sdyn_keypoint*>(pKF->mvKeysUn.data());
    v.desc = pKF->mDescriptors.data; v.scale_factors = pKF->mvScaleFactors.data();
    v.min_x = pKF->mnMinX; v.min_y = pKF->mnMinY; v.max_x = pKF->mnMaxX; v.max_y = pKF->mnMaxY;
    v.fx = pKF->fx; v.fy = pKF->fy; v.cx = pKF->cx; v.cy = pKF->cy;
    std::vector<int32_t> assign(pKF->N);
    for (int i = 0; i < pKF->N; ++i) assign[i] = vpMatched[i] ? -2 : -1;
    int n = 0;
    if (sdyn_match_projection_pose(ctx, &v, q.data(), (int)q.size(), &p, assign.data(), &n) != SDYN_OK) { report(ctx, __func__); return 0; }
    for (int i = 0; i < pKF->N; ++i)
        if (assign[i] >= 0) vpMatched[i] = vpPoints[assign[i]];
    return n;
}
template <class KeyFrameT>
inline sdyn_frame_view keyframe_view(KeyFrameT* pKF)
{
    sdyn_frame_view v;
    std::memset(&v, 0, sizeof(v));
    v.n = pKF->N; v.nlevels = pKF->mnScaleLevels;
    v.keys = v.keys_un = reinterpret_cast<const sdyn_keypoint*>(pKF->mvKeysUn.data());
    v.desc = pKF->mDescriptors.data; v.scale_factors = pKF->mvScaleFactors.data();
    v.u_right = pKF->mvuRight.empty() ? nullptr : pKF->mvuRight.data();
    v.min_x = pKF->mnMinX; v.min_y = pKF->mnMinY; v.max_x = pKF->mnMaxX; v.max_y = pKF->mnMaxY;
    v.fx = pKF->fx; v.fy = pKF->fy; v.cx = pKF->cx; v.cy = pKF->cy; v.bf = pKF->mbf;
    return v;
}

inline void rigid_rows(const cv::Mat& R, const cv::Mat& t, float out[12])
{
    for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) out[4 * r + k] = R.at<float>(r, k); out[4 * r + 3] = t.at<float>(r); }
}

/* ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th)
 * reference: src/ORBmatcher.cc:982-1130 (LocalMapping::SearchInNeighbors).  The keypoint choice of every MapPoint
 * (:1000-1100) runs on the device; the Replace / AddObservation bookkeeping (:1103-1126) runs here in the
 * reference's order, re-testing isBad() / IsInKeyFrame() at its turn because earlier iterations change them. */
template <class KeyFrameT, class MapPointT>
inline int Fuse(sdyn_ctx* ctx, KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, float th)
{
    const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
    std::vector<sdyn_proj_point> q(vpMapPoints.size());
    for (size_t i = 0; i < vpMapPoints.size(); ++i)
        proj_point(vpMapPoints[i], vpMapPoints[i] && !vpMapPoints[i]->isBad() && !vpMapPoints[i]->IsInKeyFrame(pKF), 0.f, q[i]);
    sdyn_best_params p;
    std::memset(&p, 0, sizeof(p));
    rigid_rows(Rcw, tcw, p.t1);
    for (int k = 0; k < 3; ++k) p.ow[k] = Ow.at<float>(k);
    p.check_normal = 1; p.chi2_gate = 1; p.bf = pKF->mbf; p.th = th;
    for (int l = 0; l < pKF->mnScaleLevels && l < SDYN_MAX_LEVELS; ++l) p.inv_level_sigma2[l] = pKF->mvInvLevelSigma2[l];
    p.log_scale_factor = pKF->mfLogScaleFactor; p.nlevels = pKF->mnScaleLevels;
    sdyn_frame_view v = keyframe_view(pKF);
    std::vector<int32_t> bestIdx(q.size()), bestDist(q.size());
    if (sdyn_match_projection_best(ctx, &v, q.data(), (int)q.size(), &p, bestIdx.data(), bestDist.data()) != SDYN_OK) { report(ctx, __func__); return 0; }
    int nFused = 0;
    for (size_t i = 0; i < vpMapPoints.size(); ++i) {
        MapPointT* pMP = vpMapPoints[i];
        if (!pMP || !q[i].valid) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;           /* state may have changed since the call started */
        if (bestIdx[i] < 0 || bestDist[i] > SDYN_TH_LOW) continue;
        MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, bestIdx[i]);
            pKF->AddMapPoint(pMP, bestIdx[i]);
        }
        ++nFused;
    }
    return nFused;
}

/* ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint)
 * reference: src/ORBmatcher.cc:1132-1257 (LoopClosing::SearchAndFuse) */
template <class KeyFrameT, class MapPointT>
inline int Fuse(sdyn_ctx* ctx, KeyFrameT* pKF, cv::Mat Scw, const std::vector<MapPointT*>& vpPoints, float th,
                std::vector<MapPointT*>& vpReplacePoint)
{
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    const std::set<MapPointT*> spAlreadyFound = pKF->GetMapPoints();
    std::vector<sdyn_proj_point> q(vpPoints.size());
    for (size_t i = 0; i < vpPoints.size(); ++i)
        proj_point(vpPoints[i], !vpPoints[i]->isBad() && !spAlreadyFound.count(vpPoints[i]), 0.f, q[i]);
    sdyn_best_params p;
    std::memset(&p, 0, sizeof(p));
    rigid_rows(Rcw, tcw, p.t1);
    for (int k = 0; k < 3; ++k) p.ow[k] = Ow.at<float>(k);
    p.invz_double = 1; p.check_normal = 1; p.th = th;
    p.log_scale_factor = pKF->mfLogScaleFactor; p.nlevels = pKF->mnScaleLevels;
    sdyn_frame_view v = keyframe_view(pKF);
    v.u_right = nullptr;
    std::vector<int32_t> bestIdx(q.size()), bestDist(q.size());
    if (sdyn_match_projection_best(ctx, &v, q.data(), (int)q.size(), &p, bestIdx.data(), bestDist.data()) != SDYN_OK) { report(ctx, __func__); return 0; }
    int nFused = 0;
    for (size_t i = 0; i < vpPoints.size(); ++i) {
        if (!q[i].valid || bestIdx[i] < 0 || bestDist[i] > SDYN_TH_LOW) continue;
        MapPointT* pMPinKF = pKF->GetMapPoint(bestIdx[i]);
        if (pMPinKF) { if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF; }
        else { vpPoints[i]->AddObservation(pKF, bestIdx[i]); pKF->AddMapPoint(vpPoints[i], bestIdx[i]); }
        ++nFused;
    }
    return nFused;
}

/* ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12, s12, R12, t12, th)
 * reference: src/ORBmatcher.cc:1259-1483 (LoopClosing::ComputeSim3) */
template <class KeyFrameT, class MapPointT>
inline int SearchBySim3(sdyn_ctx* ctx, KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const float& s12,
                        const cv::Mat& R12, const cv::Mat& t12, float th)
{
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;                                           /* :1276-1278 */
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    const std::vector<MapPointT*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
    const int N1 = (int)mp1.size(), N2 = (int)mp2.size();
    std::vector<bool> done1(N1, false), done2(N2, false);
    for (int i = 0; i < N1; ++i)
        if (vpMatches12[i]) {
            done1[i] = true;
            const int idx2 = vpMatches12[i]->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2) done2[idx2] = true;
        }
    std::vector<sdyn_proj_point> q1(N1), q2(N2);
    for (int i = 0; i < N1; ++i) proj_point(mp1[i], mp1[i] && !done1[i] && !mp1[i]->isBad(), 0.f, q1[i]);
    for (int i = 0; i < N2; ++i) proj_point(mp2[i], mp2[i] && !done2[i] && !mp2[i]->isBad(), 0.f, q2[i]);
    sdyn_best_params a, b;
    std::memset(&a, 0, sizeof(a)); std::memset(&b, 0, sizeof(b));
    rigid_rows(R1w, t1w, a.t1); rigid_rows(sR21, t21, a.t2);
    rigid_rows(R2w, t2w, b.t1); rigid_rows(sR12, t12, b.t2);
    a.use_t2 = b.use_t2 = 1; a.invz_double = b.invz_double = 1; a.dist_from_camera = b.dist_from_camera = 1; a.th = b.th = th;
    a.log_scale_factor = pKF2->mfLogScaleFactor; a.nlevels = pKF2->mnScaleLevels;
    b.log_scale_factor = pKF1->mfLogScaleFactor; b.nlevels = pKF1->mnScaleLevels;
    sdyn_frame_view v1 = keyframe_view(pKF1), v2 = keyframe_view(pKF2);
    v1.u_right = v2.u_right = nullptr;
    std::vector<int32_t> i12(N1), d12(N1), i21(N2), d21(N2);
    if (sdyn_match_projection_best(ctx, &v2, q1.data(), N1, &a, i12.data(), d12.data()) != SDYN_OK) { report(ctx, __func__); return 0; }
    if (sdyn_match_projection_best(ctx, &v1, q2.data(), N2, &b, i21.data(), d21.data()) != SDYN_OK) { report(ctx, __func__); return 0; }
    int nFound = 0;
    for (int i1 = 0; i1 < N1; ++i1) {                                   /* :1462-1478 */
        const int idx2 = (i12[i1] >= 0 && d12[i1] <= SDYN_TH_HIGH) ? i12[i1] : -1;
        if (idx2 < 0) continue;
        const int idx1 = (i21[idx2] >= 0 && d21[idx2] <= SDYN_TH_HIGH) ? i21[idx2] : -1;
        if (idx1 == i1) { vpMatches12[i1] = mp2[idx2]; ++nFound; }
    }
    return nFound;
}

/* ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vector<pair<size_t,size_t>> &vMatchedPairs,
 * const bool bOnlyStereo) — reference: src/ORBmatcher.cc:814-980 (LocalMapping::CreateNewMapPoints, LocalMapping.cc:269) */
template <class KeyFrameT>
inline int SearchForTriangulation(sdyn_ctx* ctx, KeyFrameT* pKF1, KeyFrameT* pKF2, cv::Mat F12,
                                  std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo, bool checkOrientation)
{
    cv::Mat Cw = pKF1->GetCameraCenter();                               /* epipole in the second image, :823-829 */
    cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    sdyn_tri_params p;
    std::memset(&p, 0, sizeof(p));
    p.epipole_x = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
    p.epipole_y = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) p.f12[3 * r + k] = F12.at<float>(r, k);
    p.only_stereo = bOnlyStereo; p.check_orientation = checkOrientation;
    for (int l = 0; l < pKF2->mnScaleLevels && l < SDYN_MAX_LEVELS; ++l) p.level_sigma2[l] = pKF2->mvLevelSigma2[l];
    std::vector<uint8_t> h1(pKF1->N), h2(pKF2->N);
    for (int i = 0; i < pKF1->N; ++i) h1[i] = pKF1->GetMapPoint(i) != nullptr;
    for (int i = 0; i < pKF2->N; ++i) h2[i] = pKF2->GetMapPoint(i) != nullptr;
    sdyn_frame_view a = keyframe_view(pKF1), b = keyframe_view(pKF2);
    FeatureVectorCSR fa(pKF1->mFeatVec), fb(pKF2->mFeatVec);
    sdyn_feature_vector av = fa.view(), bv = fb.view();
    std::vector<int32_t> m12(pKF1->N, -1);
    int n = 0;
    vMatchedPairs.clear();
    if (sdyn_match_triangulation(ctx, &a, h1.data(), &av, &b, h2.data(), &bv, &p, m12.data(), &n) != SDYN_OK) { report(ctx, __func__); return 0; }
    vMatchedPairs.reserve(n);
    for (size_t i = 0; i < m12.size(); ++i)
        if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
    return n;
}
#endif  /* SDYN_HAVE_OPENCV */

/* Frame::ComputeBoW() / KeyFrame::ComputeBoW() — reference: src/Frame.cc:803-810, src/KeyFrame.cc:76-86.  The tree
 * descent of all descriptors runs on the device; the two DBoW2 containers are then filled through their own
 * methods in the insertion order of TemplatedVocabulary::transform (TemplatedVocabulary.h:1147-1164, 1194), so
 * mBowVec / mFeatVec are the reference's objects with the reference's values. */
template <class FrameT, class LNormT>
inline bool ComputeBoW(sdyn_ctx* ctx, const sdyn_vocab* voc, FrameT& F, LNormT l1Norm /* DBoW2::L1 */)
{
    if (!F.mBowVec.empty()) return true;
    const int n = F.mDescriptors.rows;
    std::vector<uint32_t> word((size_t)std::max(n, 1)), node((size_t)std::max(n, 1));
    std::vector<double> weight((size_t)std::max(n, 1));
    if (sdyn_bow_transform(ctx, voc, F.mDescriptors.data, n, 4, word.data(), weight.data(), node.data()) != SDYN_OK) return false;
    F.mBowVec.clear(); F.mFeatVec.clear();
    for (int i = 0; i < n; ++i)
        if (weight[i] > 0) { F.mBowVec.addWeight(word[i], weight[i]); F.mFeatVec.addFeature(node[i], (unsigned)i); }
    F.mBowVec.normalize(l1Norm);
    return true;
}

/* Frame::ComputeStereoMatches() — reference: src/Frame.cc:874-1048.  Called where the reference calls it, right
 * after the two ExtractORB threads joined (src/Frame.cc:151-160): it works on the keypoints, descriptors and
 * pyramids the left / right ORBextractor contexts still hold on the device, so mvImagePyramid is not read on the host. */
template <class FrameT>
inline bool ComputeStereoMatches(FrameT& F)
{
    F.mvuRight.assign(F.N, -1.0f);
    F.mvDepth.assign(F.N, -1.0f);
    if (F.N == 0) return true;
    return sdyn_stereo_match(F.mpORBextractorLeft->Context(), F.mpORBextractorRight->Context(), 1, F.mb, F.mbf,
                             F.mvuRight.data(), F.mvDepth.data(), F.N, nullptr) == SDYN_OK;
}

/* Frame::firstSeparate (src/Frame.cc:555-604): the keypoint-in-box test runs on the device; the reorder and
 * the reference's box bookkeeping (incl. its erase-while-iterating behaviour) stay host code in Frame. */
template <class KeyPointT, class RectT>
inline bool BoxMask(sdyn_ctx* ctx, const std::vector<KeyPointT>& keys, const std::vector<RectT>& boxes, std::vector<uint64_t>& mask)
{
    static_assert(sizeof(RectT) == 4 * sizeof(double), "cv::Rect2d layout");
    mask.assign(keys.size(), 0);
    return sdyn_dyn_box_mask(ctx, reinterpret_cast<const sdyn_keypoint*>(keys.data()), (int)keys.size(),
                             reinterpret_cast<const double*>(boxes.data()), (int)boxes.size(), mask.data()) == SDYN_OK;
}

/* Frame::firstSeparate (src/Frame.cc:555-604) as a whole: box test on the device (one 64-bit word per keypoint), then the
 * reference's own bookkeeping on the host — class_id stamping, the static-first reorder of keys and descriptors, the erase of
 * boxes without keypoints with its erase-while-iterating behaviour (the loop index advances after an erase and hasKpts is not
 * erased in step) and N_d.  Members touched are exactly those the reference function touches.  <= 64 boxes per frame. */
template <class FrameT, class RectT>
inline bool FirstSeparate(sdyn_ctx* ctx, FrameT& F, std::vector<RectT>& boxes, std::vector<std::vector<int>>& index,
                          std::vector<bool>& hasKpts)
{
    std::vector<uint64_t> mask;
    if (boxes.size() > 64 || !BoxMask(ctx, F.mvKeys, boxes, mask)) { report(ctx, "Frame::firstSeparate (box mask)"); mask.assign(F.mvKeys.size(), 0); }
    typedef typename std::remove_reference<decltype(F.mvKeys)>::type KeyVec;
    KeyVec statKeys, dynKeys;
    std::vector<int> statRows, dynRows;
    for (size_t i = 0; i < F.mvKeys.size(); ++i) {
        std::vector<int> idx;
        for (size_t j = 0; j < boxes.size(); ++j)
            if ((mask[i] >> j) & 1ull) {
                hasKpts[j] = true;
                idx.push_back((int)j);
                if (F.mvKeys[i].class_id == -1) F.mvKeys[i].class_id = (int)i;
            }
        if (!idx.empty()) { index.push_back(idx); dynKeys.push_back(F.mvKeys[i]); dynRows.push_back((int)i); }
        else { statKeys.push_back(F.mvKeys[i]); statRows.push_back((int)i); }
    }
    bool empty_box = false;
    for (size_t i = 0; i < boxes.size(); i++) {
        if (hasKpts[i] == true) continue;
        boxes.erase(boxes.begin() + i);
        F.box_idx.erase(F.box_idx.begin() + i);
        F.omit.erase(F.omit.begin() + i);
        F.box_velocity.erase(F.box_velocity.begin() + i);
        empty_box = true;
    }
    F.N_d = (int)index.size();
    if (F.N_d > 0) {
        F.mvKeys = statKeys;
        F.mvKeys.insert(F.mvKeys.end(), dynKeys.begin(), dynKeys.end());
        decltype(F.mDescriptors) D(F.mDescriptors.rows, 32, F.mDescriptors.type());
        int r = 0;
        for (int src : statRows) std::memcpy(D.ptr(r++), F.mDescriptors.ptr(src), 32);
        for (int src : dynRows) std::memcpy(D.ptr(r++), F.mDescriptors.ptr(src), 32);
        F.mDescriptors = D;
    }
    return empty_box;
}

/* Tracking::Separate (src/Tracking.cc:1093-1239): per current box joined to a box of the reference frame, cv::BFMatcher
 * cross-check + classifyF / classifyH run on the device for ALL boxes in one call (sdyn_dyn_separate); the >= 3 / 20 %
 * gates, the "static if more than max(1, 20 %) survive" rule and the box_status bookkeeping are the reference's, on the host.
 * Debug output of the reference (drawMatches, imwrite, putText) is not reproduced.  Returns what Separate returns. */
template <class FrameT, class MatT>
inline int Separate(sdyn_ctx* ctx, FrameT& Cur, FrameT& Ref, FrameT& Last, const MatT& HorF, int flag, std::vector<std::vector<int>>& dynStatus)
{
    const size_t nb = Cur.objects.size();
    dynStatus.resize(nb);
    struct Job { size_t box, ref; std::vector<float> qxy, txy; std::vector<int32_t> mq, mt, md, fd; };
    std::vector<Job> jobs;
    std::vector<sdyn_box_pair> pairs;
    for (size_t b = 0; b < nb; ++b) {
        auto it = std::find(Ref.box_idx.begin(), Ref.box_idx.end(), Cur.box_idx[b]);
        if (it == Ref.box_idx.end()) continue;
        const size_t r = it - Ref.box_idx.begin();
        if (Cur.mdynDescriptors[b].cols == 0 || Ref.mdynDescriptors[r].cols == 0) continue;
        Job j; j.box = b; j.ref = r;
        const auto& ck = Cur.mvdynKeysUn[b]; const auto& rk = Ref.mvdynKeysUn[r];
        for (const auto& k : ck) { j.qxy.push_back(k.pt.x); j.qxy.push_back(k.pt.y); }
        for (const auto& k : rk) { j.txy.push_back(k.pt.x); j.txy.push_back(k.pt.y); }
        const size_t cap = std::max<size_t>(ck.size(), 1);
        j.mq.assign(cap, -1); j.mt.assign(cap, -1); j.md.assign(cap, 0); j.fd.assign(cap, -1);
        jobs.push_back(std::move(j));
    }
    pairs.resize(jobs.size());
    std::vector<MatT> keepQ(jobs.size()), keepT(jobs.size());
    for (size_t k = 0; k < jobs.size(); ++k) {
        Job& j = jobs[k];
        /* mdynDescriptors is built row by row (Mat::push_back): take continuous copies */
        keepQ[k] = Cur.mdynDescriptors[j.box].clone(); keepT[k] = Ref.mdynDescriptors[j.ref].clone();
        sdyn_box_pair& p = pairs[k];
        p.nq = keepQ[k].rows; p.nt = keepT[k].rows;
        p.q_desc = keepQ[k].data; p.t_desc = keepT[k].data;
        p.q_xy = j.qxy.data(); p.t_xy = j.txy.data();
        p.match_query = j.mq.data(); p.match_train = j.mt.data(); p.match_dist = j.md.data(); p.false_dyn = j.fd.data();
        p.nmatches = 0;
    }
    float m[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[3 * r + c] = HorF.template at<float>(r, c);
    if (!pairs.empty() && sdyn_dyn_separate(ctx, pairs.data(), (int)pairs.size(), m, flag == 1 ? 1 : 0) != SDYN_OK) {
        report(ctx, "Tracking::Separate");
        return 0;
    }
    bool static_exit = false;
    for (size_t k = 0; k < jobs.size(); ++k) {
        const Job& j = jobs[k];
        const size_t good = (size_t)pairs[k].nmatches;
        if (good < 3 || good < 0.2 * Cur.mvdynKeys[j.box].size()) continue;
        std::vector<int>& st = dynStatus[j.box];
        st.assign(j.fd.begin(), j.fd.begin() + good);
        const int num0 = (int)st.size() - (int)std::count(st.begin(), st.end(), -1);
        if (num0 > std::max((double)1, 0.2 * good)) {
            static_exit = true;          /* `mCurrentFrame.box_status[n_box] == 1;` is a comparison in the reference (Tracking.cc:1188) */
        } else {
            const size_t last_idx = std::find(Last.box_idx.begin(), Last.box_idx.end(), Cur.box_idx[j.box]) - Last.box_idx.begin();
            const int ls = last_idx < Last.box_status.size() ? Last.box_status[last_idx] : -1;   /* the reference reads past the end here */
            Cur.box_status[j.box] = (ls == 0 || ls == 2) ? 2 : 0;
        }
    }
    return static_exit ? 1 : 0;
}

}  // namespace sdyn_host
#endif
