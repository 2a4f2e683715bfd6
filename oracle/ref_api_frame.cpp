/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * Flat C entry points over the REFERENCE'S OWN Frame (src/Frame.cc), ORBmatcher (src/ORBmatcher.cc) and
 * Tracking::Separate / classifyF / classifyH (src/Tracking.cc:1093-1367), compiled unchanged against oracle/ref_shim
 * into oracle/_ref/libref.so.  tests/test_oracle_ref.py drives them next to the restatement (oracle/orc_*.cpp) on the
 * same inputs.  Pointers of the reference come back as indices: a MapPoint* is reported as its position in the list
 * it was created from.  Compiled with -fno-access-control so the private Frame members can be called directly.
 */
#include "ref_shim/ref_entities.h"
#include "Frame.h"
#include "ORBmatcher.h"
#include "ORBextractor.h"
#include "ref_shim/ref_arena.h"
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>

using namespace ORB_SLAM2;

namespace refapi {
int tracking_separate(Frame& cur, Frame& ref, Frame& last, const cv::Mat& HorF, int flag, std::vector<std::vector<int>>& dynStatus,
                      std::vector<int>& curStatusOut);
void classify(int flag, const cv::Mat& M, const std::vector<cv::KeyPoint>& cur, const std::vector<cv::KeyPoint>& ref,
              std::vector<cv::DMatch>& matches, std::vector<int>& falseDyn);
}

namespace {

/* The reference prints progress lines to std::cout.  While any call is inside the reference's code, cout is pointed at a
 * process-lifetime null buffer (counted under a mutex: bench.py's CPU arm calls in from many threads, and a per-call sink whose
 * restore races with another thread's save leaves cout holding a dead stack buffer). */
struct NullBuf : std::streambuf {
    int overflow(int c) override { return c == EOF ? 0 : c; }
    std::streamsize xsputn(const char*, std::streamsize n) override { return n; }
};
NullBuf g_null;
std::mutex g_quietMutex;
int g_quietDepth = 0;
std::streambuf* g_quietOld = nullptr;
struct Quiet {
    Quiet() { std::lock_guard<std::mutex> l(g_quietMutex); if (g_quietDepth++ == 0) g_quietOld = std::cout.rdbuf(&g_null); }
    ~Quiet() { std::lock_guard<std::mutex> l(g_quietMutex); if (--g_quietDepth == 0) std::cout.rdbuf(g_quietOld); }
};

struct PointList {
    std::vector<std::unique_ptr<MapPoint>> owned;
    std::vector<MapPoint*> ptrs;            /* may contain NULL */
    std::map<MapPoint*, int> index;
};

struct FrameBox {
    std::unique_ptr<Frame> f;
    std::unique_ptr<KeyFrame> kf;           /* built on demand from f */
    std::shared_ptr<PointList> pts;         /* the list mvpMapPoints was last filled from */
    ORBVocabulary voc;
};

cv::Mat mat3x1(const float* p) { cv::Mat m(3, 1, CV_32F); for (int i = 0; i < 3; ++i) m.at<float>(i) = p[i]; return m; }
cv::Mat matRows(const float* p, int r, int c) { cv::Mat m(r, c, CV_32F); for (int i = 0; i < r * c; ++i) m.at<float>(i / c, i % c) = p[i]; return m; }
cv::Mat pose4x4(const float* t12)
{
    cv::Mat m = cv::Mat::eye(4, 4, CV_32F);
    for (int i = 0; i < 12; ++i) m.at<float>(i / 4, i % 4) = t12[i];
    return m;
}
cv::Mat descRow(const uint8_t* d) { cv::Mat m(1, 32, CV_8UC1); std::memcpy(m.data, d, 32); return m; }

void fill_scale_info(Frame& F, ORBextractor* ex)
{
    F.mnScaleLevels = ex->GetLevels();
    F.mfScaleFactor = ex->GetScaleFactor();
    F.mfLogScaleFactor = log(F.mfScaleFactor);          /* Frame.cc:314 */
    F.mvScaleFactors = ex->GetScaleFactors();
    F.mvInvScaleFactors = ex->GetInverseScaleFactors();
    F.mvLevelSigma2 = ex->GetScaleSigmaSquares();
    F.mvInvLevelSigma2 = ex->GetInverseScaleSigmaSquares();
}

int mp_index(const FrameBox* fb, MapPoint* p)
{
    if (!p) return -1;
    if (!fb->pts) return -2;
    auto it = fb->pts->index.find(p);
    return it == fb->pts->index.end() ? -2 : it->second;
}

}  // namespace

extern "C" {

/* ---- statics of Frame (computed once per calibration in the reference) ---- */
void ref_frame_reset_statics() { Frame::mbInitialComputations = true; Frame::nNextId = 0; }
void ref_frame_get_statics(float* out10)
{
    out10[0] = Frame::mnMinX; out10[1] = Frame::mnMinY; out10[2] = Frame::mnMaxX; out10[3] = Frame::mnMaxY;
    out10[4] = Frame::mfGridElementWidthInv; out10[5] = Frame::mfGridElementHeightInv;
    out10[6] = Frame::fx; out10[7] = Frame::fy; out10[8] = Frame::cx; out10[9] = Frame::cy;
}

/* ---- construction ---- */

/* The fork's RGB-D constructor with boxes, Frame.cc:297-403 (extraction, boxTrack, firstSeparate, undistortion, depth
 * association, tail split, grid).  depth: w*h floats or NULL (all zero).  last: a frame handle or NULL (an empty Frame). */
void* ref_frame_rgbd_boxes(void* extractor, const uint8_t* gray, int w, int h, const float* depth, const double* boxes, int nboxes,
                           void* last, const float* K4, const float* dist, int ndist, float bf, float thDepth)
{
    Quiet q;
    FrameBox* fb = new FrameBox;
    cv::Mat im(h, w, CV_8UC1, (void*)gray, (size_t)w);
    cv::Mat rgb(1, 1, CV_8UC3), mask;
    cv::Mat dep(h, w, CV_32F);
    if (depth) std::memcpy(dep.data, depth, sizeof(float) * (size_t)w * h); else dep.setTo(0);
    std::vector<cv::Rect2d> bx(nboxes);
    for (int i = 0; i < nboxes; ++i) bx[i] = cv::Rect2d(boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]);
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = K4[0]; K.at<float>(1, 1) = K4[1]; K.at<float>(0, 2) = K4[2]; K.at<float>(1, 2) = K4[3];
    cv::Mat D(ndist, 1, CV_32F);
    for (int i = 0; i < ndist; ++i) D.at<float>(i) = dist[i];
    Frame empty;
    Frame& lastF = last ? *((FrameBox*)last)->f : empty;
    refapi::BumpScope arena(true);      /* the extraction inside the constructor sees the pinned allocation order (B-1) */
    fb->f.reset(new Frame(im, rgb, dep, mask, bx, lastF, 0.0, (ORBextractor*)extractor, &fb->voc, K, D, bf, thDepth));
    fb->f->mpORBvocabulary = nullptr;
    return fb;
}

/* The stereo constructor, Frame.cc:65-128 (two extraction threads, ComputeStereoMatches, grid). */
void* ref_frame_stereo(void* exL, void* exR, const uint8_t* left, const uint8_t* right, int w, int h, const float* K4, float bf, float thDepth)
{
    Quiet q;
    FrameBox* fb = new FrameBox;
    cv::Mat imL(h, w, CV_8UC1, (void*)left, (size_t)w), imR(h, w, CV_8UC1, (void*)right, (size_t)w), rgb(1, 1, CV_8UC3);
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = K4[0]; K.at<float>(1, 1) = K4[1]; K.at<float>(0, 2) = K4[2]; K.at<float>(1, 2) = K4[3];
    cv::Mat D = cv::Mat::zeros(4, 1, CV_32F);
    /* The constructor calls ComputeStereoMatches (Frame.cc:99), which reads `mb`, before it assigns `mb = mbf/fx` (:124).  In
     * the reference the Frame is a stack temporary that lands on the previous frame's bytes, so the field still holds the
     * previous (identical) baseline; here the value is planted in the raw storage before the constructor runs. */
    void* mem = ::operator new(sizeof(Frame));
    std::memset(mem, 0, sizeof(Frame));
    reinterpret_cast<Frame*>(mem)->mb = bf / K4[0];
    fb->f.reset(new (mem) Frame(imL, imR, rgb, 0.0, (ORBextractor*)exL, (ORBextractor*)exR, &fb->voc, K, D, bf, thDepth));
    fb->f->mpORBvocabulary = nullptr;
    return fb;
}

/* A Frame filled from arrays (the members the image constructors would have produced), then Frame::AssignFeaturesToGrid
 * (Frame.cc:463-478).  bounds = mnMinX, mnMinY, mnMaxX, mnMaxY; cam = fx, fy, cx, cy, mbf, mb.  The statics are shared by
 * all frames, as in the reference. */
void* ref_frame_from_arrays(void* extractor, const cv::KeyPoint* keys, const cv::KeyPoint* keysUn, const uint8_t* desc, const float* uRight,
                            int n, const float* bounds, const float* cam, const float* tcw12)
{
    FrameBox* fb = new FrameBox;
    fb->f.reset(new Frame());
    Frame& F = *fb->f;
    F.mpORBvocabulary = nullptr; F.mpORBextractorLeft = (ORBextractor*)extractor; F.mpORBextractorRight = nullptr;
    F.mTimeStamp = 0; F.mnId = Frame::nNextId++; F.mpReferenceKF = nullptr; F.mThDepth = 0; F.N_d = 0; F.N_ori = 0;
    fill_scale_info(F, (ORBextractor*)extractor);
    F.N = n;
    F.mvKeys.assign(keys, keys + n);
    F.mvKeysUn.assign(keysUn ? keysUn : keys, (keysUn ? keysUn : keys) + n);
    F.mDescriptors = cv::Mat(n, 32, CV_8UC1);
    if (n) std::memcpy(F.mDescriptors.data, desc, (size_t)32 * n);
    F.mvuRight.assign(n, -1.f);
    if (uRight) F.mvuRight.assign(uRight, uRight + n);
    F.mvDepth.assign(n, -1.f);
    F.mvpMapPoints.assign(n, nullptr);
    F.mvbOutlier.assign(n, false);
    Frame::mnMinX = bounds[0]; Frame::mnMinY = bounds[1]; Frame::mnMaxX = bounds[2]; Frame::mnMaxY = bounds[3];
    /* Frame.cc:383-384 */
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3];
    Frame::invfx = 1.0f / Frame::fx; Frame::invfy = 1.0f / Frame::fy;
    Frame::mbInitialComputations = false;
    F.mbf = cam[4]; F.mb = cam[5];
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.mDistCoef = cv::Mat::zeros(4, 1, CV_32F);
    F.AssignFeaturesToGrid();
    if (tcw12) F.SetPose(pose4x4(tcw12));
    return fb;
}

void ref_frame_destroy(void* f) { delete (FrameBox*)f; }

/* ---- read-back ---- */
int ref_frame_n(void* f) { return ((FrameBox*)f)->f->mvKeys.size(); }
int ref_frame_n_right(void* f) { return ((FrameBox*)f)->f->mvKeysRight.size(); }
void ref_frame_keys(void* f, int which, cv::KeyPoint* out)
{
    Frame& F = *((FrameBox*)f)->f;
    const std::vector<cv::KeyPoint>& v = which == 0 ? F.mvKeys : which == 1 ? F.mvKeysUn : F.mvKeysRight;
    if (!v.empty()) std::memcpy(out, v.data(), sizeof(cv::KeyPoint) * v.size());
}
void ref_frame_descriptors(void* f, int right, uint8_t* out)
{
    Frame& F = *((FrameBox*)f)->f;
    const cv::Mat& D = right ? F.mDescriptorsRight : F.mDescriptors;
    for (int i = 0; i < D.rows; ++i) std::memcpy(out + 32 * (size_t)i, D.ptr(i), 32);
}
void ref_frame_stereo_values(void* f, float* uRight, float* depth)
{
    Frame& F = *((FrameBox*)f)->f;
    for (size_t i = 0; i < F.mvuRight.size(); ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}
/* per cell (x-major: cell = ix*48 + iy) the entry count, then the concatenated entries in cell order */
int ref_frame_grid(void* f, int* counts, int* entries, int cap)
{
    Frame& F = *((FrameBox*)f)->f;
    int n = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
        for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
            counts[ix * FRAME_GRID_ROWS + iy] = (int)F.mGrid[ix][iy].size();
            for (size_t v : F.mGrid[ix][iy]) { if (n < cap) entries[n] = (int)v; ++n; }
        }
    return n;
}
int ref_frame_features_in_area(void* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap)
{
    const std::vector<size_t> v = ((FrameBox*)f)->f->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}

/* boxes after boxTrack + firstSeparate: objects (4 doubles), box_idx, omit, velocity (2 doubles), status */
int ref_frame_boxes(void* f, double* objects, int* box_idx, uint8_t* omit, double* vel, int* status, int cap)
{
    Frame& F = *((FrameBox*)f)->f;
    const int n = (int)F.objects.size();
    for (int i = 0; i < n && i < cap; ++i) {
        objects[4 * i] = F.objects[i].x; objects[4 * i + 1] = F.objects[i].y; objects[4 * i + 2] = F.objects[i].width; objects[4 * i + 3] = F.objects[i].height;
        box_idx[i] = F.box_idx[i]; omit[i] = F.omit[i];
        vel[2 * i] = F.box_velocity[i].x; vel[2 * i + 1] = F.box_velocity[i].y;
        status[i] = i < (int)F.box_status.size() ? F.box_status[i] : -99;
    }
    return n;
}
void ref_frame_set_box_status(void* f, const int* status, int n) { ((FrameBox*)f)->f->box_status.assign(status, status + n); }
int ref_frame_n_dyn(void* f) { return ((FrameBox*)f)->f->N_d; }
int ref_frame_dyn_count(void* f, int box)
{
    Frame& F = *((FrameBox*)f)->f;
    return box < (int)F.mvdynKeys.size() ? (int)F.mvdynKeys[box].size() : 0;
}
void ref_frame_dyn(void* f, int box, cv::KeyPoint* keys, cv::KeyPoint* keysUn, uint8_t* desc, float* uRight, float* depth)
{
    Frame& F = *((FrameBox*)f)->f;
    const size_t n = F.mvdynKeys[box].size();
    for (size_t i = 0; i < n; ++i) {
        keys[i] = F.mvdynKeys[box][i]; keysUn[i] = F.mvdynKeysUn[box][i];
        std::memcpy(desc + 32 * i, F.mdynDescriptors[box].ptr((int)i), 32);
        uRight[i] = F.mvudynRight[box][i]; depth[i] = F.mvdynDepth[box][i];
    }
}

/* ---- pose / map points ---- */
void ref_frame_set_pose(void* f, const float* tcw12) { ((FrameBox*)f)->f->SetPose(pose4x4(tcw12)); }

void* ref_points_create(int n, const uint8_t* present, const float* world, const float* normal, const uint8_t* desc, const float* minDist,
                        const float* maxDist, const int* nobs, const uint8_t* bad)
{
    auto* pl = new std::shared_ptr<PointList>(new PointList);
    PointList& L = **pl;
    L.ptrs.assign(n, nullptr);
    const float zero3[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        if (present && !present[i]) continue;
        L.owned.emplace_back(new MapPoint(mat3x1(world + 3 * i), mat3x1(normal ? normal + 3 * i : zero3), descRow(desc + 32 * (size_t)i),
                                          minDist ? minDist[i] : 0.f, maxDist ? maxDist[i] : 0.f, nobs ? nobs[i] : 1, bad ? bad[i] != 0 : false));
        L.ptrs[i] = L.owned.back().get();
        L.index[L.ptrs[i]] = i;
    }
    return pl;
}
void ref_points_destroy(void* p) { delete (std::shared_ptr<PointList>*)p; }
/* the per-frame tracking fields Frame::isInFrustum would have written (Frame.cc:718-727) */
void ref_points_set_track(void* p, const uint8_t* inView, const float* projX, const float* projY, const float* projXR, const int* level, const float* viewCos)
{
    PointList& L = **(std::shared_ptr<PointList>*)p;
    for (size_t i = 0; i < L.ptrs.size(); ++i) {
        MapPoint* m = L.ptrs[i];
        if (!m) continue;
        m->mbTrackInView = inView[i] != 0; m->mTrackProjX = projX[i]; m->mTrackProjY = projY[i]; m->mTrackProjXR = projXR[i];
        m->mnTrackScaleLevel = level[i]; m->mTrackViewCos = viewCos[i];
    }
}
/* Frame::isInFrustum on every point of the list (Frame.cc:677-733): returns the flags and the fields it wrote */
void ref_points_in_frustum(void* f, void* p, float viewingCosLimit, uint8_t* inView, float* projX, float* projY, float* projXR, int* level, float* viewCos)
{
    Frame& F = *((FrameBox*)f)->f;
    PointList& L = **(std::shared_ptr<PointList>*)p;
    for (size_t i = 0; i < L.ptrs.size(); ++i) {
        MapPoint* m = L.ptrs[i];
        inView[i] = 0;
        if (!m) continue;
        inView[i] = F.isInFrustum(m, viewingCosLimit);
        projX[i] = m->mTrackProjX; projY[i] = m->mTrackProjY; projXR[i] = m->mTrackProjXR; level[i] = m->mnTrackScaleLevel; viewCos[i] = m->mTrackViewCos;
    }
}
/* F.mvpMapPoints[i] = list[i] (a frame that owns one map point per keypoint, like LastFrame); outlier flags optional */
void ref_frame_set_points(void* f, void* p, const uint8_t* outlier)
{
    FrameBox* fb = (FrameBox*)f;
    fb->pts = *(std::shared_ptr<PointList>*)p;
    Frame& F = *fb->f;
    F.mvpMapPoints = fb->pts->ptrs;
    F.mvpMapPoints.resize(F.N, nullptr);
    F.mvbOutlier.assign(F.N, false);
    if (outlier) for (int i = 0; i < F.N; ++i) F.mvbOutlier[i] = outlier[i] != 0;
}
/* which list the pointers found in mvpMapPoints are reported against */
void ref_frame_report_against(void* f, void* p) { ((FrameBox*)f)->pts = *(std::shared_ptr<PointList>*)p; }
void ref_frame_clear_points(void* f) { Frame& F = *((FrameBox*)f)->f; F.mvpMapPoints.assign(F.N, nullptr); }
/* pre-occupy keypoints: assign[i] >= 0 -> list[assign[i]] */
void ref_frame_preassign(void* f, void* p, const int* assign)
{
    FrameBox* fb = (FrameBox*)f;
    fb->pts = *(std::shared_ptr<PointList>*)p;
    Frame& F = *fb->f;
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = assign[i] >= 0 ? fb->pts->ptrs[assign[i]] : nullptr;
}
void ref_frame_assignment(void* f, int* out)
{
    FrameBox* fb = (FrameBox*)f;
    for (int i = 0; i < fb->f->N; ++i) out[i] = mp_index(fb, fb->f->mvpMapPoints[i]);
}
void ref_frame_set_featvec(void* f, int nnodes, const uint32_t* nodeId, const int* offset, const uint32_t* index)
{
    Frame& F = *((FrameBox*)f)->f;
    F.mFeatVec.clear();
    for (int k = 0; k < nnodes; ++k)
        for (int j = offset[k]; j < offset[k + 1]; ++j) F.mFeatVec.addFeature(nodeId[k], index[j]);
}

/* ---- ORBmatcher ---- */
int ref_descriptor_distance(const uint8_t* a, const uint8_t* b) { return ORBmatcher::DescriptorDistance(descRow(a), descRow(b)); }

/* ORBmatcher.cc:45-129 */
int ref_search_by_projection_map(void* f, void* points, float th, float nnratio)
{
    FrameBox* fb = (FrameBox*)f;
    PointList& L = **(std::shared_ptr<PointList>*)points;
    std::vector<MapPoint*> v;
    for (MapPoint* m : L.ptrs) if (m) v.push_back(m);   /* the reference's list never holds NULL */
    ORBmatcher matcher(nnratio, true);
    return matcher.SearchByProjection(*fb->f, v, th);
}

/* ORBmatcher.cc:1485-1627 (wantPairs = 0) and the fork's overload :407-559 (wantPairs = 1: pairs = last.x, last.y, cur.x, cur.y) */
int ref_search_by_projection_frame(void* cur, void* last, float th, int mono, float nnratio, int checkOri, int wantPairs, float* pairs, int cap, int* npairs)
{
    Quiet q;
    Frame& C = *((FrameBox*)cur)->f;
    Frame& Lf = *((FrameBox*)last)->f;
    ORBmatcher matcher(nnratio, checkOri != 0);
    if (!wantPairs) return matcher.SearchByProjection(C, Lf, th, mono != 0);
    std::vector<cv::Point2f> pl, pc;
    const int n = matcher.SearchByProjection(C, Lf, th, mono != 0, pl, pc);
    *npairs = (int)pl.size();
    for (size_t i = 0; i < pl.size() && (int)i < cap; ++i) { pairs[4 * i] = pl[i].x; pairs[4 * i + 1] = pl[i].y; pairs[4 * i + 2] = pc[i].x; pairs[4 * i + 3] = pc[i].y; }
    return n;
}

/* ORBmatcher.cc:562-677 */
int ref_search_for_initialization(void* f1, void* f2, float* prevMatched, int* matches12, int windowSize, float nnratio, int checkOri)
{
    Frame& F1 = *((FrameBox*)f1)->f;
    Frame& F2 = *((FrameBox*)f2)->f;
    std::vector<cv::Point2f> prev(F1.mvKeysUn.size());
    for (size_t i = 0; i < prev.size(); ++i) prev[i] = cv::Point2f(prevMatched[2 * i], prevMatched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
    for (size_t i = 0; i < m12.size(); ++i) matches12[i] = m12[i];
    for (size_t i = 0; i < prev.size(); ++i) { prevMatched[2 * i] = prev[i].x; prevMatched[2 * i + 1] = prev[i].y; }
    return n;
}

/* the frame seen as a KeyFrame (KeyFrame.cc:30-58): copies keys, descriptors, grid, FeatureVector, mvpMapPoints, pose */
void ref_frame_make_keyframe(void* f)
{
    FrameBox* fb = (FrameBox*)f;
    if (fb->f->mTcw.empty()) fb->f->SetPose(cv::Mat::eye(4, 4, CV_32F));
    fb->kf.reset(new KeyFrame(*fb->f, nullptr, nullptr));
}

/* ORBmatcher.cc:159-288: assign[fIdx] = kfIdx | -1 */
int ref_search_by_bow_frame(void* kf, void* f, float nnratio, int checkOri, int* assign)
{
    FrameBox* K = (FrameBox*)kf;
    FrameBox* Fb = (FrameBox*)f;
    std::vector<MapPoint*> matches;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(K->kf.get(), *Fb->f, matches);
    std::map<MapPoint*, int> where;
    for (int i = 0; i < K->kf->N; ++i) if (K->kf->mvpMapPoints[i]) where[K->kf->mvpMapPoints[i]] = i;
    for (size_t i = 0; i < matches.size(); ++i) assign[i] = matches[i] ? where[matches[i]] : -1;
    return n;
}

/* ORBmatcher.cc:679-812: matches12[idx1] = idx2 | -1 */
int ref_search_by_bow_kf(void* kf1, void* kf2, float nnratio, int checkOri, int* matches12)
{
    FrameBox* A = (FrameBox*)kf1;
    FrameBox* B = (FrameBox*)kf2;
    std::vector<MapPoint*> m12;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByBoW(A->kf.get(), B->kf.get(), m12);
    std::map<MapPoint*, int> where;
    for (int i = 0; i < B->kf->N; ++i) if (B->kf->mvpMapPoints[i]) where[B->kf->mvpMapPoints[i]] = i;
    for (size_t i = 0; i < m12.size(); ++i) matches12[i] = m12[i] ? where[m12[i]] : -1;
    return n;
}

/* ORBmatcher.cc:814-980: matches12[idx1] = idx2 | -1 from vMatchedPairs */
int ref_search_for_triangulation(void* kf1, void* kf2, const float* F12, int onlyStereo, float nnratio, int checkOri, int* matches12)
{
    FrameBox* A = (FrameBox*)kf1;
    FrameBox* B = (FrameBox*)kf2;
    std::vector<std::pair<size_t, size_t>> pairs;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchForTriangulation(A->kf.get(), B->kf.get(), matRows(F12, 3, 3), pairs, onlyStereo != 0);
    for (int i = 0; i < A->kf->N; ++i) matches12[i] = -1;
    for (auto& p : pairs) matches12[p.first] = (int)p.second;
    return n;
}

/* ORBmatcher.cc:1629-1756: candidates = map points of the keyframe `kf`; assign[curIdx] = kf keypoint index */
int ref_search_by_projection_reloc(void* cur, void* kf, float th, int ORBdist, float nnratio, int checkOri, int* assign)
{
    FrameBox* C = (FrameBox*)cur;
    FrameBox* K = (FrameBox*)kf;
    std::set<MapPoint*> found;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int n = matcher.SearchByProjection(*C->f, K->kf.get(), found, th, ORBdist);
    std::map<MapPoint*, int> where;
    for (int i = 0; i < K->kf->N; ++i) if (K->kf->mvpMapPoints[i]) where[K->kf->mvpMapPoints[i]] = i;
    for (int i = 0; i < C->f->N; ++i) assign[i] = C->f->mvpMapPoints[i] ? where[C->f->mvpMapPoints[i]] : -1;
    return n;
}

/* ORBmatcher.cc:290-403: vpMatched[kfIdx] = point index | -1 */
int ref_search_by_projection_sim3(void* kf, const float* Scw16, void* points, int th, float nnratio, int* matched)
{
    FrameBox* K = (FrameBox*)kf;
    PointList& L = **(std::shared_ptr<PointList>*)points;
    static MapPoint dummy(mat3x1(Scw16), mat3x1(Scw16), descRow((const uint8_t*)"0123456789012345678901234567890123"), 0, 0, 1, false);
    std::vector<MapPoint*> vpMatched(K->kf->N, nullptr);
    for (int i = 0; i < K->kf->N; ++i) if (matched[i] == -2) vpMatched[i] = &dummy;      /* in: -2 = already matched elsewhere */
    ORBmatcher matcher(nnratio, true);
    std::vector<MapPoint*> v;
    for (MapPoint* m : L.ptrs) if (m) v.push_back(m);      /* the reference's list never holds NULL (no check at :318) */
    const int n = matcher.SearchByProjection(K->kf.get(), matRows(Scw16, 4, 4), v, vpMatched, th);
    for (int i = 0; i < K->kf->N; ++i) {
        auto it = L.index.find(vpMatched[i]);
        matched[i] = !vpMatched[i] ? -1 : it != L.index.end() ? it->second : -2;
    }
    return n;
}

/* ORBmatcher.cc:982-1130 Fuse(pKF, vpMapPoints, th) on a keyframe whose keypoints are all free and candidates without
 * observations: a fused candidate either gets AddObservation(pKF, bestIdx) or, when an earlier candidate already sits
 * there, is Replace()d by it — both reveal bestIdx.  outIdx[i] = bestIdx | -1. */
int ref_fuse(void* kf, void* points, float th, float nnratio, int* outIdx)
{
    FrameBox* K = (FrameBox*)kf;
    PointList& L = **(std::shared_ptr<PointList>*)points;
    K->kf->mvpMapPoints.assign(K->kf->N, nullptr);
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.Fuse(K->kf.get(), L.ptrs, th);
    for (size_t i = 0; i < L.ptrs.size(); ++i) {
        MapPoint* m = L.ptrs[i];
        outIdx[i] = -1;
        if (!m) continue;
        if (m->mObservations.count(K->kf.get())) outIdx[i] = (int)m->mObservations[K->kf.get()];
        else if (m->mpReplaced && m->mpReplaced->mObservations.count(K->kf.get())) outIdx[i] = (int)m->mpReplaced->mObservations[K->kf.get()];
    }
    return n;
}

/* ORBmatcher.cc:1132-1257 Fuse(pKF, Scw, vpPoints, th, vpReplacePoint), same set-up: outIdx[i] = bestIdx | -1 */
int ref_fuse_sim3(void* kf, const float* Scw16, void* points, float th, float nnratio, int* outIdx)
{
    FrameBox* K = (FrameBox*)kf;
    PointList& L = **(std::shared_ptr<PointList>*)points;
    K->kf->mvpMapPoints.assign(K->kf->N, nullptr);
    std::vector<MapPoint*> v;
    std::vector<int> at;
    for (size_t i = 0; i < L.ptrs.size(); ++i) if (L.ptrs[i]) { v.push_back(L.ptrs[i]); at.push_back((int)i); }   /* no NULL check at :1163 */
    std::vector<MapPoint*> repl(v.size(), nullptr);
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.Fuse(K->kf.get(), matRows(Scw16, 4, 4), v, th, repl);
    for (size_t i = 0; i < L.ptrs.size(); ++i) outIdx[i] = -1;
    for (size_t c = 0; c < v.size(); ++c) {
        MapPoint* m = v[c];
        if (m->mObservations.count(K->kf.get())) outIdx[at[c]] = (int)m->mObservations[K->kf.get()];
        else if (repl[c] && repl[c]->mObservations.count(K->kf.get())) outIdx[at[c]] = (int)repl[c]->mObservations[K->kf.get()];
    }
    return n;
}

/* ORBmatcher.cc:1259-1483: matches12[idx1] = idx2 | -1 (in: -2 = already matched to something else) */
int ref_search_by_sim3(void* kf1, void* kf2, int* matches12, float s12, const float* R12, const float* t12, float th, float nnratio)
{
    FrameBox* A = (FrameBox*)kf1;
    FrameBox* B = (FrameBox*)kf2;
    static MapPoint dummy(mat3x1(R12), mat3x1(R12), descRow((const uint8_t*)"0123456789012345678901234567890123"), 0, 0, 1, false);
    std::vector<MapPoint*> m12(A->kf->N, nullptr);
    for (int i = 0; i < A->kf->N; ++i) if (matches12[i] == -2) m12[i] = &dummy;
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchBySim3(A->kf.get(), B->kf.get(), m12, s12, matRows(R12, 3, 3), matRows(t12, 3, 1), th);
    std::map<MapPoint*, int> where;
    for (int i = 0; i < B->kf->N; ++i) if (B->kf->mvpMapPoints[i]) where[B->kf->mvpMapPoints[i]] = i;
    for (int i = 0; i < A->kf->N; ++i) matches12[i] = !m12[i] ? -1 : m12[i] == &dummy ? -2 : where[m12[i]];
    return n;
}

/* ---- Frame::UpdateFrame (Frame.cc:607-653) and Tracking::Separate (Tracking.cc:1093-1239) ---- */
/* dynStatus as CSR: off[nboxes+1], values */
void ref_frame_update(void* f, const int* off, const int* values, int nboxes)
{
    Frame& F = *((FrameBox*)f)->f;
    std::vector<std::vector<int>> ds(nboxes);
    for (int b = 0; b < nboxes; ++b) ds[b].assign(values + off[b], values + off[b + 1]);
    F.UpdateFrame(ds);
}

/* Returns Separate's result; dynStatus comes back as CSR (off has nboxes+1 entries) and the current frame's box_status
 * after the call in statusOut. */
int ref_tracking_separate(void* cur, void* ref, void* last, const float* HorF, int flag, int* off, int* values, int cap, int* statusOut)
{
    Quiet q;
    Frame& C = *((FrameBox*)cur)->f;
    std::vector<std::vector<int>> ds;
    std::vector<int> st;
    const int r = refapi::tracking_separate(C, *((FrameBox*)ref)->f, *((FrameBox*)last)->f, matRows(HorF, 3, 3), flag, ds, st);
    int n = 0;
    for (size_t b = 0; b < ds.size(); ++b) {
        off[b] = n;
        for (int v : ds[b]) { if (n < cap) values[n] = v; ++n; }
    }
    off[ds.size()] = n;
    for (size_t b = 0; b < st.size(); ++b) statusOut[b] = st[b];
    return r;
}

/* classifyF (flag 2) / classifyH (flag 1) on explicit pairs: curXY / refXY are 2 floats per keypoint */
void ref_classify(int flag, const float* M, const float* curXY, int ncur, const float* refXY, int nref, const int* query, const int* train, int nm, int* falseDyn)
{
    std::vector<cv::KeyPoint> c(ncur), r(nref);
    for (int i = 0; i < ncur; ++i) c[i].pt = cv::Point2f(curXY[2 * i], curXY[2 * i + 1]);
    for (int i = 0; i < nref; ++i) r[i].pt = cv::Point2f(refXY[2 * i], refXY[2 * i + 1]);
    std::vector<cv::DMatch> m(nm);
    for (int i = 0; i < nm; ++i) { m[i].queryIdx = query[i]; m[i].trainIdx = train[i]; }
    std::vector<int> fd(nm, -1);
    refapi::classify(flag, matRows(M, 3, 3), c, r, m, fd);
    for (int i = 0; i < nm; ++i) falseDyn[i] = fd[i];
}

/* ---- probes of the shim's own arithmetic (checked against cv2 in tests/test_oracle_ref.py) ---- */
void ref_cv_gemm(const float* A, int ar, int ac, int aT, const float* B, int br, int bc, double alpha, const float* Cm, double beta, float* out)
{
    cv::Mat a = matRows(A, ar, ac), b = matRows(B, br, bc), c;
    if (Cm) c = matRows(Cm, aT ? ac : ar, bc);
    cv::Mat d = cv::minicv_gemm(a, b, alpha, c, beta, aT ? 1 : 0);
    for (int i = 0; i < d.rows * d.cols; ++i) out[i] = d.at<float>(i / d.cols, i % d.cols);
}
void ref_cv_invert3x3(const float* M, float* out)
{
    cv::Mat d = matRows(M, 3, 3).inv();
    for (int i = 0; i < 9; ++i) out[i] = d.at<float>(i / 3, i % 3);
}
double ref_cv_norm(const float* v, int n) { return cv::norm(matRows(v, n, 1)); }
double ref_cv_dot(const float* a, const float* b, int n) { return matRows(a, n, 1).dot(matRows(b, n, 1)); }
int ref_cv_bfmatch(const uint8_t* q, int nq, const uint8_t* t, int nt, int* query, int* train, int* dist)
{
    cv::Mat Q(nq, 32, CV_8UC1, (void*)q), T(nt, 32, CV_8UC1, (void*)t);
    std::vector<cv::DMatch> m;
    cv::BFMatcher(cv::NORM_HAMMING, true).match(Q, T, m);
    for (size_t i = 0; i < m.size(); ++i) { query[i] = m[i].queryIdx; train[i] = m[i].trainIdx; dist[i] = (int)m[i].distance; }
    return (int)m.size();
}

}  // extern "C"
