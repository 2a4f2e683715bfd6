/* sdyn — B200-native tracking front end for li-guihai/slam-dynamic: the C-ABI drop-in boundary.
 *
 * This is the ONLY interface between the reference-facing C++ adapter classes
 * (slam-dynamic_b200/host/: ORB_SLAM2::ORBextractor, ORBmatcher, Frame helpers — same signatures
 * as the reference) and the hand-written sm_100a CUDA kernels (slam-dynamic_b200/csrc/).
 * Plain C types only: int, float, pointers, POD structs.  Every function returns 0 on success and a
 * negative sdyn_status on failure; nothing throws, nothing calls exit().  There is no CPU fallback:
 * when no CUDA device is usable sdyn_create() fails with SDYN_ERR_CUDA.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference tree).
 */
#ifndef SDYN_H_
#define SDYN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDYN_MAX_LEVELS 16
#define SDYN_EDGE 19            /* EDGE_THRESHOLD, src/ORBextractor.cc:74 — border of every pyramid level */
#define SDYN_GRID_COLS 64       /* FRAME_GRID_COLS, include/Frame.h:40 */
#define SDYN_GRID_ROWS 48       /* FRAME_GRID_ROWS, include/Frame.h:39 */
#define SDYN_TH_HIGH 100        /* ORBmatcher::TH_HIGH, src/ORBmatcher.cc:37 */
#define SDYN_TH_LOW 50          /* ORBmatcher::TH_LOW,  src/ORBmatcher.cc:38 */
#define SDYN_HISTO_LENGTH 30    /* ORBmatcher::HISTO_LENGTH, src/ORBmatcher.cc:39 */

typedef enum {
    SDYN_OK = 0,
    SDYN_ERR_ARG = -1,          /* bad argument (null pointer, size out of the limits given to sdyn_create) */
    SDYN_ERR_CUDA = -2,         /* CUDA runtime error or no device; sdyn_last_error() has the text */
    SDYN_ERR_CAPACITY = -3,     /* caller's output capacity too small; required count is still reported */
    SDYN_ERR_GEOMETRY = -4,     /* image shape on which the reference itself is undefined */
    SDYN_ERR_NOMEM = -5
} sdyn_status;

/* Layout-identical to cv::KeyPoint (28 bytes): the adapter reinterpret_casts between the two. */
typedef struct {
    float x, y;                 /* pt, level-0 coordinates */
    float size;                 /* int(31 * scale[octave]) */
    float angle;                /* degrees, [0,360) */
    float response;             /* FAST score */
    int32_t octave;
    int32_t class_id;           /* -1 out of the extractor */
} sdyn_keypoint;

/* Constructor arguments of ORBextractor (include/ORBextractor.h:51-52). */
typedef struct {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
} sdyn_orb_params;

/* Values of the getters GetScaleFactors()/GetInverseScaleFactors()/GetScaleSigmaSquares()/
 * GetInverseScaleSigmaSquares() (include/ORBextractor.h:60-78) and of mnFeaturesPerLevel. */
typedef struct {
    int32_t nlevels;
    float scale[SDYN_MAX_LEVELS];
    float inv_scale[SDYN_MAX_LEVELS];
    float sigma2[SDYN_MAX_LEVELS];
    float inv_sigma2[SDYN_MAX_LEVELS];
    int32_t features_per_level[SDYN_MAX_LEVELS];
} sdyn_scale_info;

/* Geometry of one pyramid level for the current image size (ComputePyramid, src/ORBextractor.cc:1107-1132). */
typedef struct {
    int32_t width, height;      /* un-bordered level size */
    int32_t pitch;              /* bytes per row of the bordered device buffer */
    int32_t reserved;
    size_t offset;              /* byte offset of bordered pixel (-19,-19) inside one frame's pyramid block */
} sdyn_level_info;

typedef struct sdyn_ctx sdyn_ctx;

/* ---- context ------------------------------------------------------------------------------------
 * One context per ORBextractor instance (the reference runs the left and right instances on two
 * threads, src/Frame.cc:151-154: use two contexts).  Owns its device buffers, pinned staging and
 * stream; after sdyn_create nothing is allocated per frame.  max_batch = frames processed per call
 * by the *_batch entry points (1 for the plain drop-in). */
int sdyn_create(const sdyn_orb_params* params, int max_width, int max_height, int max_batch,
                int device, sdyn_ctx** out);
int sdyn_destroy(sdyn_ctx* ctx);
/* Text of the last error on this context (never NULL).  sdyn_create failures: pass NULL. */
const char* sdyn_last_error(const sdyn_ctx* ctx);
int sdyn_scale_info_get(const sdyn_ctx* ctx, sdyn_scale_info* out);
/* Capacity needed for the keypoint / descriptor outputs of one frame: the extractor can return a few
 * more than nfeatures (the octree overshoots by up to 3 per level). */
int sdyn_max_keypoints(const sdyn_ctx* ctx);

/* Host-only: the tables the ORBextractor constructor builds (src/ORBextractor.cc:410-470) — scale factors,
 * per-level quotas and the 16 half-widths of the orientation disc.  Needs no device and no context. */
int sdyn_orb_tables(const sdyn_orb_params* params, sdyn_scale_info* out, int32_t umax[16]);

/* Pinned host memory helpers (optional; any host pointer works, pinned ones copy asynchronously). */
int sdyn_host_alloc(void** ptr, size_t bytes);
int sdyn_host_free(void* ptr);

/* ---- ORBextractor::operator() -----------------------------------------------------------------
 * Replaces ORBextractor::operator()(image, mask, keypoints, descriptors), src/ORBextractor.cc:1043-1105.
 * gray: 8-bit single channel, `stride` bytes per row (host memory).  Writes at most `cap` keypoints
 * (level-major, octree list order) and 32-byte descriptors; *n_out receives the true count.
 * An empty image (w==0 || h==0 || gray==NULL) is a silent no-op with *n_out = 0, as in the reference.
 * pyr_out (nullable): SDYN_MAX_LEVELS host pointers, each receiving the bordered level
 * ((w_l+38) x (h_l+38), tightly packed) — the public mvImagePyramid member (include/ORBextractor.h:85). */
int sdyn_extract(sdyn_ctx* ctx, const uint8_t* gray, int width, int height, int stride,
                 sdyn_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out,
                 uint8_t* const* pyr_out);

/* Batched form: nframes (<= max_batch) images of identical size, frame f at gray + f*frame_stride.
 * Outputs: frame f at kp_out + f*cap and desc_out + f*cap*32; n_out[f]. */
int sdyn_extract_batch(sdyn_ctx* ctx, int nframes, const uint8_t* gray, size_t frame_stride,
                       int width, int height, int stride,
                       sdyn_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out);

/* Device-resident form: d_gray is a device pointer; results stay on the device (see
 * sdyn_device_results) until sdyn_fetch_results.  `stream` is a cudaStream_t (NULL = the context's own
 * stream); the call only enqueues work. */
int sdyn_extract_batch_device(sdyn_ctx* ctx, int nframes, const uint8_t* d_gray, size_t frame_stride,
                              int width, int height, int stride, void* stream);
typedef struct {
    const sdyn_keypoint* kp;    /* [max_batch][cap] */
    const uint8_t* desc;        /* [max_batch][cap][32] */
    const int32_t* count;       /* [max_batch] */
    const int32_t* level_count; /* [max_batch][SDYN_MAX_LEVELS] keypoints kept per level */
    const uint8_t* pyramid;     /* [max_batch] frames of `pyramid_frame_bytes`, levels at sdyn_level_info.offset */
    const uint8_t* blurred;     /* same geometry, GaussianBlur(7x7, sigma 2) of every level (interior only) */
    size_t pyramid_frame_bytes;
    int32_t cap;
    int32_t nlevels;
    sdyn_level_info level[SDYN_MAX_LEVELS];
} sdyn_device_view;
int sdyn_device_results(const sdyn_ctx* ctx, sdyn_device_view* out);
int sdyn_fetch_results(sdyn_ctx* ctx, int nframes, sdyn_keypoint* kp_out, uint8_t* desc_out, int cap,
                       int* n_out, void* stream);
/* Copies bordered level `level` of frame `frame` of the last batch to host, tightly packed. */
int sdyn_fetch_level(sdyn_ctx* ctx, int frame, int level, uint8_t* out, int* width, int* height);
/* Stage-level introspection for parity tests: FAST candidates of (frame, level) after the per-cell
 * threshold fallback, as (x, y, response) int triples in level coordinates relative to the FAST window
 * origin (minBorderX/Y = 16).  Order is unspecified (the selection stage is order-independent). */
int sdyn_fetch_candidates(sdyn_ctx* ctx, int frame, int level, int32_t* xyv, int cap, int* n_out);
/* Latency mode of one-frame contexts (max_batch == 1, the drop-in ORBextractor): sdyn_extract / sdyn_extract_batch(1 frame)
 * replay the whole call as ONE CUDA graph captured per image size (on by default; 0 switches back to stream launches).
 * Outputs are bit-identical either way. */
int sdyn_set_latency_mode(sdyn_ctx* ctx, int on);
/* The context's own stream (a cudaStream_t), for ordering caller-side copies (e.g. sdyn_map_update) with its steps. */
void* sdyn_stream(sdyn_ctx* ctx);
/* Blocks until all work enqueued on the context's stream has finished. */
int sdyn_sync(sdyn_ctx* ctx);
/* Number of kernel launches enqueued by this context since creation (bench.py's gpu_launches). */
long long sdyn_launch_count(const sdyn_ctx* ctx);

/* ---- ORBmatcher ----------------------------------------------------------------------------------
 * Pointers of the reference become indices across the ABI: Frame::mvpMapPoints[i] is `assign[i]`
 * (-1 = NULL, otherwise the index of the query that claimed keypoint i) and "the occupant has
 * Observations() > 0" (src/ORBmatcher.cc:87-89, :1560-1562) is `locked[i]`.  Both arrays are in/out:
 * the adapter fills them from the frame before the call and maps indices back to MapPoint* after it. */

/* What the searches read from a Frame (include/Frame.h:49-211). */
typedef struct {
    int32_t n;                      /* N */
    int32_t nlevels;                /* mnScaleLevels */
    const sdyn_keypoint* keys;      /* mvKeys */
    const sdyn_keypoint* keys_un;   /* mvKeysUn */
    const uint8_t* desc;            /* mDescriptors, n x 32 */
    const float* u_right;           /* mvuRight, or NULL (monocular: all -1) */
    const float* scale_factors;     /* mvScaleFactors */
    float min_x, min_y, max_x, max_y;   /* mnMinX, mnMinY, mnMaxX, mnMaxY */
    float fx, fy, cx, cy, bf, b;    /* fx, fy, cx, cy, mbf, mb */
    float tcw[12];                  /* rows 0..2 of mTcw, row-major */
} sdyn_frame_view;

/* MapPoint fields read by SearchByProjection(Frame&, vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129 */
typedef struct {
    float proj_x, proj_y, proj_xr;  /* mTrackProjX, mTrackProjY, mTrackProjXR */
    float view_cos;                 /* mTrackViewCos */
    int32_t level;                  /* mnTrackScaleLevel */
    uint8_t track_in_view, bad, obs_positive, pad;   /* mbTrackInView, isBad(), Observations() > 0 */
    uint8_t desc[32];               /* GetDescriptor() */
} sdyn_mappoint_query;

/* Per keypoint of LastFrame, read by SearchByProjection(Cur, Last, th, bMono), src/ORBmatcher.cc:1485-1627 */
typedef struct {
    uint8_t has_mp, outlier, obs_positive, pad;   /* mvpMapPoints[i] != NULL, mvbOutlier[i], Observations() > 0 */
    float world[3];                 /* pMP->GetWorldPos() */
    uint8_t desc[32];               /* pMP->GetDescriptor() */
} sdyn_last_point;

/* MapPoint fields read by the two pose-projection searches (src/ORBmatcher.cc:290-403, 1629-1756). */
typedef struct {
    uint8_t valid, pad[3];          /* pMP != NULL && !isBad() && not in sAlreadyFound / spAlreadyFound */
    float world[3];                 /* GetWorldPos() */
    float normal[3];                /* GetNormal() (KeyFrame/Scw overload only) */
    float min_distance, max_distance;   /* GetMinDistanceInvariance(), GetMaxDistanceInvariance() */
    float max_distance_raw;         /* mfMaxDistance, the numerator of MapPoint::PredictScale (src/MapPoint.cc:385-418) */
    float angle;                    /* Frame/KeyFrame overload: pKF->mvKeysUn[i].angle */
    uint8_t desc[32];               /* GetDescriptor() */
} sdyn_proj_point;

enum { SDYN_PROJ_FRAME_KEYFRAME = 0,   /* SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*>&, th, ORBdist) */
       SDYN_PROJ_KEYFRAME_SIM3 = 1 };  /* SearchByProjection(KeyFrame* pKF, cv::Mat Scw, vpPoints, vpMatched, th) */

/* Pose and thresholds of one pose-projection search.  rcw / tcw / ow are the values of the reference's own host
 * expressions (Rcw, tcw, Ow = -Rcw.t()*tcw; for the Scw overload after dividing by the scale, :300-304) — the
 * adapter evaluates them with cv::Mat exactly as the reference does and passes the floats. */
typedef struct {
    float rcw[9], tcw[3], ow[3];
    float th;                       /* search radius factor (the int th of the Scw overload converted to float) */
    int32_t max_descriptor_distance;/* ORBdist, or TH_LOW for the Scw overload */
    int32_t variant;                /* SDYN_PROJ_* */
    int32_t check_orientation;      /* mbCheckOrientation (Frame/KeyFrame overload only) */
    float log_scale_factor;         /* mfLogScaleFactor of the searched frame */
    int32_t nlevels;                /* mnScaleLevels of the searched frame */
} sdyn_proj_params;

/* DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>, Thirdparty/DBoW2/DBoW2/FeatureVector.h:22)
 * flattened to CSR with ascending node ids. */
typedef struct {
    int32_t nnodes;
    const uint32_t* node_id;
    const int32_t* offset;          /* nnodes + 1 */
    const uint32_t* index;
} sdyn_feature_vector;

/* ORBmatcher::DescriptorDistance (static, src/ORBmatcher.cc:1804-1820): 256-bit Hamming distance of one pair.
 * Host-side popcount — the static member is called pair-at-a-time by MapPoint.cc:281 and friends; the
 * searches below do their distance work on the device. */
int sdyn_hamming(const uint8_t* a, const uint8_t* b);

/* Hamming-distance evaluations (ORBmatcher::DescriptorDistance calls of the reference loop, src/ORBmatcher.cc:1804-1820) the
 * LAST search call on this context performed: the unit of the popc roofline. */
long long sdyn_match_last_evals(const sdyn_ctx* ctx);

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th),
 * src/ORBmatcher.cc:45-129.  *nmatches = return value. */
int sdyn_match_projection_map(sdyn_ctx* ctx, const sdyn_frame_view* frame, const sdyn_mappoint_query* mps,
                              int nmp, float th, float nnratio, int32_t* assign, uint8_t* locked, int* nmatches);

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
 * src/ORBmatcher.cc:1485-1627, and the fork's overload with two extra vector<cv::Point2f>& outputs
 * (:407-559) when pairs != NULL: pairs receives (last.x, last.y, cur.x, cur.y) per accepted match, in
 * query order, appended before the rotation cull; capacity 4 * last->n floats. */
int sdyn_match_projection_frame(sdyn_ctx* ctx, const sdyn_frame_view* cur, const sdyn_frame_view* last,
                                const sdyn_last_point* last_points, float th, int mono, int check_orientation,
                                int32_t* assign, uint8_t* locked, int* nmatches, float* pairs, int* npairs);

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist),
 * src/ORBmatcher.cc:1629-1756 (relocalisation, Tracking.cc:2323,2337) and
 * ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th),
 * src/ORBmatcher.cc:290-403 (loop closing, LoopClosing.cc:376).  target: the frame / keyframe whose keypoints are
 * searched (grid bounds, intrinsics, mvKeysUn, descriptors).  assign in: -1 = keypoint free, anything else =
 * occupied (mvpMapPoints[i] / vpMatched[i] != NULL); out: index into pts of the point that took the keypoint. */
int sdyn_match_projection_pose(sdyn_ctx* ctx, const sdyn_frame_view* target, const sdyn_proj_point* pts, int npts,
                               const sdyn_proj_params* params, int32_t* assign, int* nmatches);

/* Independent best-keypoint search behind ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:
 * 982-1130), ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (:1132-1257) and the two passes of
 * ORBmatcher::SearchBySim3 (:1259-1483): every candidate MapPoint is projected, gated (depth, image, distance range,
 * viewing angle, predicted level, reprojection chi-square) and given the keypoint of smallest descriptor distance in
 * its window.  Nothing here depends on the other points, so the pointer-graph side effects of the reference (Replace,
 * AddObservation, the mutual-agreement check) stay in the adapter and run over best_idx / best_dist.
 * t1 (and t2 when use_t2): row-major [R|t] the point goes through, values of the reference's host expressions. */
typedef struct {
    float t1[12], t2[12];
    int32_t use_t2;                 /* SearchBySim3: p3Dc2 = sR21*(R1w*p3Dw + t1w) + t21 */
    float ow[3];                    /* camera centre (distance / viewing-angle gates) */
    int32_t invz_double;            /* invz = 1.0/z in double (Fuse Scw, SearchBySim3) instead of float 1/z (Fuse) */
    int32_t dist_from_camera;       /* dist3D = norm(p3Dc) (SearchBySim3) instead of norm(p3Dw - Ow) */
    int32_t check_normal;           /* PO.dot(Pn) < 0.5*dist3D gate (both Fuse overloads) */
    int32_t chi2_gate;              /* Fuse: e2*mvInvLevelSigma2[level] > 5.99 (mono) / 7.8 (stereo, target->u_right >= 0) */
    float bf;                       /* mbf (stereo reprojection) */
    float inv_level_sigma2[SDYN_MAX_LEVELS];
    float th, log_scale_factor;
    int32_t nlevels;
} sdyn_best_params;
/* best_idx[i] = keypoint of target or -1, best_dist[i] = its descriptor distance (256 when none). */
int sdyn_match_projection_best(sdyn_ctx* ctx, const sdyn_frame_view* target, const sdyn_proj_point* pts, int npts,
                               const sdyn_best_params* params, int32_t* best_idx, int32_t* best_dist);

/* ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, vector<pair<size_t,size_t>> &vMatchedPairs,
 * const bool bOnlyStereo), src/ORBmatcher.cc:814-980 with CheckDistEpipolarLine (:140-157) — LocalMapping::CreateNewMapPoints.
 * has_mp1 / has_mp2: pKF->GetMapPoint(idx) != NULL; u_right of the two views = mvuRight (NULL = monocular).  f12: row-major
 * float F12; epipole: (ex, ey) of :823-829 as the reference's cv::Mat expressions give them; level_sigma2: pKF2->mvLevelSigma2.
 * matches12[i1] = KeyFrame-2 index or -1 (vMatchedPairs = the pairs (i1, matches12[i1]) in ascending i1). */
typedef struct {
    float f12[9];
    float epipole_x, epipole_y;
    int32_t only_stereo, check_orientation;
    float level_sigma2[SDYN_MAX_LEVELS];
} sdyn_tri_params;
int sdyn_match_triangulation(sdyn_ctx* ctx, const sdyn_frame_view* kf1, const uint8_t* has_mp1, const sdyn_feature_vector* fv1,
                             const sdyn_frame_view* kf2, const uint8_t* has_mp2, const sdyn_feature_vector* fv2,
                             const sdyn_tri_params* params, int32_t* matches12, int* nmatches);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), src/ORBmatcher.cc:562-677.
 * prev_matched: f1->n x 2 floats, updated in place; matches12: f1->n. */
int sdyn_match_init(sdyn_ctx* ctx, const sdyn_frame_view* f1, const sdyn_frame_view* f2, float* prev_matched,
                    int32_t* matches12, int window_size, float nnratio, int check_orientation, int* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, vector<MapPoint*> &vpMapPointMatches), src/ORBmatcher.cc:159-288.
 * kf_valid[i] = keyframe keypoint i has a non-bad MapPoint; assign[f_idx] = kf_idx | -1 (f->n entries). */
int sdyn_match_bow(sdyn_ctx* ctx, const sdyn_frame_view* kf, const uint8_t* kf_valid,
                   const sdyn_feature_vector* kf_fv, const sdyn_frame_view* f, const sdyn_feature_vector* f_fv,
                   float nnratio, int check_orientation, int32_t* assign, int* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12), src/ORBmatcher.cc:679-812
 * (LoopClosing.cc:266).  valid1[i] / valid2[i]: feature i of that keyframe has a MapPoint that is not bad.
 * matches12[i1] = index of the matched KeyFrame-2 feature (vpMatches12[i1] = vpMapPoints2[that index]) or -1. */
int sdyn_match_bow_kf(sdyn_ctx* ctx, const sdyn_frame_view* kf1, const uint8_t* valid1, const sdyn_feature_vector* fv1,
                      const sdyn_frame_view* kf2, const uint8_t* valid2, const sdyn_feature_vector* fv2, float nnratio,
                      int check_orientation, int32_t* matches12, int* nmatches);

/* ---- dynamic-keypoint rejection ---------------------------------------------------------------------
 * Frame::firstSeparate's keypoint-in-box test (src/Frame.cc:555-572): bit b of mask[i] is set iff
 * cv::Rect2d boxes[b] contains keys[i].pt (x <= px < x+w, y <= py < y+h, doubles).  nboxes <= 64.
 * The reorder / per-box bookkeeping that follows is host code in the Frame adapter. */
int sdyn_dyn_box_mask(sdyn_ctx* ctx, const sdyn_keypoint* keys, int n, const double* boxes_xywh, int nboxes,
                      uint64_t* mask);

/* One (current box, reference box) pair of Tracking::Separate (src/Tracking.cc:1093-1239). */
typedef struct {
    int32_t nq, nt;                 /* keypoints in the current / reference box */
    const uint8_t* q_desc;          /* mdynDescriptors[n_box], nq x 32 */
    const uint8_t* t_desc;          /* mRefFrame->mdynDescriptors[ref_idx], nt x 32 */
    const float* q_xy;              /* mvdynKeysUn[n_box][k].pt, nq x 2 */
    const float* t_xy;              /* nt x 2 */
    int32_t* match_query;           /* out, capacity nq: cv::DMatch::queryIdx of match m */
    int32_t* match_train;           /* out: trainIdx */
    int32_t* match_dist;            /* out: Hamming distance */
    int32_t* false_dyn;             /* out: classifyF/H result per match: queryIdx (static) or -1 */
    int32_t nmatches;               /* out */
} sdyn_box_pair;

/* cv::BFMatcher(NORM_HAMMING, crossCheck=true).match (Tracking.cc:1096,1122) followed by classifyF
 * (mode 0, F21, threshold 5.841, Tracking.cc:1311-1367) or classifyH (mode 1, H21 and its inverse,
 * threshold 5.991, :1241-1309) for every pair.  m3x3: row-major float.  The >=3 / 20 % gates and the box
 * status logic stay in the adapter. */
int sdyn_dyn_separate(sdyn_ctx* ctx, sdyn_box_pair* pairs, int npairs, const float* m3x3, int mode);

/* ---- batched front end: extract + track + dynamic mask, device-resident ---------------------------------
 * One call processes nframes independent frames (frame / sequence sharding is the caller's: there is no
 * cross-frame or cross-GPU exchange).  Per frame f, in this order on one stream:
 *   1. ORBextractor::operator()                                   (as sdyn_extract_batch_device)
 *   2. SearchByProjection(CurrentFrame, LastFrame, th, bMono)     src/ORBmatcher.cc:1485  (TrackWithMotionModel)
 *   3. SearchByProjection(CurrentFrame, vpMapPoints, th)          src/ORBmatcher.cc:45    (SearchLocalPoints),
 *      on the mvpMapPoints left by step 2
 *   4. dynamic mask: Frame::firstSeparate's box test, Tracking::Separate's per-box BFMatcher + classifyF and
 *      Frame::UpdateFrame's re-admission:  mask[i] = in_box(i) && !readmitted(i)
 * Searches see every extracted keypoint (the stereo constructor's behaviour, where firstSeparate is disabled,
 * src/Frame.cc:166) unless rgbd_split is set (below).  All pointers are DEVICE pointers; per-frame arrays are `*_stride`
 * elements apart.
 *
 * What persists across frames in the reference stays on the device here.  In the reference the query data of the two
 * searches lives in MapPoint objects (world position, normal, descriptor, scale-invariance distances) and in LastFrame
 * (its keypoints); only the per-frame bookkeeping changes.  Two resident forms mirror that:
 *   - `map`: a device-resident MapPoint table (sdyn_map_create / sdyn_map_update), addressed by point id;
 *   - resident LastFrame: the keypoints each batch slot extracted in the PREVIOUS step of this context (slot f of step k is
 *     the LastFrame of slot f of step k+1: one sequence per slot).
 * With them a step uploads, per frame, the image, the ids / flag bytes / projection records and the pose. */
typedef struct {
    float world[3];                 /* GetWorldPos() */
    float normal[3];                /* GetNormal() */
    float min_distance, max_distance;   /* mfMinDistance, mfMaxDistance (the 0.8 / 1.2 factors are applied on the device) */
    uint8_t desc[32];               /* GetDescriptor() */
} sdyn_map_point;                   /* 64 bytes */
typedef struct sdyn_map sdyn_map;
int sdyn_map_create(int device, int capacity, sdyn_map** out);
int sdyn_map_destroy(sdyn_map* map);
/* points[0..count) become ids first_id .. first_id+count-1.  HOST array; the copy is ordered on `stream` (NULL = default). */
int sdyn_map_update(sdyn_map* map, int first_id, int count, const sdyn_map_point* points, void* stream);
int sdyn_map_capacity(const sdyn_map* map);

/* per LastFrame keypoint in the resident form: bit flags next to the MapPoint id (-1 = mvpMapPoints[i] == NULL) */
enum { SDYN_LP_OUTLIER = 1, SDYN_LP_OBS_POSITIVE = 2 };
enum { SDYN_MP_BAD = 1, SDYN_MP_OBS_POSITIVE = 2, SDYN_MP_SKIP = 4 };
/* per local-map point in the resident form: what Frame::isInFrustum left in the MapPoint for this frame
 * (src/Frame.cc:718-727) plus the two per-frame state bits; = sdyn_mappoint_query without the descriptor */
typedef struct {
    float proj_x, proj_y, proj_xr, view_cos;
    int32_t level;
    uint8_t track_in_view, bad, obs_positive, pad;
} sdyn_map_proj;                    /* 24 bytes */

typedef struct {
    /* LastFrame of every current frame, explicit form.  last_keys_un may be NULL or equal to last_keys (mvKeysUn == mvKeys
     * for an undistorted camera, src/Frame.cc:814-818): the host-buffer entry points then upload the array once. */
    const sdyn_last_point* last_points; const sdyn_keypoint* last_keys; const sdyn_keypoint* last_keys_un;
    const int32_t* n_last; int32_t last_stride;
    /* local map of every current frame, explicit form */
    const sdyn_mappoint_query* map_points; const int32_t* n_map; int32_t map_stride;
    /* detection boxes (cv::Rect2d, 64 x 4 doubles per frame) and, per box, the reference frame's box joined by
     * Frame::boxTrack (ref_box[f*64+b] = index into that frame's reference boxes or -1) */
    const double* boxes; const int32_t* n_boxes; const int32_t* ref_box;
    /* reference frame's per-box dynamic keypoints (mdynDescriptors / mvdynKeysUn): box r of frame f occupies
     * [ref_off[f*65+r], ref_off[f*65+r+1]) of that frame's ref_desc / ref_xy block */
    const uint8_t* ref_desc; const float* ref_xy; const int32_t* ref_off; int32_t ref_stride;
    const float* fmat;              /* F21 per frame, 9 floats (classifyF) */
    /* camera and search parameters shared by the batch */
    float min_x, min_y, max_x, max_y, fx, fy, cx, cy, bf, b;
    float th_frame, th_map, nnratio_map;
    int32_t mono, check_orientation;
    /* poses: per frame rows 0..2 of CurrentFrame.mTcw then of LastFrame.mTcw (24 floats).  The forward / backward level
     * rule of ORBmatcher.cc:1497-1506 is evaluated per frame on the device. */
    const float* poses;
    /* resident forms (see above); each replaces the explicit array of the same search when that array is NULL:
     *   last_points == NULL: last_ids[f*last_stride + i] / last_flags[..] describe keypoint i of slot f's previous frame,
     *                        whose keypoints (and count) are taken from the previous step; last_keys / n_last are unused;
     *   map_points == NULL:  map_ids / map_proj [f*map_stride + q], n_map as before. */
    const sdyn_map* map;
    const int32_t* last_ids; const uint8_t* last_flags;
    const int32_t* map_ids; const sdyn_map_proj* map_proj;
    /* map_points == NULL and map_proj == NULL: Frame::isInFrustum (src/Frame.cc:677-733, viewing-cosine limit below; 0.5 in
     * Tracking::SearchLocalPoints) runs on the device for every local-map id from the table entry and the frame's pose, and
     * only one state byte per point travels: SDYN_MP_BAD | SDYN_MP_OBS_POSITIVE | SDYN_MP_SKIP (SKIP = the point is not to be
     * searched in this frame: mnLastFrameSeen == mCurrentFrame.mnId in SearchLocalPoints). */
    const uint8_t* map_flags; float viewing_cos_limit;
    /* 0: stereo-constructor semantics (above).  1: RGB-D-constructor semantics (src/Frame.cc:297-403 followed by
     * Tracking::Separate and Frame::UpdateFrame :607-653): the searches run on the frame the reference tracks with — the
     * keypoints outside every box in extraction order, then the re-admitted in-box keypoints in UpdateFrame's push order —
     * and assign / locked / the resident LastFrame of the next step index THAT list (sdyn_track_frame_order gives the
     * extraction index of each entry). */
    int32_t rgbd_split;
    /* 0: array-major — array X of frame f starts X_stride elements after frame f-1's (every array contiguous over the batch).
     * > 0: frame-major records — array X of frame f starts at (char*)X + f * frame_pitch: each frame's inputs form one
     * contiguous record of frame_pitch bytes (sdyn_track_record_layout), so any run of consecutive frames of a host-side
     * pool uploads with ONE copy. */
    int64_t frame_pitch;
} sdyn_track_inputs;

typedef struct {
    const int32_t* assign;          /* [max_batch][cap]: mvpMapPoints after steps 2+3: -1, last index, or n_last + map index */
    const uint8_t* locked;          /* [max_batch][cap] */
    const uint8_t* dyn_mask;        /* [max_batch][cap] */
    const int32_t* counts;          /* [max_batch][4]: nmatches of step 2, nmatches of step 3, #in-box, #masked */
} sdyn_track_view;

int sdyn_track_batch_device(sdyn_ctx* ctx, int nframes, const uint8_t* d_gray, size_t frame_stride,
                            int width, int height, int stride, const sdyn_track_inputs* in, void* stream);
/* Stereo form of the same step (the stereo Frame constructor, src/Frame.cc:140-175, then the searches): both views are
 * extracted side by side on the two contexts' streams, Frame::ComputeStereoMatches runs on the device, and the two
 * searches apply their mvuRight gates (ORBmatcher.cc:94-99, 1567-1573) to the result.  Inputs are device pointers;
 * in->mono must be 0 for the forward / backward level rule; results as for sdyn_track_batch_device plus
 * sdyn_stereo_results on `ctx` (the left context). */
int sdyn_track_batch_stereo_device(sdyn_ctx* ctx, sdyn_ctx* right, int nframes, const uint8_t* d_gray_left,
                                   const uint8_t* d_gray_right, size_t frame_stride, int width, int height, int stride,
                                   const sdyn_track_inputs* in, float mb, float mbf);
/* Host-buffer form of the same step (what a caller holding frames and map data in host memory uses): every
 * pointer of `in` and `gray` is HOST memory (pinned memory copies asynchronously), results are written to the
 * host arrays (kp_out/desc_out: [nframes][cap], n_out: [nframes]; assign/locked/dyn_mask: [nframes][cap];
 * counts: [nframes][4]; any output may be NULL).  Host->device copies of the frames and of all query arrays
 * and the device->host copies of the results happen inside the call. */
int sdyn_track_batch(sdyn_ctx* ctx, int nframes, const uint8_t* gray, size_t frame_stride, int width, int height,
                     int stride, const sdyn_track_inputs* in, sdyn_keypoint* kp_out, uint8_t* desc_out, int* n_out,
                     int32_t* assign, uint8_t* locked, uint8_t* dyn_mask, int32_t* counts, int cap);
/* Asynchronous form: returns once the copies and kernels are enqueued on the context's stream; the output
 * arrays are valid after sdyn_track_wait().  Two contexts used alternately overlap one step's PCIe transfers
 * with the other's kernels (pin the host buffers with sdyn_host_alloc for truly asynchronous copies). */
int sdyn_track_batch_async(sdyn_ctx* ctx, int nframes, const uint8_t* gray, size_t frame_stride, int width, int height,
                           int stride, const sdyn_track_inputs* in, sdyn_keypoint* kp_out, uint8_t* desc_out, int* n_out,
                           int32_t* assign, uint8_t* locked, uint8_t* dyn_mask, int32_t* counts, int cap);
/* sdyn_track_batch / sdyn_track_wait / sdyn_track_fetch return SDYN_ERR_CAPACITY when a search of the step ran out of candidate
 * pool (a very dense frame or a very wide window): that step's matches are incomplete, the pool has been doubled for the next
 * call, and NO slot's resident LastFrame was advanced by the failed step — submit the same step again. */
int sdyn_track_wait(sdyn_ctx* ctx);
/* Layout of the host-buffer entry points' input staging block for `nframes` frames: offsets[i] is where array i of
 * sdyn_track_inputs starts (order: last_points, last_keys, last_keys_un, n_last, map_points, n_map, boxes, n_boxes, ref_box,
 * ref_desc, ref_xy, ref_off, fmat, poses, last_ids, last_flags, map_ids, map_proj, map_flags; each on a 256-byte boundary; an array takes
 * no room when its form is not in use: `forms` bit 0 = separate last_keys_un, bit 1 = resident LastFrame (last_ids /
 * last_flags instead of last_points / last_keys), bit 2 = resident map (map_ids / map_proj instead of map_points)), *total
 * the block size.  A caller that builds its arrays inside ONE host block at these offsets (pinned with sdyn_host_alloc) and
 * passes base + offsets[i] as the array pointers has them uploaded with a single copy instead of one copy per array — the
 * per-frame query data the reference keeps in MapPoint / Frame objects (src/ORBmatcher.cc:45-129, 1485-1627) is gathered
 * by the caller either way. */
#define SDYN_TRACK_INPUT_ARRAYS 19
enum { SDYN_FORM_SEPARATE_KEYS_UN = 1, SDYN_FORM_RESIDENT_LAST = 2, SDYN_FORM_RESIDENT_MAP = 4,
       SDYN_FORM_DEVICE_FRUSTUM = 8 /* with RESIDENT_MAP: map_flags instead of map_proj */ };
int sdyn_track_input_layout(int nframes, int last_stride, int map_stride, int ref_stride, int forms,
                            size_t offsets[SDYN_TRACK_INPUT_ARRAYS], size_t* total);
/* Frame-major counterpart: offsets[i] = where array i starts inside ONE frame's record, *pitch = the record size (a multiple
 * of 256).  A host pool of records (pinned) hands any run of consecutive frames to the host-buffer entry points as one copy:
 * pass array_i = pool + first_frame * pitch + offsets[i] and frame_pitch = pitch. */
int sdyn_track_record_layout(int last_stride, int map_stride, int ref_stride, int forms,
                             size_t offsets[SDYN_TRACK_INPUT_ARRAYS], size_t* pitch);
/* Statistics of the last fetched step: Hamming-distance evaluations of the frame search and of the map search,
 * summed over the step's frames (the work the popc roofline is computed from). */
int sdyn_track_stats(const sdyn_ctx* ctx, int nframes, long long evals[2]);
int sdyn_track_results(const sdyn_ctx* ctx, sdyn_track_view* out);
/* rgbd_split steps: order[f*cap + j] = extraction index of entry j of frame f's tracked keypoint list, n[f] its length
 * (N after UpdateFrame), n_static[f] = N_s.  HOST arrays ([nframes][cap], [nframes], [nframes]); synchronises the stream. */
int sdyn_track_frame_order(sdyn_ctx* ctx, int nframes, int32_t* order, int32_t* n, int32_t* n_static, int cap);
/* D2H of the step results next to sdyn_fetch_results (host arrays sized [nframes][cap] / [nframes][4]). */
int sdyn_track_fetch(sdyn_ctx* ctx, int nframes, int32_t* assign, uint8_t* locked, uint8_t* dyn_mask,
                     int32_t* counts, int cap, void* stream);

/* ---- detection boxes at the edge of the dynamic-keypoint path (host code) ---------------------------------------
 * The detector's per-frame text file: one "id cx cy w h" line per detection (Examples/RGB-D/rgbd_my.cc:232-252);
 * each becomes cv::Rect2d(MAX(cx - w/2, 0), MAX(cy - h/2, 0), w, h), written as 4 doubles.  Empty lines are skipped,
 * as in the reference; lines with fewer than 5 numbers are skipped too (the reference reads uninitialised values).
 * *n_out = detections found; SDYN_ERR_CAPACITY if more than cap.  A missing file is a frame without boxes. */
int sdyn_boxes_parse(const char* text, size_t len, double* xywh, int cap, int* n_out);
int sdyn_boxes_read(const char* path, double* xywh, int cap, int* n_out);
/* Frame::boxTrack (src/Frame.cc:481-552): joins this frame's boxes with the last frame's objects by IoU, carries
 * unmatched objects over once and numbers new ones.  boxes: in/out, nboxes valid of capacity cap (carried-over boxes
 * are appended); last_*: objects, box_idx, omit, box_velocity of the last frame (nlast entries; nlast == 0 = the
 * re-initialised case).  Outputs box_idx / omit / velocity (2 doubles per box) for *n_out boxes. */
int sdyn_box_track(double* boxes, int nboxes, int cap, const double* last_objects, const int32_t* last_box_idx,
                   const uint8_t* last_omit, const double* last_velocity, int nlast, int img_width, int img_height,
                   int32_t* box_idx, uint8_t* omit, double* velocity, int* n_out);

/* ---- Frame::UndistortKeyPoints / ComputeImageBounds -----------------------------------------------------------
 * Camera of the context: mK = (fx, fy, cx, cy) and mDistCoef = (k1, k2, p1, p2[, k3]) as floats (ncoef = 4 or 5; 0
 * = no distortion).  With k1 != 0 (the reference's test, src/Frame.cc:814,846) every extraction also produces
 * mvKeysUn on the device — cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) per keypoint — and the
 * batched front end searches the undistorted keypoints; otherwise mvKeysUn == mvKeys. */
int sdyn_set_camera(sdyn_ctx* ctx, float fx, float fy, float cx, float cy, const float* dist_coef, int ncoef);
/* mvKeysUn of the last extraction: [nframes][cap] (equals the keypoints when the camera has no distortion). */
int sdyn_fetch_keypoints_un(sdyn_ctx* ctx, int nframes, sdyn_keypoint* kp_out, int cap, void* stream);
/* n (x, y) float pairs through the same device routine (host arrays). */
int sdyn_undistort_points(sdyn_ctx* ctx, const float* xy, int n, float* out_xy);
/* Frame::ComputeImageBounds (src/Frame.cc:844-872): bounds = {mnMinX, mnMinY, mnMaxX, mnMaxY}. */
int sdyn_image_bounds(sdyn_ctx* ctx, int width, int height, float bounds[4]);

/* ---- Frame::ComputeStereoMatches ------------------------------------------------------------------------
 * Replaces Frame::ComputeStereoMatches (src/Frame.cc:874-1048), the step that follows the two extractions in the
 * stereo Frame constructor (src/Frame.cc:151-160).  `left` and `right` are the contexts of mpORBextractorLeft /
 * mpORBextractorRight; the call works on the keypoints, descriptors and pyramids their LAST extraction of
 * `nframes` frames left on the device (same image size and ORB parameters on both), so no pyramid level travels
 * to the host.  mb / mbf: Frame::mb, Frame::mbf.  Outputs per left keypoint: mvuRight, mvDepth (-1 = none).
 * The device form only enqueues on `stream` (NULL = left's stream) after ordering it behind both extractions;
 * results stay device-resident (sdyn_stereo_results) and can feed sdyn_frame_view.u_right of later searches. */
typedef struct {
    const float* u_right;           /* [max_batch][cap]  mvuRight */
    const float* depth;             /* [max_batch][cap]  mvDepth */
    const int32_t* kept;            /* [max_batch] stereo points that survive the median cut (:1034-1047) */
    int32_t cap;
} sdyn_stereo_view;
int sdyn_stereo_match_device(sdyn_ctx* left, sdyn_ctx* right, int nframes, float mb, float mbf, void* stream);
int sdyn_stereo_results(const sdyn_ctx* left, sdyn_stereo_view* out);
/* D2H of mvuRight / mvDepth: host arrays [nframes][cap]; kept (nullable) [nframes].  Synchronises `stream`.
 * SDYN_ERR_GEOMETRY if a right keypoint's row band left the image (the reference writes out of bounds there). */
int sdyn_stereo_fetch(sdyn_ctx* left, int nframes, float* u_right, float* depth, int cap, int32_t* kept, void* stream);
/* Host-output form: device step + fetch. */
int sdyn_stereo_match(sdyn_ctx* left, sdyn_ctx* right, int nframes, float mb, float mbf, float* u_right, float* depth,
                      int cap, int32_t* kept);

/* ---- Frame::ComputeBoW -------------------------------------------------------------------------------------------
 * Replaces mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) (src/Frame.cc:803-810; DBoW2
 * TemplatedVocabulary.h:1127-1263), the step in front of SearchByBoW.  The vocabulary tree is uploaded once.
 * Nodes are given in DBoW2 node-id order, node 0 = root: parent[i] (< i), is_leaf[i], descriptor (32 bytes), weight —
 * exactly the columns of ORBvoc.txt (loadFromTextFile, :1338-1425); children keep file order and word ids number the
 * leaves in file order.  Scoring L1_NORM + weighting TF_IDF (the ORBvoc header "10 6 0 0") are what is implemented. */
typedef struct sdyn_vocab sdyn_vocab;
int sdyn_vocab_create(int device, int nnodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                      const double* weight, int k, int L, sdyn_vocab** out);
int sdyn_vocab_load_text(int device, const char* path, sdyn_vocab** out);
int sdyn_vocab_destroy(sdyn_vocab* voc);
int sdyn_vocab_info(const sdyn_vocab* voc, int32_t info[4]);      /* nodes, words, k, L */
/* Tree descent of every descriptor (TemplatedVocabulary::transform(feature, id, weight, nid, levelsup), :1210-1258):
 * per feature the word id, the word's weight and the node id levelsup levels above the leaves.  Host-array form: */
int sdyn_bow_transform(sdyn_ctx* ctx, const sdyn_vocab* voc, const uint8_t* desc, int n, int levelsup,
                       uint32_t* word_id, double* weight, uint32_t* node_id);
/* Device-resident form on the descriptors of the context's last extraction (results: sdyn_bow_fetch, [nframes][cap]). */
int sdyn_bow_transform_device(sdyn_ctx* ctx, const sdyn_vocab* voc, int nframes, int levelsup, void* stream);
int sdyn_bow_fetch(sdyn_ctx* ctx, int nframes, uint32_t* word_id, double* weight, uint32_t* node_id, int cap, void* stream);
/* Host assembly of the two containers from those arrays, as transform(features, BowVector&, FeatureVector&, levelsup)
 * does it: BowVector = ascending word ids with the TF-IDF weights summed in feature order and L1-normalised
 * (bow_ids / bow_values, capacity n); FeatureVector = CSR over ascending node ids (fv_nodes capacity n, fv_offset n+1,
 * fv_index n).  Host code — the containers are std::map in the drop-in. */
int sdyn_bow_assemble(const uint32_t* word_id, const double* weight, const uint32_t* node_id, int n, uint32_t* bow_ids,
                      double* bow_values, int* n_words, uint32_t* fv_nodes, int32_t* fv_offset, uint32_t* fv_index,
                      int* n_fv_nodes);

/* ---- per-stage device timing (CUDA events on the launching stream) -------------------------------
 * While enabled, every enqueue brackets each stage with events; sdyn_profile_read synchronises and
 * returns the accumulated milliseconds and launch counts since the last read. */
enum { SDYN_STAGE_PYRAMID = 0, SDYN_STAGE_FAST, SDYN_STAGE_OCTREE, SDYN_STAGE_BLUR, SDYN_STAGE_DESCRIBE,
       SDYN_STAGE_MATCH, SDYN_STAGE_DYNAMIC, SDYN_STAGE_LEVEL0 /* clears + level 0; PYRAMID = the resize chain */,
       SDYN_STAGE_STEREO, SDYN_STAGE_BOW,
       SDYN_STAGE_CANDIDATES /* k_match_candidates alone; MATCH = the whole search stage around it */,
       SDYN_STAGE_COUNT };
typedef struct {
    double ms[SDYN_STAGE_COUNT];
    long long calls[SDYN_STAGE_COUNT];
} sdyn_stage_times;
int sdyn_profile_enable(sdyn_ctx* ctx, int on);
int sdyn_profile_read(sdyn_ctx* ctx, sdyn_stage_times* out);

#ifdef __cplusplus
}
#endif
#endif /* SDYN_H_ */
