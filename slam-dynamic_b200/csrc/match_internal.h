/* Device-side job descriptors of the matcher kernels (k_match.cu) — shared with the host glue. */
#pragma once
#include "sdyn_internal.h"

namespace sdyn {

constexpr int kGridCells = SDYN_GRID_COLS * SDYN_GRID_ROWS;   /* 3072 */

enum MatchMode { MM_FRAME = 0, MM_MAP = 1, MM_INIT = 2, MM_BOW = 3, MM_POSE = 4, MM_BEST = 5, MM_TRI = 6 };

/* query of the BoW search: one keyframe feature against the frame features of the same vocabulary node */
struct BowQuery { int32_t kfIdx, fOff, fCnt; };

/* One search call.  All pointers are device pointers.  Several jobs run per launch (blockIdx.y). */
struct MatchJob {
    int mode;
    /* frame that is searched (CurrentFrame / F / F2 / F) */
    const sdyn_keypoint* keysUn;
    const uint8_t* desc;
    const float* uRight;             /* may be null */
    const int32_t* nPtr;             /* device-resident keypoint count, or null -> n */
    int n;
    float minX, minY, maxX, maxY, gridWInv, gridHInv;
    float scale[SDYN_MAX_LEVELS];
    /* grid CSR built by k_grid_build: keypoints sorted by (cell = ix*48+iy, index) */
    int32_t* cellOff;                /* kGridCells + 1 */
    int32_t* sorted;                 /* n */
    int32_t* cellOf;                 /* n, scratch */
    float4* gridEntry;               /* n: (x, y, bits(index | octave << 24), 0) of sorted[p] — one load per candidate */
    /* queries */
    const void* queries;             /* sdyn_mappoint_query / sdyn_last_point / F1 keypoints / BowQuery */
    const sdyn_keypoint* qKeys;      /* FRAME: LastFrame.mvKeys (octave), INIT: F1.mvKeysUn, BOW: KF.mvKeysUn */
    const sdyn_keypoint* qKeysUn;    /* FRAME: LastFrame.mvKeysUn (angle, pt for pairs) */
    const uint8_t* qDesc;            /* INIT: F1 descriptors, BOW: KF descriptors */
    const uint32_t* fIndex;          /* BOW: F feature-vector index array */
    const int32_t* nqPtr;            /* device-resident query count, or null -> nq */
    int nq;
    float* prevMatched;              /* INIT: in/out, nq x 2 */
    /* parameters */
    float th, nnratio;
    int checkOri, forward, backward, window;
    int distTh;                      /* FRAME / MAP / POSE: accept best <= distTh (TH_HIGH, or ORBdist / TH_LOW of the pose searches) */
    int poseVariant;                 /* POSE: SDYN_PROJ_FRAME_KEYFRAME or SDYN_PROJ_KEYFRAME_SIM3 */
    float Ow[3], logScaleFactor;     /* POSE: camera centre, Frame::mfLogScaleFactor */
    int predLevels;                  /* POSE: mnScaleLevels (PredictScale clamp) */
    /* BEST (Fuse x2, SearchBySim3 passes): independent best keypoint per point, no claims */
    float T2[12]; int useT2, invzDouble, distFromCamera, checkNormal, chi2Gate;
    float invSigma2[SDYN_MAX_LEVELS];
    /* TRI (SearchForTriangulation): F12 row-major, epipole, per-feature flags */
    float F12[9], epiX, epiY, sigma2[SDYN_MAX_LEVELS];
    const uint8_t* fValid;           /* searched feature may be matched (no MapPoint, stereo filter) */
    const float* qURight;            /* mvuRight of the query keyframe, or null */
    int strictLow;                   /* BOW: accept best < TH_LOW (KeyFrame-KeyFrame overload) instead of <= (KeyFrame-Frame) */
    int assignBase;                  /* value written to assign[] = assignBase + query index (FRAME / MAP) */
    float Tcw[12], fx, fy, cx, cy, bf;
    /* FRAME, batched front end: per-frame pose table entry (24 floats: rows 0..2 of CurrentFrame.mTcw, then of LastFrame.mTcw).
     * When set, Tcw / forward / backward above are derived from it on the device (ORBmatcher.cc:1495-1506). */
    const float* pose; float mb; int mono;
    /* state / outputs */
    int32_t* assign;                 /* FRAME/MAP/BOW: per searched keypoint; INIT: matches12 per query */
    uint8_t* locked;                 /* FRAME/MAP */
    int32_t* matchedDist;            /* INIT: per F2 keypoint */
    int32_t* m21;                    /* INIT */
    int32_t* result;                 /* [0] nmatches, [1] npairs, [2] status (1 = pool overflow) */
    float* pairs;                    /* FRAME (fork overload), may be null */
    /* scratch */
    int32_t* qperm;                  /* visiting order of k_match_candidates (k_query_order), or null = identity */
    int2* qspan;                     /* per query: (offset into pool, count) */
    int32_t* qAccepted;              /* per query: claimed index or -1 */
    int32_t* qBin;                   /* per query: rotation-histogram bin */
    uint32_t* pool;                  /* candidate records: idx:16 | dist:9 | level:5 */
    int poolCap;
    int32_t* poolUsed;
    int32_t* qNext;                  /* next unvisited query slot (k_match_candidates work counter), zero on entry */
};

/* maxN: largest MatchJob::n of the launch (sizes the shared-memory sort buffer) */
/* orderJobs / nOrder: MM_MAP jobs whose query order (k_query_order) is produced by the same launch */
cudaError_t launch_grid_build(const MatchJob* dJobs, int njobs, int maxN, cudaStream_t st, const MatchJob* orderJobs = nullptr, int nOrder = 0);
/* fills MatchJob::qperm of MM_MAP jobs that have one */
cudaError_t launch_query_order(const MatchJob* dJobs, int njobs, cudaStream_t st);
/* maxN: largest MatchJob::n of the launch (sizes the shared-memory staging of the searched frame) */
cudaError_t launch_match_candidates(const MatchJob* dJobs, int njobs, int maxQueries, int maxN, cudaStream_t st);
/* mode: all jobs of one launch share a mode; maxN / maxQ: largest MatchJob::n / ::nq of the launch */
/* second > 0 (FRAME / MAP / POSE): job b + second is resolved by the same CTA right after job b */
cudaError_t launch_match_resolve(const MatchJob* dJobs, int njobs, int mode, int maxN, int maxQ, cudaStream_t st, int second = 0);

/* per-frame array of the batched front end: frame f's part of array `base` (array-major: `bytes` apart; frame-major records:
 * `pitch` apart — sdyn_track_inputs::frame_pitch) */
template <class T>
__host__ __device__ __forceinline__ const T* frame_part(const T* base, int f, size_t bytes, long long pitch)
{
    return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + (size_t)f * (pitch > 0 ? (size_t)pitch : bytes));
}

/* dynamic-keypoint kernels (k_dynamic.cu) */
struct BoxPairJob {
    int nq, nt;
    const uint8_t* qDesc; const uint8_t* tDesc;
    const float* qXY; const float* tXY;
    int32_t* nnQ; int32_t* dQ; int32_t* nnT;          /* scratch: nq, nq, nt */
    int32_t* outQuery; int32_t* outTrain; int32_t* outDist; int32_t* outFalseDyn; int32_t* outCount;
};
/* boxPitch / nbPitch: bytes between two jobs' box arrays / box counts (0 = boxStride * 32 / 4) */
cudaError_t launch_box_mask(const sdyn_keypoint* dKeys, const int32_t* nPtr, int n, int keyStride,
                            const double* dBoxes, const int32_t* nBoxesPtr, int nboxes, int boxStride,
                            uint64_t* dMask, int njobs, cudaStream_t st, size_t boxPitch = 0, size_t nbPitch = 0);
/* kp = mvKeys (box containment, Frame.cc:562), kpUn = mvKeysUn (classifyF coordinates, Tracking.cc:1129-1131) */
cudaError_t launch_dyn_stage(const sdyn_track_inputs& in, const sdyn_keypoint* kp, const sdyn_keypoint* kpUn, const uint8_t* desc, const int32_t* count,
                             int cap, uint64_t* mask, unsigned long long* has, int8_t* slotMap, int32_t* boxList, int32_t* nnQ, int32_t* nnT,
                             int nnTStride, int32_t* readmit, int32_t* staticExit, uint8_t* dynMask, int32_t* counts,
                             int nframes, cudaStream_t st);
/* RGB-D-constructor form of the tracked frame (k_track.cu) */
cudaError_t launch_frame_compact(const sdyn_keypoint* kp, const sdyn_keypoint* kpUn, const uint8_t* desc, const int32_t* count, int cap,
                                 const uint64_t* mask, const int32_t* readmit, const int32_t* staticExit, sdyn_keypoint* fKp,
                                 sdyn_keypoint* fKpUn, uint8_t* fDesc, int32_t* fOrder, int32_t* fCount, int32_t* fStatic, int nframes,
                                 cudaStream_t st);
/* camera / pyramid constants of Frame::isInFrustum on the device */
struct FrustumParams { float fx, fy, cx, cy, bf, minX, minY, maxX, maxY, cosLimit, logScaleFactor; int nlevels; };
/* resident query forms: ids + MapPoint table -> the searches' query records (k_track.cu) */
cudaError_t launch_gather_queries(const sdyn_map_point* table, int tableCap, const int32_t* lastIds, const uint8_t* lastFlags,
                                  const int32_t* nLast, int lastStride, sdyn_last_point* gLast, const int32_t* mapIds,
                                  const sdyn_map_proj* mapProj, const int32_t* nMap, int mapStride, sdyn_mappoint_query* gMap,
                                  const uint8_t* mapFlags, const float* poses, const FrustumParams& fp,
                                  int nframes, long long framePitch, cudaStream_t st);
cudaError_t launch_box_pairs(const BoxPairJob* dJobs, int njobs, const float* dM, const float* dMinv, int mode,
                             cudaStream_t st);

}  // namespace sdyn
