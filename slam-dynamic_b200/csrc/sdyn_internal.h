/* Internal declarations shared by the host side of libsdyn and its CUDA kernels. */
#pragma once
#include "../../include/sdyn.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

namespace sdyn {

constexpr int kEdge = SDYN_EDGE;
constexpr int kFastBorder = kEdge - 3;     /* minBorderX/Y = 16, src/ORBextractor.cc:773-774 */
constexpr int kRowAlign = 32;              /* interior column 0 of every level row is 32-byte aligned */
constexpr int kLeftPad = 32;               /* bytes before interior column 0 (19 of them are the border) */

/* Geometry of one level, valid for one input size.  Passed to kernels by value inside Geom. */
struct LevelGeom {
    int w, h;                /* level size  (ComputePyramid, src/ORBextractor.cc:1111-1112) */
    int pitch;               /* row pitch in bytes */
    int pad0;
    long long off;           /* byte offset, inside one frame's pyramid block, of interior pixel (0,0) */
    /* FAST window (ComputeKeyPointsOctTree, src/ORBextractor.cc:773-787) */
    int fw, fh;              /* maxBorder - minBorder */
    int nCols, nRows, wCell, hCell;
    int cellOff;             /* first cell flag of this level in the per-frame flag array */
    int candOff, candCap;    /* slice of the per-frame candidate array */
    /* octree (DistributeOctTree, :539-763) */
    int quota, nIni, nodeCap;
    float hX;
    int kpOff;               /* slice of the per-frame level-keypoint scratch (nodeCap entries) */
    /* keypoint post-processing (:835-846, :1095-1101) */
    float scale;
    float patchSize;
    /* resize tables (level > 0): byte offsets into the table blob */
    int xtab, ytab;
    /* shared-memory staging of the resize kernel: worst-case source rectangle of a 128 x 32 output tile */
    int rsPitch, rsRows;
    /* int16 tables (byte offsets into the blob): first staged source column of every tile column; (first staged source row,
     * row count) of every tile row */
    int xTile, yTile;
};

struct Geom {
    int nlevels, W, H, pad;
    long long frameBytes;    /* pyramid block per frame */
    int cellsPerFrame, candPerFrame, kpPerFrame, maxNodeCap;
    LevelGeom L[SDYN_MAX_LEVELS];
};

/* One column / row of the bilinear resize (cv::resize INTER_LINEAR 8-bit path, SURVEY A-2),
 * indexed by BORDERED destination coordinate so the REFLECT_101 frame is produced in the same pass. */
struct ResizeTap {
    int16_t s0, s1;          /* the two source indices */
    int16_t c0, c1;          /* 11-bit fixed-point coefficients */
};

struct TileRef { int16_t level, tx, ty, pad; };

struct Candidate {           /* 4 bytes: x:12 | y:12 | score:8, window-relative */
    uint32_t v;
};

/* Pinhole + Brown distortion as Frame holds it (mK, mDistCoef: float values, used in double by OpenCV) */
struct CameraModel {
    float fx, fy, cx, cy;
    float k[5];              /* k1 k2 p1 p2 k3 (k3 = 0 for 4-coefficient cameras) */
    int enabled;             /* mDistCoef.at<float>(0) != 0 (src/Frame.cc:814) */
};

struct LevelKp {             /* output of the octree stage, per (frame, level, slot) */
    int16_t x, y;            /* level coordinates (border added) */
    int32_t score;
};

}  // namespace sdyn

struct sdyn_ctx {
    sdyn_orb_params params;
    sdyn_scale_info scales;
    int umax[16];
    int maxW, maxH, maxBatch, device;
    int maxKp;                         /* per-frame output capacity */
    cudaStream_t stream;
    cudaStream_t aux;                          /* second stream: stages without mutual dependence run beside the main one */
    cudaEvent_t evFork, evJoin, evFork2, evJoin2;
    std::string err;
    long long launches;
    /* latency mode: the one-frame extraction (pinned staging -> kernels -> pinned results) as ONE CUDA graph */
    cudaGraphExec_t graph1; int graphW, graphH, graphCam; uint8_t* hIn; size_t hInCap; int latencyMode;
    long long lastEvals;                       /* Hamming evaluations of the last single-search call (sdyn_match_last_evals) */

    /* geometry for the current image size */
    sdyn::Geom geom;
    bool geomValid;
    std::vector<sdyn::TileRef> fastTiles, blurTiles;
    sdyn::TileRef* dFastTiles; sdyn::TileRef* dBlurTiles;
    int nFastTiles, nBlurTiles;
    uint8_t* dTables; size_t tablesCap;

    /* device buffers sized at create() for (maxW, maxH, maxBatch) */
    uint8_t* dIn;  size_t inFrameCap;          /* staged input frames (host entry points) */
    uint8_t* dPyr; uint8_t* dBlur; size_t pyrFrameCap;
    void* tma;                                 /* host TmaMaps (tma.h): per-level tensor maps of the blur tiles and descriptor patches */
    uint8_t* dCellFlag; int cellCap;           /* per frame */
    uint32_t* dCand; int32_t* dCandNode; int candCap;   /* per frame */
    int32_t* dCandCount;                       /* [maxBatch][SDYN_MAX_LEVELS] maxima emitted by FAST */
    int32_t* dSelCount;                        /* [maxBatch][SDYN_MAX_LEVELS] survivors of the per-cell threshold rule */
    sdyn::LevelKp* dLevelKp; int levelKpCap;   /* per frame */
    int32_t* dLevelCount;                      /* [maxBatch][SDYN_MAX_LEVELS] */
    int32_t* dCount;                           /* [maxBatch] */
    int32_t* dStatus;                          /* [maxBatch] sticky per-frame error flags */
    sdyn_keypoint* dKp; uint8_t* dDesc;        /* [maxBatch][maxKp] */
    /* matcher / dynamic-mask arena (grown on demand, reused across calls) */
    uint8_t* dArena; size_t arenaCap;
    sdyn::CameraModel camera;                  /* distortion model for mvKeysUn (sdyn_set_camera); disabled = mvKeysUn == mvKeys */
    sdyn_keypoint* dKpUn;                      /* [maxBatch][maxKp], allocated when a distorted camera is set */
    void* track;                               /* TrackState of the batched front end (sdyn_track.cpp) */
    void* stereo;                              /* StereoState of ComputeStereoMatches (sdyn_stereo.cpp) */
    void* bow;                                 /* BowState of ComputeBoW (sdyn_bow.cpp) */
    /* per-stage profiling */
    bool profiling;
    std::vector<cudaEvent_t> evPool;           /* free events */
    struct Span { cudaEvent_t a, b; int stage; };
    std::vector<Span> spans;                   /* recorded, not yet read */
    sdyn_stage_times acc;
    /* pinned staging */
    sdyn_keypoint* hKp; uint8_t* hDesc; int32_t* hCount; int32_t* hStatus;
};

namespace sdyn {

/* sdyn_api.cpp */
int ensure_geometry(sdyn_ctx* c, int W, int H);
cudaError_t upload_frames(sdyn_ctx* c, int nframes, const uint8_t* gray, size_t frameStride, int W, int H, int stride,
                          cudaStream_t st);
int fetch_enqueue(sdyn_ctx* c, int nframes, sdyn_keypoint* kpOut, uint8_t* descOut, int cap, cudaStream_t st);
int fetch_finish(sdyn_ctx* c, int nframes, sdyn_keypoint* kpOut, uint8_t* descOut, int cap, int* nOut);
int enqueue_extract(sdyn_ctx* c, int nframes, const uint8_t* dGray, size_t frameStride, int rowStride, cudaStream_t st);
struct StageTimer {          /* brackets a stage with CUDA events while profiling is enabled */
    sdyn_ctx* c; cudaStream_t st; int stage; cudaEvent_t a;
    StageTimer(sdyn_ctx* c, cudaStream_t st, int stage);
    ~StageTimer();
};
/* sdyn_track.cpp */
void free_track_state(sdyn_ctx* c);
/* sdyn_stereo.cpp */
void free_stereo_state(sdyn_ctx* c);
/* sdyn_bow.cpp */
void free_bow_state(sdyn_ctx* c);

/* geometry.cpp */
void compute_scale_info(const sdyn_orb_params& p, sdyn_scale_info& s, int umax[16]);
/* Fills g and the resize tables for image size W x H.  Returns SDYN_OK or SDYN_ERR_GEOMETRY. */
int compute_geometry(const sdyn_orb_params& p, const sdyn_scale_info& s, int W, int H,
                     Geom& g, std::vector<uint8_t>& tables);
int max_keypoints_per_frame(const sdyn_orb_params& p, const sdyn_scale_info& s, int maxW, int maxH);

/* kernels (each returns the cudaError_t of its launch) */
cudaError_t launch_level0(const Geom& g, const uint8_t* dIn, size_t inFrameStride, int inRowStride,
                          uint8_t* dPyr, int nframes, cudaStream_t st);
cudaError_t launch_resize(const Geom& g, int level, const uint8_t* dTables, const void* tmaMaps, uint8_t* dPyr,
                          int nframes, cudaStream_t st);
cudaError_t launch_fast(const Geom& g, const TileRef* tiles, int ntiles, const void* tmaMaps,
                        int iniTh, int minTh, uint8_t* dCellFlag, uint32_t* dCand, int32_t* dCandCount,
                        int nframes, cudaStream_t st);
cudaError_t launch_octree(const Geom& g, int iniTh, int minTh, const uint8_t* dCellFlag, uint32_t* dCand,
                          const int32_t* dCandCount, int32_t* dCandNode, int32_t* dSelCount, LevelKp* dLevelKp,
                          int32_t* dLevelCount, int32_t* dStatus, int nframes, cudaStream_t st);
cudaError_t launch_blur(const Geom& g, const TileRef* tiles, int ntiles, const void* tmaMaps,
                        uint8_t* dBlur, int nframes, cudaStream_t st);
cudaError_t launch_orient_describe(const Geom& g, const void* tmaMaps,
                                   const LevelKp* dLevelKp, const int32_t* dLevelCount,
                                   sdyn_keypoint* dKp, uint8_t* dDesc, int32_t* dCount, int maxKp,
                                   int nframes, cudaStream_t st);
size_t octree_smem_bytes(int nodeCap);
cudaError_t launch_undistort(const CameraModel& cam, const sdyn_keypoint* dKp, const int32_t* dCount, int cap,
                             sdyn_keypoint* dKpUn, int nframes, cudaStream_t st);
cudaError_t launch_undistort_xy(const CameraModel& cam, const float* dSrc, int n, float* dDst, cudaStream_t st);

/* FAST tiles: 124 x 30 window pixels (+ halo = 126 x 32 of the 128 x 32 score positions a CTA computes: 4 x 4 per thread).
 * Tile tx starts at window column 124 tx - kFastLead: 124 = 0 and -3 = 1 (mod 4) put the centre pixel of score column 0
 * on a 4-byte boundary of the padded row for every tile, so a TMA box (x origin a multiple of 16) delivers the tile with
 * the compass test's words aligned. */
constexpr int kFastTileW = 124, kFastTileH = 30, kFastLead = 3;
constexpr int kFastStageW = 160, kFastStageH = kFastTileH + 8;   /* staged box: 16 + 128 + 4 needed bytes; tile + 1 (NMS halo) + 3 (ring) rows each side */
constexpr int kBlurTileW = 128, kBlurTileH = 32;
constexpr int kBlurStageW = kBlurTileW + 32, kBlurStageH = kBlurTileH + 6;   /* staged box: columns x0-16 .. x0+W+15, rows y0-3 .. y0+H+2 */
constexpr int kPatchPitch = 80;            /* descriptor-stage patch box: 64 needed bytes, rows 20 banks apart */
constexpr int kOrientRows = 31, kDescRows = 37;
constexpr int kResizePitch = 192;          /* staging pitch of k_resize's common case (128-column tile, scale >= ~1.15) = its TMA box width */

}  // namespace sdyn
