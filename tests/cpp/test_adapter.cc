/* Exercises the drop-in adapter classes on a GPU and dumps inputs + outputs for tests/test_gpu_adapter.py, which
 * replays the inputs through the CPU oracle.  usage: test_adapter <out.bin> */
#include "../../slam-dynamic_b200/host/ORBextractor.h"
#include "../../slam-dynamic_b200/host/sdyn_adapters.hpp"
#include "../../slam-dynamic_b200/synth/sdyn_synth.h"
#include "ref_stubs.h"
#include <cstdio>
#include <string>

static FILE* g_out;
static void dump(const std::string& name, const void* p, size_t bytes)
{
    uint32_t n = (uint32_t)name.size(); uint64_t b = bytes;
    fwrite(&n, 4, 1, g_out); fwrite(name.data(), 1, n, g_out); fwrite(&b, 8, 1, g_out); if (bytes) fwrite(p, 1, bytes, g_out);
}
static uint64_t rnd(uint64_t& s) { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

using ref_stub::Frame; using ref_stub::MapPoint;

static void fill_frame(Frame& F, ORB_SLAM2::ORBextractor& ex, const cv::Mat& img)
{
    ex(img, cv::Mat(), F.mvKeys, F.mDescriptors);
    F.N = (int)F.mvKeys.size(); F.mvKeysUn = F.mvKeys;
    F.mvScaleFactors = ex.GetScaleFactors(); F.mnScaleLevels = ex.GetLevels();
    F.mvuRight.assign(F.N, -1.f);
    F.mvpMapPoints.assign(F.N, nullptr); F.mvbOutlier.assign(F.N, false);
    F.mTcw = cv::Mat(4, 4, CV_32F);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) F.mTcw.at<float>(r, c) = r == c ? 1.f : 0.f;
    F.mbf = 379.8145f; F.mb = F.mbf / Frame::fx;
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    g_out = fopen(argv[1], "wb");
    if (!g_out) return 2;
    const int W = 640, H = 480;
    Frame::mnMinX = 0; Frame::mnMinY = 0; Frame::mnMaxX = W; Frame::mnMaxY = H;
    Frame::fx = 707.0912f; Frame::fy = 707.0912f; Frame::cx = 601.8873f; Frame::cy = 183.1104f;
    cv::Mat img0(H, W, CV_8UC1), img1(H, W, CV_8UC1);
    sdyn_synth_frame(1007, 1000, W, H, 120, 0, 0, 0, img0.data, W);
    sdyn_synth_frame(1007, 1001, W, H, 120, 4, 1, 1, img1.data, W);
    ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    Frame last, cur;
    fill_frame(last, ex, img0);
    fill_frame(cur, ex, img1);
    if (cur.N == 0 || !ex.Context()) { fprintf(stderr, "extraction failed\n"); return 1; }
    dump("img1", img1.data, (size_t)W * H);
    dump("cur_keys", cur.mvKeys.data(), cur.mvKeys.size() * sizeof(cv::KeyPoint));
    dump("cur_desc", cur.mDescriptors.data, (size_t)cur.N * 32);
    dump("last_keys", last.mvKeys.data(), last.mvKeys.size() * sizeof(cv::KeyPoint));
    dump("last_desc", last.mDescriptors.data, (size_t)last.N * 32);
    const cv::Mat& L3 = ex.mvImagePyramid[3];
    std::vector<unsigned char> lvl((size_t)L3.rows * L3.cols);
    for (int r = 0; r < L3.rows; ++r) memcpy(&lvl[(size_t)r * L3.cols], L3.ptr(r), L3.cols);
    int dims[2] = {L3.cols, L3.rows};
    dump("level3_dims", dims, sizeof(dims)); dump("level3", lvl.data(), lvl.size());
    std::vector<float> sc = ex.GetScaleFactors();
    dump("scale", sc.data(), sc.size() * 4);

    /* LastFrame map points: back-projected so they land near their keypoint in the current frame */
    uint64_t seed = 42;
    std::vector<MapPoint> pts(last.N);
    std::vector<sdyn_last_point> lpDump(last.N);
    for (int i = 0; i < last.N; ++i) {
        memset(&lpDump[i], 0, sizeof(sdyn_last_point));
        if (rnd(seed) % 100 < 15) continue;
        MapPoint& p = pts[i];
        const float z = 4.f + (float)(rnd(seed) % 3600) / 100.f;
        p.pos = cv::Mat(3, 1, CV_32F);
        p.pos.at<float>(0, 0) = (last.mvKeys[i].pt.x - 4.f - Frame::cx) * z / Frame::fx;
        p.pos.at<float>(1, 0) = (last.mvKeys[i].pt.y - 1.f - Frame::cy) * z / Frame::fy;
        p.pos.at<float>(2, 0) = z;
        p.desc = last.mDescriptors.row(i).clone();
        p.nObs = (rnd(seed) % 10) ? 1 : 0;
        last.mvpMapPoints[i] = &p;
        last.mvbOutlier[i] = rnd(seed) % 20 == 0;
        lpDump[i].has_mp = 1; lpDump[i].outlier = last.mvbOutlier[i]; lpDump[i].obs_positive = p.nObs > 0;
        for (int k = 0; k < 3; ++k) lpDump[i].world[k] = p.pos.at<float>(k, 0);
        memcpy(lpDump[i].desc, p.desc.data, 32);
    }
    dump("last_points", lpDump.data(), lpDump.size() * sizeof(sdyn_last_point));
    std::vector<cv::Point2f> pl, pc;
    const int n1 = sdyn_host::SearchByProjection(ex.Context(), cur, last, 15.f, true, true, &pl, &pc);
    std::vector<int> assign1(cur.N, -1);
    for (int i = 0; i < cur.N; ++i) if (cur.mvpMapPoints[i]) assign1[i] = (int)(cur.mvpMapPoints[i] - pts.data());
    dump("frame_n", &n1, 4); dump("frame_assign", assign1.data(), assign1.size() * 4);
    std::vector<float> pairs;
    for (size_t k = 0; k < pl.size(); ++k) { pairs.push_back(pl[k].x); pairs.push_back(pl[k].y); pairs.push_back(pc[k].x); pairs.push_back(pc[k].y); }
    dump("frame_pairs", pairs.data(), pairs.size() * 4);

    /* local map */
    const int nmp = 1500;
    std::vector<MapPoint> mps(nmp); std::vector<MapPoint*> vp(nmp);
    std::vector<sdyn_mappoint_query> mq(nmp);
    for (int i = 0; i < nmp; ++i) {
        MapPoint& p = mps[i]; vp[i] = &p;
        const int src = (int)(rnd(seed) % cur.N);
        p.mbTrackInView = rnd(seed) % 10 != 0; p.bad = rnd(seed) % 30 == 0;
        p.mTrackProjX = cur.mvKeys[src].pt.x + (float)((int)(rnd(seed) % 9) - 4) * 0.5f;
        p.mTrackProjY = cur.mvKeys[src].pt.y + (float)((int)(rnd(seed) % 9) - 4) * 0.5f;
        p.mTrackProjXR = p.mTrackProjX - 10.f;
        p.mTrackViewCos = (rnd(seed) & 1) ? 0.9995f : 0.99f;
        p.mnTrackScaleLevel = std::min(7, cur.mvKeys[src].octave + (int)(rnd(seed) & 1));
        p.nObs = (rnd(seed) % 10) ? 2 : 0;
        p.desc = cur.mDescriptors.row(src).clone();
        p.desc.data[rnd(seed) % 32] ^= (unsigned char)(1u << (rnd(seed) % 8));
        sdyn_mappoint_query& m = mq[i]; memset(&m, 0, sizeof(m));
        m.proj_x = p.mTrackProjX; m.proj_y = p.mTrackProjY; m.proj_xr = p.mTrackProjXR; m.view_cos = p.mTrackViewCos;
        m.level = p.mnTrackScaleLevel; m.track_in_view = p.mbTrackInView; m.bad = p.bad; m.obs_positive = p.nObs > 0;
        memcpy(m.desc, p.desc.data, 32);
    }
    dump("map_points", mq.data(), mq.size() * sizeof(sdyn_mappoint_query));
    const int n2 = sdyn_host::SearchByProjection(ex.Context(), cur, vp, 3.f, 0.8f);
    std::vector<int> assign2(cur.N, -1);
    for (int i = 0; i < cur.N; ++i) {
        MapPoint* p = cur.mvpMapPoints[i];
        if (!p) continue;
        assign2[i] = (p >= mps.data() && p < mps.data() + nmp) ? 100000 + (int)(p - mps.data()) : (int)(p - pts.data());
    }
    dump("map_n", &n2, 4); dump("map_assign", assign2.data(), assign2.size() * 4);

    /* stereo constructor path (src/Frame.cc:151-160): left and right extraction, then ComputeStereoMatches */
    {
        cv::Mat imgR(H, W, CV_8UC1);
        sdyn_synth_frame(1007, 2001, W, H, 120, 4 + 9, 1, 1, imgR.data, W);     /* img1 content, 9 px disparity */
        ORB_SLAM2::ORBextractor exL(1000, 1.2f, 8, 20, 7), exR(1000, 1.2f, 8, 20, 7);
        Frame S;
        fill_frame(S, exL, img1);
        exR(imgR, cv::Mat(), S.mvKeysRight, S.mDescriptorsRight);
        S.mpORBextractorLeft = &exL; S.mpORBextractorRight = &exR;
        if (!sdyn_host::ComputeStereoMatches(S)) { fprintf(stderr, "stereo failed: %s\n", sdyn_last_error(exL.Context())); return 1; }
        dump("imgR", imgR.data, (size_t)W * H);
        dump("stereo_uright", S.mvuRight.data(), S.mvuRight.size() * 4);
        dump("stereo_depth", S.mvDepth.data(), S.mvDepth.size() * 4);
    }
    /* context re-creation: a larger image re-creates the extractor's GPU context; the camera model set through the
     * extractor must survive it, and a wide-then-tall sequence must settle on the envelope instead of thrashing */
    {
        ORB_SLAM2::ORBextractor ex2(500, 1.2f, 8, 20, 7);
        const float dist[5] = {0.262383f, -0.953104f, -0.005358f, 0.002628f, 1.163314f};
        if (ex2.SetCamera(517.306408f, 516.469215f, 318.643040f, 255.313989f, dist, 5) != SDYN_OK) { fprintf(stderr, "SetCamera failed\n"); return 1; }
        cv::Mat wide(240, 640, CV_8UC1), tall(480, 320, CV_8UC1);
        sdyn_synth_frame(1007, 3001, 640, 240, 60, 0, 0, 0, wide.data, 640);
        sdyn_synth_frame(1007, 3002, 320, 480, 60, 0, 0, 0, tall.data, 320);
        std::vector<cv::KeyPoint> k; cv::Mat d;
        ex2(wide, cv::Mat(), k, d);
        sdyn_ctx* c1 = ex2.Context();
        ex2(tall, cv::Mat(), k, d);                       /* taller than anything seen: re-created as 640 x 480 */
        sdyn_ctx* c2 = ex2.Context();
        ex2(wide, cv::Mat(), k, d);                       /* fits the envelope: no re-creation */
        if (!c1 || !c2 || ex2.Context() != c2 || k.empty()) { fprintf(stderr, "context envelope: unexpected re-creation\n"); return 1; }
        std::vector<cv::KeyPoint> ku(k.size());
        if (sdyn_fetch_keypoints_un(ex2.Context(), 1, reinterpret_cast<sdyn_keypoint*>(ku.data()), (int)ku.size(), nullptr) != SDYN_OK) {
            fprintf(stderr, "fetch_keypoints_un: %s\n", sdyn_last_error(ex2.Context())); return 1;
        }
        size_t moved = 0;
        for (size_t i = 0; i < k.size(); ++i) moved += ku[i].pt.x != k[i].pt.x || ku[i].pt.y != k[i].pt.y;
        if (moved * 2 < k.size()) { fprintf(stderr, "camera model lost on context re-creation (%zu of %zu keypoints undistorted)\n", moved, k.size()); return 1; }
        dump("recreate_keys", k.data(), k.size() * sizeof(cv::KeyPoint));
        dump("recreate_keys_un", ku.data(), ku.size() * sizeof(cv::KeyPoint));
        dump("recreate_img", wide.data, (size_t)640 * 240);
    }
    fclose(g_out);
    printf("adapter ok: %d keypoints, %d frame matches, %d map matches\n", cur.N, n1, n2);
    return 0;
}
