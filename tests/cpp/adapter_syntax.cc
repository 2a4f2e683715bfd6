/* Syntax / type check of EVERY adapter template in slam-dynamic_b200/host/sdyn_adapters.hpp, including the ones that
 * need the real OpenCV matrix algebra (-DSDYN_HAVE_OPENCV): compiled with `g++ -fsyntax-only` against declaration-only
 * mocks of cv::Mat and of the reference's Frame / KeyFrame / MapPoint / DBoW2 containers (member names as in the
 * reference headers).  Nothing here is linked or run.  TEST INFRASTRUCTURE. */
#include <cstddef>
#include <map>
#include <set>
#include <vector>

namespace cv {
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
class MatExpr;
class Mat {
public:
    int rows, cols; unsigned char* data;
    Mat();
    Mat(const MatExpr&);
    bool empty() const;
    Mat rowRange(int, int) const; Mat colRange(int, int) const; Mat row(int) const; Mat col(int) const;
    MatExpr t() const;
    double dot(const Mat&) const;
    template <class T> T& at(int);
    template <class T> const T& at(int) const;
    template <class T> T& at(int, int);
    template <class T> const T& at(int, int) const;
};
class MatExpr { public: MatExpr(const Mat&); operator Mat() const; MatExpr t() const; };
MatExpr operator*(const Mat&, const Mat&); MatExpr operator*(const MatExpr&, const Mat&); MatExpr operator*(double, const Mat&);
MatExpr operator*(double, const MatExpr&);
MatExpr operator+(const MatExpr&, const Mat&); MatExpr operator+(const Mat&, const Mat&);
MatExpr operator/(const Mat&, double); MatExpr operator-(const MatExpr&); MatExpr operator-(const Mat&);
}  // namespace cv

#define SDYN_HAVE_OPENCV 1
#include "../../slam-dynamic_b200/host/sdyn_adapters.hpp"

namespace DBoW2 {
enum LNorm { L1, L2 };
struct BowVector { void addWeight(unsigned, double); void normalize(LNorm); void clear(); bool empty() const; };
struct FeatureVector : std::map<unsigned, std::vector<unsigned> > { void addFeature(unsigned, unsigned); };
}

struct KeyFrame;
struct MapPoint {
    bool mbTrackInView; int mnTrackScaleLevel; float mTrackViewCos, mTrackProjX, mTrackProjY, mTrackProjXR;
    bool isBad(); int Observations(); cv::Mat GetDescriptor(); cv::Mat GetWorldPos(); cv::Mat GetNormal();
    float GetMinDistanceInvariance(); float GetMaxDistanceInvariance();
protected:
    float mfMaxDistance;
public:
    bool IsInKeyFrame(KeyFrame*); void Replace(MapPoint*); void AddObservation(KeyFrame*, size_t); int GetIndexInKeyFrame(KeyFrame*);
};
struct Extractor { sdyn_ctx* Context(); };
struct KeyFrame {
    int N, mnScaleLevels; std::vector<cv::KeyPoint> mvKeysUn; cv::Mat mDescriptors; std::vector<float> mvuRight, mvScaleFactors,
        mvLevelSigma2, mvInvLevelSigma2; float mfLogScaleFactor, fx, fy, cx, cy, mbf; int mnMinX, mnMinY, mnMaxX, mnMaxY;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> GetMapPointMatches(); std::set<MapPoint*> GetMapPoints(); MapPoint* GetMapPoint(size_t);
    void AddMapPoint(MapPoint*, size_t); cv::Mat GetRotation(); cv::Mat GetTranslation(); cv::Mat GetCameraCenter();
};
struct Frame {
    int N, mnScaleLevels; std::vector<cv::KeyPoint> mvKeys, mvKeysUn; cv::Mat mDescriptors, mTcw;
    std::vector<float> mvuRight, mvDepth, mvScaleFactors; std::vector<MapPoint*> mvpMapPoints; std::vector<bool> mvbOutlier;
    DBoW2::FeatureVector mFeatVec; DBoW2::BowVector mBowVec; float mfLogScaleFactor, mb, mbf;
    static float mnMinX, mnMinY, mnMaxX, mnMaxY, fx, fy, cx, cy;
    Extractor* mpORBextractorLeft; Extractor* mpORBextractorRight;
};

/* one explicit use of every adapter, with the reference's argument types */
void use_all(sdyn_ctx* ctx, sdyn_vocab* voc, Frame& F, Frame& G, KeyFrame* k1, KeyFrame* k2, std::vector<MapPoint*>& mps,
             std::set<MapPoint*>& found, cv::Mat M, std::vector<cv::Point2f>& pts, std::vector<int>& m12,
             std::vector<std::pair<size_t, size_t> >& pairs, std::vector<cv::KeyPoint>& keys, std::vector<unsigned long long>& mask)
{
    sdyn_host::SearchByProjection(ctx, F, mps, 3.f, 0.8f);
    sdyn_host::SearchByProjection<Frame, cv::Point2f>(ctx, F, G, 7.f, false, true, &pts, &pts);
    sdyn_host::SearchForInitialization(ctx, F, G, pts, m12, 100, 0.9f, true);
    sdyn_host::SearchByBoW(ctx, k1, F, mps, 0.7f, true);
    sdyn_host::SearchByBoW(ctx, k1, k2, mps, 0.75f, true);
    sdyn_host::SearchByProjection(ctx, F, k1, found, 10.f, 100, true);
    sdyn_host::SearchByProjection(ctx, k1, M, mps, mps, 10);
    sdyn_host::Fuse(ctx, k1, mps, 3.f);
    sdyn_host::Fuse(ctx, k1, M, mps, 4.f, mps);
    float s12 = 1.f;
    sdyn_host::SearchBySim3(ctx, k1, k2, mps, s12, M, M, 7.5f);
    sdyn_host::SearchForTriangulation(ctx, k1, k2, M, pairs, false, true);
    sdyn_host::ComputeStereoMatches(F);
    sdyn_host::ComputeBoW(ctx, voc, F, DBoW2::L1);
    struct Rect2d { double x, y, w, h; };
    std::vector<Rect2d> boxes;
    std::vector<uint64_t> m64;
    sdyn_host::BoxMask(ctx, keys, boxes, m64);
    (void)mask;
}
