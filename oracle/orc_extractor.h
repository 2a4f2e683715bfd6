/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * CPU restatement of ORBextractor (reference: src/ORBextractor.cc, include/ORBextractor.h). */
#pragma once
#include <cstdint>
#include <vector>

namespace orc {

/* Same layout as cv::KeyPoint (28 bytes). */
struct KeyPoint {
    float x, y, size, angle, response;
    int octave, class_id;
};

struct Level {
    int w = 0, h = 0, stride = 0;         /* stride of the bordered buffer = w + 38 */
    std::vector<uint8_t> buf;             /* (w+38) x (h+38), interior at +19,+19 */
    const uint8_t* roi() const { return buf.data() + 19 * (size_t)stride + 19; }
    uint8_t* roi() { return buf.data() + 19 * (size_t)stride + 19; }
};

class Extractor {
public:
    /* ORBextractor.cc:410-470 */
    Extractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);

    /* ORBextractor.cc:1043-1105.  Returns #keypoints (keypoints level-major, octree list order);
     * desc is n x 32.  Returns -1 on an input the reference would crash on (see .cpp). */
    int run(const uint8_t* img, int w, int h, int stride,
            std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc);

    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;                   /* the reference stores the float ctor arg in a double member */
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;               /* mnFeaturesPerLevel */
    std::vector<int> umax;                /* orientation disc half-widths */
    std::vector<Level> pyr;               /* mvImagePyramid (bordered) */

    /* intermediate products kept for stage-level parity tests */
    std::vector<std::vector<int>> cand;   /* per level: FAST candidates (x,y,response) triples, cell order */
    std::vector<int> perLevel;            /* keypoints kept per level */

    void computePyramid(const uint8_t* img, int w, int h, int stride);    /* :1107-1132 */
    bool computeKeyPoints(std::vector<std::vector<KeyPoint>>& all);       /* :765-853 */
    /* :539-763; tie-break of equal-size nodes pinned to "later-created first" (Appendix B-1) */
    std::vector<KeyPoint> distributeOctTree(const std::vector<KeyPoint>& keys, int minX, int maxX,
                                            int minY, int maxY, int N, bool& ok);
};

float ic_angle(const uint8_t* center, int stride, const std::vector<int>& umax);      /* :77-104 */
void orb_descriptor(float angle_deg, const uint8_t* center, int stride, uint8_t* out); /* :108-147 */

}  // namespace orc
