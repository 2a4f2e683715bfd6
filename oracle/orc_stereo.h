/* ORACLE — TEST INFRASTRUCTURE ONLY (see orc_prims.h).
 * CPU restatement of Frame::ComputeStereoMatches (reference: src/Frame.cc:874-1048).
 * Parity pin: tests/test_oracle_stereo.py — an independent Python twin whose window arithmetic runs through cv2
 * (convertTo, subtract, cv2.norm(NORM_L1)) must agree bit for bit (tests/stereo_twin.py). */
#pragma once
#include "orc_extractor.h"
#include <cstdint>

namespace orc {

/* One pyramid level as the reference sees it: mvImagePyramid[l] is the un-bordered ROI. */
struct LevelRef { const uint8_t* roi; int w, h, stride; };

/* Writes mvuRight / mvDepth (N floats each, -1 = no match); returns the number of stereo points kept, or -1 on
 * inputs the reference would index out of range with (a keypoint row outside the image). */
int compute_stereo_matches(const KeyPoint* keysL, int N, const uint8_t* descL,
                           const KeyPoint* keysR, int Nr, const uint8_t* descR,
                           const LevelRef* pyrL, const LevelRef* pyrR, int nlevels,
                           const float* scaleFactors, const float* invScaleFactors,
                           float mb, float mbf, float* uRight, float* depth);

}  // namespace orc
