"""Pins oracle/orc_stereo.cpp (Frame::ComputeStereoMatches, src/Frame.cc:874-1048): the C++ restatement must agree
bit-for-bit with an independent Python twin whose window arithmetic runs through cv2 (tests/stereo_twin.py)."""
import numpy as np
import pytest

import common
import orc
import scenario
import stereo_twin


def run_pair(cfg, idx, disparities):
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    left, right = scenario.stereo_pair(cfg, idx, disparities)
    EL, ER = orc.Extractor(nf, 1.2, 8, ini, mn), orc.Extractor(nf, 1.2, 8, ini, mn)
    kl, dl = EL(left); kr, dr = ER(right)
    return EL, ER, kl, dl, kr, dr


@pytest.mark.parametrize("cfg,idx,disp", [("small", 0, (5, 11, 23)), ("small", 3, (0, 2, 60)), ("tum", 1, (7, 19, 41))])
def test_oracle_matches_cv2_twin(cfg, idx, disp):
    EL, ER, kl, dl, kr, dr = run_pair(cfg, idx, disp)
    cam = scenario.KITTI_CAM
    mb, mbf = cam["bf"] / cam["fx"], cam["bf"]
    ur, dp, kept = orc.stereo_matches(EL, ER, kl, dl, kr, dr, mb, mbf)
    pyr_l = [EL.level(l)[19:-19, 19:-19] for l in range(8)]
    pyr_r = [ER.level(l)[19:-19, 19:-19] for l in range(8)]
    tur, tdp = stereo_twin.compute_stereo_matches(kl, dl, kr, dr, pyr_l, pyr_r, EL.scale, EL.inv_scale, mb, mbf)
    assert np.array_equal(ur.view(np.uint32), tur.view(np.uint32))
    assert np.array_equal(dp.view(np.uint32), tdp.view(np.uint32))
    assert kept == int((ur >= 0).sum())
    # the scenario must exercise the path: most keypoints find a stereo match, with band-dependent disparity
    assert kept > 0.3 * len(kl)
    if min(disp) > 1:
        d = (kl["x"] - ur)[ur >= 0]
        assert np.median(np.abs(d[:, None] - np.array(disp)[None, :]).min(1)) < 1.0


def test_no_right_keypoints_and_empty():
    EL, ER, kl, dl, kr, dr = run_pair("small", 0, (5, 11, 23))
    ur, dp, kept = orc.stereo_matches(EL, ER, kl, dl, kr[:0], dr[:0], 0.5, 380.0)
    assert kept == 0 and (ur == -1).all() and (dp == -1).all()
    ur, dp, kept = orc.stereo_matches(EL, ER, kl[:0], dl[:0], kr, dr, 0.5, 380.0)
    assert kept == 0 and len(ur) == 0
