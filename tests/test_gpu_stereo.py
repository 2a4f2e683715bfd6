"""Parity of the CUDA stereo association (Frame::ComputeStereoMatches, src/Frame.cc:874-1048) through the C ABI
against the CPU oracle.  Bar: mvuRight and mvDepth bit-exact (they are float results of integer window sums and a
handful of IEEE operations evaluated in the reference's order)."""
import numpy as np
import pytest

import common
import orc
import pysdyn
import scenario

pytestmark = pytest.mark.gpu

CAM = scenario.KITTI_CAM
MB, MBF = CAM["bf"] / CAM["fx"], CAM["bf"]


def oracle_pair(cfg, left, right):
    _, _, _, nf, ini, mn = common.CONFIGS[cfg]
    EL, ER = orc.Extractor(nf, 1.2, 8, ini, mn), orc.Extractor(nf, 1.2, 8, ini, mn)
    kl, dl = EL(left); kr, dr = ER(right)
    ur, dp, kept = orc.stereo_matches(EL, ER, kl, dl, kr, dr, MB, MBF)
    return kl, ur, dp, kept


@pytest.mark.parametrize("cfg,idx,disp", [("kitti", 0, (5, 11, 23)), ("kitti", 5, (0, 3, 90)), ("tum", 2, (7, 19, 41)),
                                          ("small", 1, (2, 6, 14))])
def test_stereo_matches_single(cfg, idx, disp):
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    left, right = scenario.stereo_pair(cfg, idx, disp)
    L = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H)
    R = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H)
    kl, _ = L(left); R(right)
    ur, dp, kept = pysdyn.stereo_match(L, R, 1, MB, MBF)
    okl, our, odp, okept = oracle_pair(cfg, left, right)
    n = len(okl)
    assert len(kl) == n
    assert np.array_equal(ur[0, :n].view(np.uint32), our.view(np.uint32))
    assert np.array_equal(dp[0, :n].view(np.uint32), odp.view(np.uint32))
    assert kept[0] == okept and okept > 0.3 * n
    L.close(); R.close()


def test_stereo_matches_batch_device():
    """Batched, device-resident: 4 stereo pairs per step through extract_batch_device on two contexts."""
    import torch
    cfg = "kitti"
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    B = 4
    pairs = [scenario.stereo_pair(cfg, 10 + i, (4 + i, 12, 30 - 3 * i)) for i in range(B)]
    L = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    R = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    dl = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    dr = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    torch.cuda.synchronize()
    for _ in range(2):      # twice: the second step reuses every buffer
        L.extract_batch_device(dl.data_ptr(), B, W * H, W, H, W)
        R.extract_batch_device(dr.data_ptr(), B, W * H, W, H, W)
        pysdyn.stereo_match_device(L, R, B, MB, MBF)
    ur, dp, kept = pysdyn.stereo_fetch(L, B)
    for i, (left, right) in enumerate(pairs):
        okl, our, odp, okept = oracle_pair(cfg, left, right)
        n = len(okl)
        assert np.array_equal(ur[i, :n].view(np.uint32), our.view(np.uint32)), i
        assert np.array_equal(dp[i, :n].view(np.uint32), odp.view(np.uint32)), i
        assert kept[i] == okept
    L.close(); R.close()


def test_stereo_rejects_mismatched_contexts():
    L = pysdyn.Extractor(500, 1.2, 8, 20, 7, max_width=320, max_height=240)
    R = pysdyn.Extractor(500, 1.2, 8, 20, 7, max_width=640, max_height=480)
    L(common.frame("small", 0)); R(common.frame("tum", 0))
    with pytest.raises(pysdyn.SdynError):
        pysdyn.stereo_match(L, R, 1, MB, MBF)
    L.close(); R.close()


def test_stereo_featureless_views():
    """No keypoints on one side (a blank image): every mvuRight / mvDepth stays -1 and nothing is kept, on both the
    device and the oracle (the reference would read vDistIdx[0] of an empty vector here)."""
    cfg = "small"
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    textured = common.frame(cfg, 2)
    blank = np.full((H, W), 127, np.uint8)
    for left, right in ((textured, blank), (blank, textured)):
        L = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H)
        R = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H)
        kl, _ = L(left); kr, _ = R(right)
        assert len(kl) == 0 or len(kr) == 0
        ur, dp, kept = pysdyn.stereo_match(L, R, 1, MB, MBF)
        assert kept[0] == 0 and (ur[0, :len(kl)] == -1).all() and (dp[0, :len(kl)] == -1).all()
        okl, our, odp, okept = oracle_pair(cfg, left, right)
        assert okept == 0 and len(okl) == len(kl)
        L.close(); R.close()
