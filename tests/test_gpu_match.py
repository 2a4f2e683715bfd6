"""Parity of the CUDA matcher and dynamic-mask kernels (through the C ABI) against the CPU oracle.
Bar: match index arrays, lock flags, match counts, point pairs and the dynamic mask are bit-exact."""
import numpy as np
import pytest

import common
import orc
import pysdyn
import scenario

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    ex = pysdyn.Extractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=1)
    yield ex
    ex.close()


def pair(cfg, shift=(4, 1)):
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    E = orc.Extractor(nf, 1.2, 8, ini, mn)
    k0, d0 = E(common.frame(cfg, 0))
    k1, d1 = E(common.frame(cfg, 1, ox=shift[0], oy=shift[1], t=1))
    return dict(W=W, H=H, scale=E.scale, k0=k0, d0=d0, k1=k1, d1=d1, shift=shift)


@pytest.fixture(scope="module")
def tum():
    return pair("tum")


@pytest.fixture(scope="module")
def kitti():
    return pair("kitti", (3, 2))


def views(p, stereo, tcw=None, d0=None, d1=None):
    cur = scenario.frame_view(p["k1"], p["d1"] if d1 is None else d1, p["scale"], p["W"], p["H"], stereo=stereo, seed=1, tcw=tcw)
    last = scenario.frame_view(p["k0"], p["d0"] if d0 is None else d0, p["scale"], p["W"], p["H"], stereo=stereo, seed=0)
    return cur, last


def test_descriptor_distance():
    r = np.random.default_rng(0)
    a = r.integers(0, 256, (64, 32), dtype=np.uint8); b = r.integers(0, 256, (64, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert pysdyn.Matcher.DescriptorDistance(x, y) == orc.hamming(x, y)


@pytest.mark.parametrize("data", ["tum", "kitti"])
@pytest.mark.parametrize("stereo,th,tz", [(True, 7.0, 0.0), (False, 15.0, 0.0), (True, 7.0, 1.5), (True, 7.0, -1.5)])
def test_search_by_projection_frame(ctx, request, data, stereo, th, tz):
    p = request.getfixturevalue(data)
    tcw = np.eye(4, dtype=np.float32)[:3].copy(); tcw[2, 3] = tz
    cur, last = views(p, stereo, tcw)
    lp = scenario.last_points(p["k0"], p["d0"], p["shift"], seed=7)
    m = pysdyn.Matcher(ctx, 0.9, True)
    got = m.SearchByProjectionFrame(cur, last, lp, th, not stereo, want_pairs=True)
    ref = orc.match_projection_frame(cur, last, lp, th, not stereo, True, want_pairs=True)
    assert got[0] == ref[0] and got[0] > 50
    assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])
    # no orientation check + pre-occupied keypoints (locked and unlocked)
    r = np.random.default_rng(5)
    a0 = np.where(r.random(cur.n) < 0.2, -2, -1).astype(np.int32); l0 = (r.random(cur.n) < 0.5).astype(np.uint8)
    m2 = pysdyn.Matcher(ctx, 0.9, False)
    got = m2.SearchByProjectionFrame(cur, last, lp, th, not stereo, a0, l0)
    ref = orc.match_projection_frame(cur, last, lp, th, not stereo, False, a0, l0)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def test_search_by_projection_frame_tie_heavy(ctx, tum):
    p = tum
    d0 = scenario.degenerate_descriptors(len(p["k0"]), 1); d1 = scenario.degenerate_descriptors(len(p["k1"]), 2)
    cur, last = views(p, True, d0=d0, d1=d1)
    for p_obs in (0.5, 1.0, 0.0):
        lp = scenario.last_points(p["k0"], d0, p["shift"], seed=9, p_obs=p_obs, noise_bits=0)
        got = pysdyn.Matcher(ctx, 0.9, True).SearchByProjectionFrame(cur, last, lp, 15.0, False, want_pairs=True)
        ref = orc.match_projection_frame(cur, last, lp, 15.0, False, True, want_pairs=True)
        assert got[0] == ref[0]
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])


@pytest.mark.parametrize("data", ["tum", "kitti"])
@pytest.mark.parametrize("th", [1.0, 3.0, 5.0])
def test_search_by_projection_map(ctx, request, data, th):
    p = request.getfixturevalue(data)
    cur, _ = views(p, True)
    mp = scenario.map_queries(p["k1"], p["d1"], 8, seed=3, count=3000)
    r = np.random.default_rng(6)
    a0 = np.where(r.random(cur.n) < 0.3, -2, -1).astype(np.int32); l0 = (r.random(cur.n) < 0.7).astype(np.uint8)
    for nnratio in (0.8, 0.6):
        got = pysdyn.Matcher(ctx, nnratio).SearchByProjectionMap(cur, mp, th, a0, l0)
        ref = orc.match_projection_map(cur, mp, th, nnratio, a0, l0)
        assert got[0] == ref[0] and got[0] > 50
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def test_search_by_projection_map_tie_heavy(ctx, tum):
    p = tum
    d1 = scenario.degenerate_descriptors(len(p["k1"]), 2)
    cur, _ = views(p, False, d1=d1)
    mp = scenario.map_queries(p["k1"], d1, 8, seed=4, count=2500, jitter=6.0)
    mp["desc"] = d1[np.random.default_rng(8).integers(0, len(d1), len(mp))]
    mp["obs_positive"] = np.random.default_rng(9).random(len(mp)) < 0.5
    got = pysdyn.Matcher(ctx, 0.8).SearchByProjectionMap(cur, mp, 5.0)
    ref = orc.match_projection_map(cur, mp, 5.0, 0.8)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


@pytest.mark.parametrize("distinct,flip,nnratio", [(6, 3, 0.8), (10, 6, 0.6), (4, 2, 0.9), (16, 10, 0.7)])
def test_search_by_projection_map_contention(ctx, tum, distinct, flip, nnratio):
    """Many map points per keypoint, descriptors a few bit flips away from a handful of patterns, every point locking:
    long claim chains, and queries whose second-best candidate gets locked so that the ratio test flips (the case in
    which the parallel resolution has to rebuild its lock times)."""
    p = tum
    r = np.random.default_rng(100 + distinct)
    n = len(p["k1"])
    d1 = scenario.degenerate_descriptors(n, 20 + distinct, distinct).copy()
    noise = r.integers(0, 256, d1.shape, dtype=np.uint8) & (r.random(d1.shape) < flip / 256.0 * 8).astype(np.uint8) * 0xff
    d1 ^= noise & r.integers(0, 256, d1.shape, dtype=np.uint8)
    cur, _ = views(p, False, d1=d1)
    mp = scenario.map_queries(p["k1"], d1, 8, seed=40 + distinct, count=6000, jitter=9.0)
    md = d1[r.integers(0, n, len(mp))].copy()
    md ^= (r.random(md.shape) < 0.03).astype(np.uint8) * r.integers(0, 256, md.shape, dtype=np.uint8)
    mp["desc"] = md
    mp["obs_positive"] = r.random(len(mp)) < 0.9
    for th in (1.0, 6.0):
        got = pysdyn.Matcher(ctx, nnratio).SearchByProjectionMap(cur, mp, th)
        ref = orc.match_projection_map(cur, mp, th, nnratio)
        assert got[0] == ref[0] and got[0] > 20, (got[0], ref[0])
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


@pytest.mark.parametrize("data", ["tum", "kitti"])
def test_search_for_initialization(ctx, request, data):
    p = request.getfixturevalue(data)
    cur, last = views(p, False)
    prev = np.stack([p["k0"]["x"], p["k0"]["y"]], 1)
    for check in (True, False):
        got = pysdyn.Matcher(ctx, 0.9, check).SearchForInitialization(last, cur, prev, 100)
        ref = orc.match_init(last, cur, prev, 100, 0.9, check)
        assert got[0] == ref[0] and got[0] > 30
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    # tie-heavy: stealing (vnMatches21) and the matched-distance gate are exercised constantly
    d0 = scenario.degenerate_descriptors(len(p["k0"]), 11, 40); d1 = scenario.degenerate_descriptors(len(p["k1"]), 12, 40)
    d1[: len(d1) // 2] = d0[: len(d1) // 2][: len(d1[: len(d1) // 2])] if len(d0) >= len(d1) // 2 else d1[: len(d1) // 2]
    cur, last = views(p, False, d0=d0, d1=d1)
    got = pysdyn.Matcher(ctx, 0.9, True).SearchForInitialization(last, cur, prev, 100)
    ref = orc.match_init(last, cur, prev, 100, 0.9, True)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


@pytest.mark.parametrize("data", ["tum", "kitti"])
def test_search_by_bow(ctx, request, data):
    p = request.getfixturevalue(data)
    cur, last = views(p, False)
    fa = pysdyn.FeatureVector(scenario.bow_nodes(p["d0"])); fb = pysdyn.FeatureVector(scenario.bow_nodes(p["d1"]))
    valid = (np.random.default_rng(2).random(last.n) < 0.8).astype(np.uint8)
    for check in (True, False):
        got = pysdyn.Matcher(ctx, 0.7, check).SearchByBoW(last, valid, fa, cur, fb)
        ref = orc.match_bow(last, valid, fa, cur, fb, 0.7, check)
        assert got[0] == ref[0] and got[0] > 30 and np.array_equal(got[1], ref[1])
    d0 = scenario.degenerate_descriptors(last.n, 21, 30); d1 = scenario.degenerate_descriptors(cur.n, 22, 30)
    cur, last = views(p, False, d0=d0, d1=d1)
    fa = pysdyn.FeatureVector(scenario.bow_nodes(d0, 3)); fb = pysdyn.FeatureVector(scenario.bow_nodes(d1, 3))
    got = pysdyn.Matcher(ctx, 0.95, True).SearchByBoW(last, valid, fa, cur, fb)
    ref = orc.match_bow(last, valid, fa, cur, fb, 0.95, True)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1])


@pytest.mark.parametrize("data", ["tum", "kitti"])
def test_search_by_bow_keyframes(ctx, request, data):
    """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), src/ORBmatcher.cc:679-812 (loop detection)."""
    p = request.getfixturevalue(data)
    kf2, kf1 = views(p, False)
    fa = pysdyn.FeatureVector(scenario.bow_nodes(p["d0"])); fb = pysdyn.FeatureVector(scenario.bow_nodes(p["d1"]))
    r = np.random.default_rng(5)
    v1 = (r.random(kf1.n) < 0.8).astype(np.uint8); v2 = (r.random(kf2.n) < 0.7).astype(np.uint8)
    for check in (True, False):
        got = pysdyn.Matcher(ctx, 0.75, check).SearchByBoWKF(kf1, v1, fa, kf2, v2, fb)
        ref = orc.match_bow_kf(kf1, v1, fa, kf2, v2, fb, 0.75, check)
        assert got[0] == ref[0] and got[0] > 30 and np.array_equal(got[1], ref[1])
        assert (v1[got[1] >= 0] == 1).all() and (v2[got[1][got[1] >= 0]] == 1).all()
    # tie-heavy descriptors, best distance exactly TH_LOW must be rejected (strict <)
    d0 = scenario.degenerate_descriptors(kf1.n, 21, 30); d1 = scenario.degenerate_descriptors(kf2.n, 22, 30)
    kf2, kf1 = views(p, False, d0=d0, d1=d1)
    fa = pysdyn.FeatureVector(scenario.bow_nodes(d0, 3)); fb = pysdyn.FeatureVector(scenario.bow_nodes(d1, 3))
    got = pysdyn.Matcher(ctx, 0.95, True).SearchByBoWKF(kf1, v1, fa, kf2, v2, fb)
    ref = orc.match_bow_kf(kf1, v1, fa, kf2, v2, fb, 0.95, True)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1])


@pytest.mark.parametrize("data", ["tum", "kitti"])
@pytest.mark.parametrize("variant,th,maxd", [(pysdyn.PROJ_FRAME_KEYFRAME, 10.0, 100), (pysdyn.PROJ_FRAME_KEYFRAME, 3.0, 64),
                                             (pysdyn.PROJ_KEYFRAME_SIM3, 10.0, 50)])
def test_search_by_projection_pose(ctx, request, data, variant, th, maxd):
    """SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist) :1629-1756 and SearchByProjection(KeyFrame*, Scw, ...) :290-403."""
    p = request.getfixturevalue(data)
    target, _ = views(p, False)
    R, tcw, ow = scenario.pose_small(seed=3)
    pts = scenario.proj_points(p["k1"], p["d1"], p["scale"], R, tcw, ow, seed=11)
    prm = pysdyn.proj_params(R, tcw, ow, th, maxd, variant, True, np.log(np.float32(1.2)), 8)
    occ = np.where(np.random.default_rng(4).random(target.n) < 0.1, -2, -1).astype(np.int32)   # already matched keypoints
    got = pysdyn.Matcher(ctx, 0.9, True).SearchByProjectionPose(target, pts, prm, occ)
    ref = orc.match_projection_pose(target, pts, prm, occ)
    assert got[0] == ref[0] and got[0] > 100 and np.array_equal(got[1], ref[1])
    # tie-heavy descriptors: claims and the rotation cull interact
    dd = scenario.degenerate_descriptors(target.n, 5, 30)
    target2, _ = views(p, False, d1=dd)
    pts2 = scenario.proj_points(p["k1"], dd, p["scale"], R, tcw, ow, seed=12, noise_bits=0)
    got = pysdyn.Matcher(ctx, 0.9, True).SearchByProjectionPose(target2, pts2, prm)
    ref = orc.match_projection_pose(target2, pts2, prm)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1])


def rt(R, t):
    return np.concatenate([np.asarray(R, np.float32), np.asarray(t, np.float32).reshape(3, 1)], 1)


@pytest.mark.parametrize("data", ["tum", "kitti"])
@pytest.mark.parametrize("which,stereo", [(0, True), (0, False), (1, False)])
def test_fuse_search(ctx, request, data, which, stereo):
    """Search parts of the two ORBmatcher::Fuse overloads (:982-1100, :1132-1237)."""
    p = request.getfixturevalue(data)
    target, _ = views(p, stereo)
    R, tcw, ow = scenario.pose_small(seed=5)
    pts = scenario.proj_points(p["k1"], p["d1"], p["scale"], R, tcw, ow, seed=13, jitter=1.5)
    inv_s2 = (1.0 / (p["scale"] * p["scale"])).astype(np.float32)
    log_sf = np.log(np.float32(1.2))
    prm = pysdyn.best_params(rt(R, tcw), 3.0, log_sf, 8, ow=ow, invz_double=(which == 1), check_normal=True, chi2_gate=(which == 0),
                             bf=target.cam[4], inv_level_sigma2=inv_s2)
    gi, gd = pysdyn.Matcher(ctx).ProjectionBest(target, pts, prm)
    ri, rd = orc.fuse_search(which, target, inv_s2, pts, R, tcw, ow, 3.0, log_sf, 8)
    assert np.array_equal(gi, ri) and np.array_equal(gd, rd)
    assert ((ri >= 0) & (rd <= 50)).sum() > 100 and (ri < 0).sum() > 10


@pytest.mark.parametrize("data", ["tum", "kitti"])
def test_search_by_sim3(ctx, request, data):
    """ORBmatcher::SearchBySim3 (:1259-1483): two chained transforms per point, both directions, agreement check."""
    p = request.getfixturevalue(data)
    kf2, kf1 = views(p, False)
    eye = np.eye(3, dtype=np.float32)
    R12, t12, _ = scenario.pose_small(seed=6, angle_deg=0.8, t=(0.03, 0.01, -0.05))
    s12 = np.float32(1.03)
    sR12 = (s12 * R12).astype(np.float32)
    sR21 = ((np.float32(1.0) / s12) * R12.T).astype(np.float32)
    t21 = (-(sR21 @ t12)).astype(np.float32)
    zero = np.zeros(3, np.float32)
    # MapPoints of each keyframe: back-projected through identity poses; the shift between the two synthetic frames
    # plays the part of the similarity error the search has to absorb
    pts1 = scenario.proj_points(p["k0"], p["d0"], p["scale"], eye, zero, zero, seed=21, p_valid=0.8)
    pts2 = scenario.proj_points(p["k1"], p["d1"], p["scale"], eye, zero, zero, seed=22, p_valid=0.8)
    log_sf = np.log(np.float32(1.2))
    args = (kf1, kf2, pts1, pts2, rt(eye, zero), rt(eye, zero), rt(sR12, t12), rt(sR21, t21), 7.5, log_sf, 8)
    got = pysdyn.Matcher(ctx).SearchBySim3(*args)
    ref = orc.search_by_sim3(*args)
    assert got[0] == ref[0] and got[0] > 50 and np.array_equal(got[1], ref[1])


@pytest.mark.parametrize("data", ["tum", "kitti"])
@pytest.mark.parametrize("stereo,only_stereo", [(False, False), (True, False), (True, True)])
def test_search_for_triangulation(ctx, request, data, stereo, only_stereo):
    """ORBmatcher::SearchForTriangulation (:814-980) incl. CheckDistEpipolarLine and the epipole exclusion zone."""
    p = request.getfixturevalue(data)
    kf2, kf1 = views(p, stereo)
    fa = pysdyn.FeatureVector(scenario.bow_nodes(p["d0"], 4)); fb = pysdyn.FeatureVector(scenario.bow_nodes(p["d1"], 4))
    r = np.random.default_rng(8)
    h1 = (r.random(kf1.n) < 0.3).astype(np.uint8); h2 = (r.random(kf2.n) < 0.3).astype(np.uint8)
    # fundamental matrix of a pure image shift (dx, dy): lines through x1 + shift with direction along the shift, wide
    # enough a band that both accepted and rejected pairs occur; epipole placed inside the image to exercise :886-892
    dx, dy = p["shift"]
    F12 = np.array([[0, 0, dy], [0, 0, -dx], [-dy, dx, 0]], np.float32) * np.float32(0.01)
    sig2 = (p["scale"] * p["scale"]).astype(np.float32)
    for check in (True, False):
        prm = pysdyn.tri_params(F12, (p["W"] * 0.4, p["H"] * 0.5), only_stereo, check, sig2)
        got = pysdyn.Matcher(ctx, 0.6, check).SearchForTriangulation(kf1, h1, fa, kf2, h2, fb, prm)
        ref = orc.match_triangulation(kf1, h1, fa, kf2, h2, fb, prm)
        assert got[0] == ref[0] and np.array_equal(got[1], ref[1])
        assert got[0] > 20 and (h1[got[1] >= 0] == 0).all() and (h2[got[1][got[1] >= 0]] == 0).all()
    # tie-heavy descriptors: the LAST candidate among equal distances wins (`dist > bestDist`)
    d0 = scenario.degenerate_descriptors(kf1.n, 31, 30); d1 = scenario.degenerate_descriptors(kf2.n, 32, 30)
    kf2, kf1 = views(p, stereo, d0=d0, d1=d1)
    fa = pysdyn.FeatureVector(scenario.bow_nodes(d0, 2)); fb = pysdyn.FeatureVector(scenario.bow_nodes(d1, 2))
    prm = pysdyn.tri_params(F12, (-500.0, -500.0), only_stereo, True, sig2)
    got = pysdyn.Matcher(ctx, 0.6, True).SearchForTriangulation(kf1, h1, fa, kf2, h2, fb, prm)
    ref = orc.match_triangulation(kf1, h1, fa, kf2, h2, fb, prm)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1])


def test_empty_inputs(ctx, tum):
    p = tum
    cur, last = views(p, False)
    m = pysdyn.Matcher(ctx)
    n, a, l = m.SearchByProjectionMap(cur, np.zeros(0, pysdyn.MAPPOINT_DTYPE), 3.0)
    assert n == 0 and (a == -1).all()
    empty = scenario.frame_view(p["k0"][:0], p["d0"][:0], p["scale"], p["W"], p["H"])
    n, a, l = m.SearchByProjectionFrame(cur, empty, np.zeros(0, pysdyn.LASTPOINT_DTYPE), 7.0, True)
    assert n == 0 and (a == -1).all()
    n, m12, prev = m.SearchForInitialization(empty, cur, np.zeros((0, 2), np.float32))
    assert n == 0 and len(m12) == 0


def test_box_mask(ctx, kitti):
    p = kitti
    boxes = pysdyn.synth_boxes(7, p["W"], p["H"], 160, 3, 2, 1, margin=10)
    assert len(boxes) >= 3
    boxes = np.concatenate([boxes, [[100.0, 50.0, 0.0, 0.0], [p["k1"]["x"][0], p["k1"]["y"][0], 1.0, 1.0]]])  # empty box; box edge on a keypoint
    got = pysdyn.box_mask(ctx, p["k1"], boxes)
    ref = orc.box_mask(p["k1"], boxes)
    assert np.array_equal(got, ref) and (ref != 0).sum() > 10
    assert np.array_equal(pysdyn.box_mask(ctx, p["k1"], np.zeros((0, 4))), np.zeros(len(p["k1"]), np.uint64))


@pytest.mark.parametrize("mode", [0, 1])
def test_separate_pairs(ctx, kitti, mode):
    """BFMatcher(crossCheck) + classifyF / classifyH per box pair."""
    p = kitti
    r = np.random.default_rng(3)
    pairs = []
    for b in range(6):
        qi = r.choice(len(p["k1"]), int(r.integers(0, 120)), replace=False)
        ti = r.choice(len(p["k0"]), int(r.integers(0, 120)), replace=False)
        qd, td = p["d1"][qi], p["d0"][ti]
        if b == 4:
            qd = scenario.degenerate_descriptors(len(qi), 5, 4); td = scenario.degenerate_descriptors(len(ti), 6, 4)
        qx = np.stack([p["k1"]["x"][qi], p["k1"]["y"][qi]], 1); tx = np.stack([p["k0"]["x"][ti], p["k0"]["y"][ti]], 1)
        if b == 5 and len(ti):
            tx = qx[: len(ti)] + np.float32([3, 2]) if len(qi) >= len(ti) else tx      # geometrically consistent pairs
        pairs.append((qd, qx, td, tx))
    if mode == 0:   # F of a pure image translation (3, 2): x2^T F x1 = 0 for x2 = x1 - t
        M = np.float32([[0, 0, 2], [0, 0, -3], [-2, 3, 0]])
    else:
        M = np.float32([[1, 0, -3], [0, 1, -2], [1e-5, 0, 1]])
    got = pysdyn.separate_pairs(ctx, pairs, M, mode)
    ref = orc.separate_pairs(pairs, M, mode)
    for g, e in zip(got, ref):
        for a, b in zip(g, e):
            assert np.array_equal(a, b)
    assert sum(len(g[0]) for g in got) > 20
