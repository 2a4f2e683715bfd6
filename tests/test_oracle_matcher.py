"""Pins for the oracle's matcher / dynamic-mask restatement (CPU only).
The reference ships no tests; these check the oracle against cv2 4.13.0 where OpenCV is the
dependency (BFMatcher, invert, gemm) and against independent brute-force re-statements elsewhere."""
import math
import numpy as np
import pytest

import common
import orc
import pysdyn
import scenario

cv2 = pytest.importorskip("cv2")
f32 = np.float32


@pytest.fixture(scope="module")
def tum_pair():
    W, H, _, nf, ini, mn = common.CONFIGS["tum"]
    E = orc.Extractor(nf, 1.2, 8, ini, mn)
    k0, d0 = E(common.frame("tum", 0))
    k1, d1 = E(common.frame("tum", 1, ox=4, oy=1, t=1))
    return dict(W=W, H=H, scale=E.scale, k0=k0, d0=d0, k1=k1, d1=d1)


def test_hamming_matches_numpy_popcount():
    r = np.random.default_rng(1)
    a = r.integers(0, 256, (200, 32), dtype=np.uint8); b = r.integers(0, 256, (200, 32), dtype=np.uint8)
    ref = np.unpackbits(a ^ b, axis=1).sum(1)
    assert [orc.hamming(x, y) for x, y in zip(a, b)] == ref.tolist()
    assert orc.hamming(a[0], a[0]) == 0 and orc.hamming(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def py_features_in_area(F, x, y, r, min_level, max_level):
    """Independent re-statement of Frame::GetFeaturesInArea + AssignFeaturesToGrid (Frame.cc:463-478, 735-800)."""
    minx, miny, maxx, maxy = (f32(v) for v in F.bounds)
    winv, hinv = f32(64) / f32(maxx - minx), f32(48) / f32(maxy - miny)
    grid = {}
    for i, kp in enumerate(F.keys_un):
        px = f32(f32(kp["x"] - minx) * winv); py = f32(f32(kp["y"] - miny) * hinv)
        gx = int(math.floor(abs(px) + 0.5)) * (1 if px >= 0 else -1)      # C round(): half away from zero
        gy = int(math.floor(abs(py) + 0.5)) * (1 if py >= 0 else -1)
        if 0 <= gx < 64 and 0 <= gy < 48:
            grid.setdefault((gx, gy), []).append(i)
    x, y, r = f32(x), f32(y), f32(r)
    cx0 = max(0, int(math.floor(f32(f32(f32(x - minx) - r) * winv))))
    cx1 = min(63, int(math.ceil(f32(f32(f32(x - minx) + r) * winv))))
    cy0 = max(0, int(math.floor(f32(f32(f32(y - miny) - r) * hinv))))
    cy1 = min(47, int(math.ceil(f32(f32(f32(y - miny) + r) * hinv))))
    if cx0 >= 64 or cx1 < 0 or cy0 >= 48 or cy1 < 0:
        return []
    check = min_level > 0 or max_level >= 0
    out = []
    for ix in range(cx0, cx1 + 1):
        for iy in range(cy0, cy1 + 1):
            for i in grid.get((ix, iy), []):
                kp = F.keys_un[i]
                if check and (kp["octave"] < min_level or (max_level >= 0 and kp["octave"] > max_level)):
                    continue
                if abs(f32(kp["x"] - x)) < r and abs(f32(kp["y"] - y)) < r:
                    out.append(i)
    return out


def test_features_in_area_order_and_filters(tum_pair):
    p = tum_pair
    F = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"])
    r = np.random.default_rng(2)
    for _ in range(60):
        x, y = float(r.uniform(-30, p["W"] + 30)), float(r.uniform(-30, p["H"] + 30))
        rad = float(r.choice([3.0, 7.0, 25.0, 100.0]))
        lv = [(-1, -1), (0, 0), (2, -1), (0, 3), (1, 2), (-1, 0)][int(r.integers(0, 6))]
        assert orc.features_in_area(F, x, y, rad, *lv).tolist() == py_features_in_area(F, x, y, rad, *lv)


def test_bf_crosscheck_equals_cv2_bfmatcher():
    """SURVEY A-7: BFMatcher(NORM_HAMMING, crossCheck=True) = strict mutual NN, lowest index on ties."""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING, True)
    for seed in range(25):
        r = np.random.default_rng(seed)
        nq, nt = int(r.integers(1, 40)), int(r.integers(1, 40))
        q = scenario.degenerate_descriptors(nq, seed, 6); t = scenario.degenerate_descriptors(nt, seed + 1, 6)
        if seed % 3 == 0:
            q = r.integers(0, 256, (nq, 32), dtype=np.uint8); t = r.integers(0, 256, (nt, 32), dtype=np.uint8)
        ref = [(m.queryIdx, m.trainIdx, int(m.distance)) for m in bf.match(q, t)]
        xy = np.zeros((max(nq, nt), 2), np.float32)
        mq, mt, md, _ = orc.separate_pairs([(q, xy[:nq], t, xy[:nt])], np.eye(3), 0)[0]
        assert list(zip(mq.tolist(), mt.tolist(), md.tolist())) == ref


def test_invert3x3_equals_cv2():
    r = np.random.default_rng(3)
    for _ in range(300):
        H = (np.eye(3) + r.normal(size=(3, 3)) * 0.2).astype(np.float32)
        assert np.array_equal(cv2.invert(H)[1], orc.invert3x3(H))


def test_cv_gemm_3x3_is_sequential_float():
    """The projection Rcw*x3Dw+tcw in ORBmatcher.cc:1519 runs OpenCV's small-matrix path: float products
    summed left to right, then the addend — the semantics the oracle and the kernel both restate."""
    r = np.random.default_rng(4)
    for _ in range(500):
        R = r.normal(size=(3, 3)).astype(np.float32); x = (r.normal(size=(3, 1)) * 10).astype(np.float32)
        t = r.normal(size=(3, 1)).astype(np.float32)
        ref = cv2.gemm(R, x, 1.0, t, 1.0)
        mine = np.array([[f32(f32(f32(f32(R[i, 0] * x[0, 0]) + f32(R[i, 1] * x[1, 0])) + f32(R[i, 2] * x[2, 0])) + t[i, 0])]
                         for i in range(3)], np.float32)
        assert np.array_equal(ref, mine)


def py_search_frame(cur, last, lp, th, mono, check_ori):
    """Independent Python re-statement of SearchByProjection(Cur, Last, th, bMono), ORBmatcher.cc:1485-1627."""
    T = cur.tcw.reshape(3, 4); Tl = last.tcw.reshape(3, 4)
    fx, fy, cx, cy, bf, b = (f32(v) for v in cur.cam)
    twc = [f32(-1.0 * sum(float(T[k, r]) * float(T[k, 3]) for k in range(3))) for r in range(3)]
    tlc = [f32(f32(f32(f32(Tl[r, 0] * twc[0]) + f32(Tl[r, 1] * twc[1])) + f32(Tl[r, 2] * twc[2])) + Tl[r, 3]) for r in range(3)]
    fwd = tlc[2] > b and not mono; bwd = -tlc[2] > b and not mono
    assign = np.full(cur.n, -1, np.int32); locked = np.zeros(cur.n, np.uint8)
    hist = [[] for _ in range(30)]
    n = 0
    for i in range(last.n):
        if not lp["has_mp"][i] or lp["outlier"][i]:
            continue
        w = lp["world"][i]
        pc = [f32(f32(f32(f32(T[r, 0] * w[0]) + f32(T[r, 1] * w[1])) + f32(T[r, 2] * w[2])) + T[r, 3]) for r in range(3)]
        invz = f32(1.0 / float(pc[2]))
        if invz < 0:
            continue
        u = f32(f32(f32(fx * pc[0]) * invz) + cx); v = f32(f32(f32(fy * pc[1]) * invz) + cy)
        if u < cur.bounds[0] or u > cur.bounds[2] or v < cur.bounds[1] or v > cur.bounds[3]:
            continue
        oc = int(last.keys["octave"][i])
        radius = f32(f32(th) * cur.scale[oc])
        lv = (oc, -1) if fwd else ((0, oc) if bwd else (oc - 1, oc + 1))
        best, bi = 256, -1
        for i2 in py_features_in_area(cur, u, v, radius, *lv):
            if assign[i2] != -1 and locked[i2]:
                continue
            if cur.u_right is not None and cur.u_right[i2] > 0:
                ur = f32(u - f32(bf * invz))
                if abs(f32(ur - cur.u_right[i2])) > radius:
                    continue
            d = int(np.unpackbits(lp["desc"][i] ^ cur.desc[i2]).sum())
            if d < best:
                best, bi = d, i2
        if best <= 100:
            assign[bi] = i; locked[bi] = lp["obs_positive"][i]; n += 1
            if check_ori:
                rot = f32(last.keys_un["angle"][i] - cur.keys_un["angle"][bi])
                if rot < 0:
                    rot = f32(rot + f32(360))
                bn = int(math.floor(float(f32(rot * f32(1.0 / 30))) + 0.5))
                hist[0 if bn == 30 else bn].append(bi)
    if check_ori:
        sizes = [len(h) for h in hist]
        m1 = m2 = m3 = 0; i1 = i2_ = i3 = -1
        for i, s in enumerate(sizes):
            if s > m1:
                m3, m2, m1, i3, i2_, i1 = m2, m1, s, i2_, i1, i
            elif s > m2:
                m3, m2, i3, i2_ = m2, s, i2_, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < f32(0.1) * f32(m1):
            i2_ = i3 = -1
        elif m3 < f32(0.1) * f32(m1):
            i3 = -1
        for i, h in enumerate(hist):
            if i in (i1, i2_, i3):
                continue
            for idx in h:
                assign[idx] = -1; locked[idx] = 0; n -= 1
    return n, assign, locked


@pytest.mark.parametrize("stereo,th,tz", [(True, 7.0, 0.0), (False, 15.0, 0.0), (True, 7.0, 1.5), (True, 7.0, -1.5)])
def test_search_frame_oracle_vs_python(tum_pair, stereo, th, tz):
    p = tum_pair
    tcw = np.eye(4, dtype=np.float32)[:3].copy(); tcw[2, 3] = tz       # forward / backward motion branches
    cur = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"], stereo=stereo, seed=1, tcw=tcw)
    last = scenario.frame_view(p["k0"], p["d0"], p["scale"], p["W"], p["H"], stereo=stereo, seed=0)
    lp = scenario.last_points(p["k0"], p["d0"], (4, 1), seed=7)
    ref = py_search_frame(cur, last, lp, th, not stereo, True)
    got = orc.match_projection_frame(cur, last, lp, th, not stereo, True)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    assert got[0] > 100


def test_search_frame_tie_heavy(tum_pair):
    """Few distinct descriptors + unlocked (Observations()==0) claims: ties, overwrites and stale histogram
    entries all occur; oracle and the independent Python re-statement must still agree."""
    p = tum_pair
    d0 = scenario.degenerate_descriptors(len(p["k0"]), 1); d1 = scenario.degenerate_descriptors(len(p["k1"]), 2)
    cur = scenario.frame_view(p["k1"], d1, p["scale"], p["W"], p["H"], stereo=True, seed=1)
    last = scenario.frame_view(p["k0"], d0, p["scale"], p["W"], p["H"], stereo=True, seed=0)
    lp = scenario.last_points(p["k0"], d0, (4, 1), seed=9, p_obs=0.5, noise_bits=0)
    ref = py_search_frame(cur, last, lp, 15.0, False, True)
    got = orc.match_projection_frame(cur, last, lp, 15.0, False, True)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def test_first_separate_erase_skip_bug():
    """Appendix B-4, literally: boxes are erased while iterating without --i and hasKpts is not erased in step,
    so with boxes [A(kp), B(empty), C(empty), D(kp)] the loop erases B, then tests hasKpts[2] (C's flag) against
    the box that slid into slot 2 (D) and erases D: the empty C survives and the occupied D is dropped."""
    keys = np.zeros(3, orc.KP_DTYPE); keys["class_id"] = -1
    keys["x"] = [10, 50, 200]; keys["y"] = [10, 50, 200]
    boxes = [[0, 0, 20, 20], [300, 300, 5, 5], [310, 310, 5, 5], [190, 190, 20, 20]]
    r = orc.first_separate(keys, boxes, [0, 1, 2, 3])
    assert r["n_dyn"] == 2 and r["order"].tolist() == [1, 0, 2]
    assert r["boxes"].tolist() == [[0, 0, 20, 20], [310, 310, 5, 5]] and r["box_idx"].tolist() == [0, 2]
    assert r["class_id"].tolist() == [0, -1, 2]
    assert r["dyn"] == [(0, 0), (1, 2)]              # index 3 - (two empty boxes before it) = 1


def test_box_track_assigns_ids_and_carries_boxes():
    b, idx, omit, vel = orc.box_track([[10, 10, 50, 50]], [], [], [], [], 640, 480)
    assert idx.tolist() == [0]
    last = [[12, 11, 50, 50], [300, 200, 40, 40]]
    b, idx, omit, vel = orc.box_track([[14, 12, 50, 50]], last, [0, 1], [0, 0], [[0, 0], [5, 0]], 640, 480)
    assert idx.tolist() == [0, 1] and omit.tolist() == [0, 1]          # second box carried over with its velocity
    assert np.allclose(b[1], [305, 200, 40, 40]) and np.allclose(vel[0], [2, 1])


def py_three_maxima(sizes):
    m1 = m2 = m3 = 0; i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    if m2 < f32(0.1) * f32(m1):
        i2 = i3 = -1
    elif m3 < f32(0.1) * f32(m1):
        i3 = -1
    return i1, i2, i3


def py_rot_bin(a, b):
    rot = f32(a - b)
    if rot < 0:
        rot = f32(rot + f32(360))
    bn = int(math.floor(float(f32(rot * f32(1.0 / 30))) + 0.5))
    return 0 if bn == 30 else bn


def py_search_bow_kf(kf1, v1, nodes1, kf2, v2, nodes2, ratio, check_ori):
    """Independent Python re-statement of SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12), ORBmatcher.cc:679-812:
    std::map iteration = ascending node id, in-node order = feature index order."""
    def fmap(nodes):
        m = {}
        for i, nd in enumerate(nodes):
            m.setdefault(int(nd), []).append(i)
        return m
    f1, f2 = fmap(nodes1), fmap(nodes2)
    m12 = np.full(kf1.n, -1, np.int32); matched2 = np.zeros(kf2.n, bool)
    hist = [[] for _ in range(30)]
    n = 0
    for node in sorted(set(f1) & set(f2)):
        for i1 in f1[node]:
            if not v1[i1]:
                continue
            b1, bi, b2 = 256, -1, 256
            for i2 in f2[node]:
                if matched2[i2] or not v2[i2]:
                    continue
                d = int(np.unpackbits(kf1.desc[i1] ^ kf2.desc[i2]).sum())
                if d < b1:
                    b2, b1, bi = b1, d, i2
                elif d < b2:
                    b2 = d
            if b1 < 50 and f32(b1) < f32(ratio) * f32(b2):
                m12[i1] = bi; matched2[bi] = True; n += 1
                if check_ori:
                    hist[py_rot_bin(kf1.keys_un["angle"][i1], kf2.keys_un["angle"][bi])].append(i1)
    if check_ori:
        keep = py_three_maxima([len(h) for h in hist])
        for i, h in enumerate(hist):
            if i in keep:
                continue
            for idx in h:
                m12[idx] = -1; n -= 1
    return n, m12


@pytest.mark.parametrize("check", [True, False])
def test_search_bow_keyframes_oracle_vs_python(tum_pair, check):
    p = tum_pair
    kf1 = scenario.frame_view(p["k0"], p["d0"], p["scale"], p["W"], p["H"])
    kf2 = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"])
    n1, n2 = scenario.bow_nodes(p["d0"]), scenario.bow_nodes(p["d1"])
    r = np.random.default_rng(5)
    v1 = (r.random(kf1.n) < 0.8).astype(np.uint8); v2 = (r.random(kf2.n) < 0.7).astype(np.uint8)
    got = orc.match_bow_kf(kf1, v1, pysdyn.FeatureVector(n1), kf2, v2, pysdyn.FeatureVector(n2), 0.75, check)
    ref = py_search_bow_kf(kf1, v1, n1, kf2, v2, n2, 0.75, check)
    assert got[0] == ref[0] and got[0] > 30 and np.array_equal(got[1], ref[1])
    # tie-heavy
    d0 = scenario.degenerate_descriptors(kf1.n, 21, 30); d1 = scenario.degenerate_descriptors(kf2.n, 22, 30)
    kf1 = scenario.frame_view(p["k0"], d0, p["scale"], p["W"], p["H"]); kf2 = scenario.frame_view(p["k1"], d1, p["scale"], p["W"], p["H"])
    n1, n2 = scenario.bow_nodes(d0, 3), scenario.bow_nodes(d1, 3)
    got = orc.match_bow_kf(kf1, v1, pysdyn.FeatureVector(n1), kf2, v2, pysdyn.FeatureVector(n2), 0.95, check)
    ref = py_search_bow_kf(kf1, v1, n1, kf2, v2, n2, 0.95, check)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1])


def py_predict_scale(raw, dist, log_sf, nlevels):
    ratio = f32(f32(raw) / f32(dist))
    n = int(math.ceil(float(f32(f32(math.log(float(ratio))) / f32(log_sf)))))     # logf = rounded double log
    return 0 if n < 0 else (nlevels - 1 if n >= nlevels else n)


def py_search_pose(target, pts, R, tcw, ow, th, maxd, variant, check_ori, log_sf, nlevels, occ):
    """Independent Python re-statement of the two pose-projection searches (ORBmatcher.cc:1629-1756 / :290-403);
    Rcw*x+tcw, cv::norm and Mat::dot run through cv2 itself."""
    fx, fy, cx, cy = (f32(v) for v in target.cam[:4])
    assign = np.asarray(occ, np.int32).copy()
    hist = [[] for _ in range(30)]
    n = 0
    R = np.ascontiguousarray(R, np.float32); t = np.asarray(tcw, np.float32).reshape(3, 1); o = np.asarray(ow, np.float32).reshape(3, 1)
    for i, p in enumerate(pts):
        if not p["valid"]:
            continue
        xw = p["world"].reshape(3, 1).astype(np.float32)
        pc = cv2.gemm(R, xw, 1.0, t, 1.0).reshape(3)
        if variant == pysdyn.PROJ_FRAME_KEYFRAME:
            invz = f32(1.0 / float(pc[2]))
            u = f32(f32(f32(fx * pc[0]) * invz) + cx); v = f32(f32(f32(fy * pc[1]) * invz) + cy)
            if u < target.bounds[0] or u > target.bounds[2] or v < target.bounds[1] or v > target.bounds[3]:
                continue
        else:
            if pc[2] < 0.0:
                continue
            invz = f32(f32(1) / pc[2])
            u = f32(f32(fx * f32(pc[0] * invz)) + cx); v = f32(f32(fy * f32(pc[1] * invz)) + cy)
            if not (u >= target.bounds[0] and u < target.bounds[2] and v >= target.bounds[1] and v < target.bounds[3]):
                continue
        po = cv2.subtract(xw, o)
        dist = f32(cv2.norm(po))
        if dist < p["min_distance"] or dist > p["max_distance"]:
            continue
        if variant == pysdyn.PROJ_KEYFRAME_SIM3 and float(po.reshape(3).astype(np.float64) @ p["normal"].astype(np.float64)) < 0.5 * float(dist):
            continue
        lvl = py_predict_scale(p["max_distance_raw"], dist, log_sf, nlevels)
        radius = f32(f32(th) * target.scale[lvl])
        if variant == pysdyn.PROJ_FRAME_KEYFRAME:
            cand = py_features_in_area(target, u, v, radius, lvl - 1, lvl + 1)
        else:
            cand = [j for j in py_features_in_area(target, u, v, radius, -1, -1)]
        best, bi = 256, -1
        for j in cand:
            if assign[j] != -1:
                continue
            if variant == pysdyn.PROJ_KEYFRAME_SIM3 and not (lvl - 1 <= target.keys_un["octave"][j] <= lvl):
                continue
            d = int(np.unpackbits(p["desc"] ^ target.desc[j]).sum())
            if d < best:
                best, bi = d, j
        if best <= maxd:
            assign[bi] = i; n += 1
            if check_ori and variant == pysdyn.PROJ_FRAME_KEYFRAME:
                hist[py_rot_bin(p["angle"], target.keys_un["angle"][bi])].append(bi)
    if check_ori and variant == pysdyn.PROJ_FRAME_KEYFRAME:
        keep = py_three_maxima([len(h) for h in hist])
        for i, h in enumerate(hist):
            if i in keep:
                continue
            for idx in h:
                assign[idx] = -1; n -= 1
    return n, assign


@pytest.mark.parametrize("variant,th,maxd", [(0, 10.0, 100), (1, 10.0, 50)])
def test_search_pose_oracle_vs_python(tum_pair, variant, th, maxd):
    p = tum_pair
    target = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"])
    R, tcw, ow = scenario.pose_small(seed=3)
    pts = scenario.proj_points(p["k1"], p["d1"], p["scale"], R, tcw, ow, seed=11)
    log_sf = np.log(np.float32(1.2))
    prm = pysdyn.proj_params(R, tcw, ow, th, maxd, variant, True, log_sf, 8)
    occ = np.where(np.random.default_rng(4).random(target.n) < 0.1, -2, -1).astype(np.int32)
    got = orc.match_projection_pose(target, pts, prm, occ)
    ref = py_search_pose(target, pts, R, tcw, ow, th, maxd, variant, True, log_sf, 8, occ)
    assert got[0] == ref[0] and got[0] > 100 and np.array_equal(got[1], ref[1])
    # every gate of the reference must actually fire in this scenario
    assert (~pts["valid"].astype(bool)).any() and (got[1] == -2).any()


def py_fuse_search(KF, inv_s2, pts, R, tcw, ow, th, log_sf, nlevels):
    """Independent Python re-statement of the search part of Fuse(pKF, vpMapPoints, th), ORBmatcher.cc:982-1100."""
    fx, fy, cx, cy, bf = (f32(v) for v in KF.cam[:5])
    R = np.ascontiguousarray(R, np.float32); t = np.asarray(tcw, np.float32).reshape(3, 1); o = np.asarray(ow, np.float32).reshape(3, 1)
    bi = np.full(len(pts), -1, np.int32); bd = np.full(len(pts), 256, np.int32)
    for i, p in enumerate(pts):
        if not p["valid"]:
            continue
        xw = p["world"].reshape(3, 1).astype(np.float32)
        pc = cv2.gemm(R, xw, 1.0, t, 1.0).reshape(3)
        if pc[2] < 0:
            continue
        invz = f32(f32(1) / pc[2])
        u = f32(f32(fx * f32(pc[0] * invz)) + cx); v = f32(f32(fy * f32(pc[1] * invz)) + cy)
        if not (u >= KF.bounds[0] and u < KF.bounds[2] and v >= KF.bounds[1] and v < KF.bounds[3]):
            continue
        ur = f32(u - f32(bf * invz))
        po = cv2.subtract(xw, o)
        dist = f32(cv2.norm(po))
        if dist < p["min_distance"] or dist > p["max_distance"]:
            continue
        if float(po.reshape(3).astype(np.float64) @ p["normal"].astype(np.float64)) < 0.5 * float(dist):
            continue
        lvl = py_predict_scale(p["max_distance_raw"], dist, log_sf, nlevels)
        radius = f32(f32(th) * KF.scale[lvl])
        best, b = 256, -1
        for j in py_features_in_area(KF, u, v, radius, -1, -1):
            kp = KF.keys_un[j]
            if not (lvl - 1 <= kp["octave"] <= lvl):
                continue
            ex = f32(u - kp["x"]); ey = f32(v - kp["y"])
            if KF.u_right is not None and KF.u_right[j] >= 0:
                er = f32(ur - KF.u_right[j])
                e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                if float(f32(e2 * inv_s2[kp["octave"]])) > 7.8:
                    continue
            else:
                e2 = f32(f32(ex * ex) + f32(ey * ey))
                if float(f32(e2 * inv_s2[kp["octave"]])) > 5.99:
                    continue
            d = int(np.unpackbits(p["desc"] ^ KF.desc[j]).sum())
            if d < best:
                best, b = d, j
        bi[i] = b; bd[i] = best
    return bi, bd


@pytest.mark.parametrize("stereo", [True, False])
def test_fuse_search_oracle_vs_python(tum_pair, stereo):
    p = tum_pair
    KF = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"], stereo=stereo, seed=1)
    R, tcw, ow = scenario.pose_small(seed=5)
    pts = scenario.proj_points(p["k1"], p["d1"], p["scale"], R, tcw, ow, seed=13, jitter=1.5)[:400]
    inv_s2 = (1.0 / (p["scale"] * p["scale"])).astype(np.float32)
    log_sf = np.log(np.float32(1.2))
    got = orc.fuse_search(0, KF, inv_s2, pts, R, tcw, ow, 3.0, log_sf, 8)
    ref = py_fuse_search(KF, inv_s2, pts, R, tcw, ow, 3.0, log_sf, 8)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and (got[0] >= 0).sum() > 50


def py_search_triangulation(kf1, h1, nodes1, kf2, h2, nodes2, F12, epi, sig2, only_stereo, check_ori):
    """Independent Python re-statement of SearchForTriangulation + CheckDistEpipolarLine, ORBmatcher.cc:814-980, 140-157."""
    def fmap(nodes):
        m = {}
        for i, nd in enumerate(nodes):
            m.setdefault(int(nd), []).append(i)
        return m
    f1, f2 = fmap(nodes1), fmap(nodes2)
    F = np.asarray(F12, np.float32)
    m12 = np.full(kf1.n, -1, np.int32); hist = [[] for _ in range(30)]; n = 0
    st1 = (lambda i: kf1.u_right is not None and kf1.u_right[i] >= 0); st2 = (lambda i: kf2.u_right is not None and kf2.u_right[i] >= 0)
    for node in sorted(set(f1) & set(f2)):
        for i1 in f1[node]:
            if h1[i1] or (only_stereo and not st1(i1)):
                continue
            k1 = kf1.keys_un[i1]
            a = f32(f32(f32(k1["x"] * F[0, 0]) + f32(k1["y"] * F[1, 0])) + F[2, 0])
            b = f32(f32(f32(k1["x"] * F[0, 1]) + f32(k1["y"] * F[1, 1])) + F[2, 1])
            c = f32(f32(f32(k1["x"] * F[0, 2]) + f32(k1["y"] * F[1, 2])) + F[2, 2])
            best, bi = 50, -1
            for i2 in f2[node]:
                if h2[i2] or (only_stereo and not st2(i2)):
                    continue
                d = int(np.unpackbits(kf1.desc[i1] ^ kf2.desc[i2]).sum())
                if d > 50 or d > best:
                    continue
                k2 = kf2.keys_un[i2]
                if not st1(i1) and not st2(i2):
                    ex = f32(f32(epi[0]) - k2["x"]); ey = f32(f32(epi[1]) - k2["y"])
                    if f32(f32(ex * ex) + f32(ey * ey)) < f32(f32(100) * kf2.scale[k2["octave"]]):
                        continue
                num = f32(f32(f32(a * k2["x"]) + f32(b * k2["y"])) + c)
                den = f32(f32(a * a) + f32(b * b))
                if den == 0:
                    continue
                dsqr = f32(f32(num * num) / den)
                if float(dsqr) < 3.84 * float(sig2[k2["octave"]]):
                    best, bi = d, i2
            if bi >= 0:
                m12[i1] = bi; n += 1
                if check_ori:
                    hist[py_rot_bin(k1["angle"], kf2.keys_un["angle"][bi])].append(i1)
    if check_ori:
        keep = py_three_maxima([len(h) for h in hist])
        for i, h in enumerate(hist):
            if i not in keep:
                for idx in h:
                    m12[idx] = -1; n -= 1
    return n, m12


@pytest.mark.parametrize("stereo,only_stereo", [(False, False), (True, True)])
def test_search_triangulation_oracle_vs_python(tum_pair, stereo, only_stereo):
    p = tum_pair
    kf1 = scenario.frame_view(p["k0"], p["d0"], p["scale"], p["W"], p["H"], stereo=stereo, seed=0)
    kf2 = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"], stereo=stereo, seed=1)
    n1, n2 = scenario.bow_nodes(p["d0"], 4), scenario.bow_nodes(p["d1"], 4)
    r = np.random.default_rng(8)
    h1 = (r.random(kf1.n) < 0.3).astype(np.uint8); h2 = (r.random(kf2.n) < 0.3).astype(np.uint8)
    F12 = np.array([[0, 0, 1], [0, 0, -4], [-1, 4, 0]], np.float32) * np.float32(0.01)
    sig2 = (p["scale"] * p["scale"]).astype(np.float32)
    epi = (p["W"] * 0.4, p["H"] * 0.5)
    prm = pysdyn.tri_params(F12, epi, only_stereo, True, sig2)
    got = orc.match_triangulation(kf1, h1, pysdyn.FeatureVector(n1), kf2, h2, pysdyn.FeatureVector(n2), prm)
    ref = py_search_triangulation(kf1, h1, n1, kf2, h2, n2, F12, epi, sig2, only_stereo, True)
    assert got[0] == ref[0] and got[0] > 50 and np.array_equal(got[1], ref[1])


def py_search_map(F, mps, th, ratio, assign0, locked0, base=0):
    """Independent Python re-statement of SearchByProjection(Frame&, vector<MapPoint*>&, th), ORBmatcher.cc:45-129."""
    assign = np.asarray(assign0, np.int32).copy(); locked = np.asarray(locked0, np.uint8).copy()
    n = 0
    for i, mp in enumerate(mps):
        if not mp["track_in_view"] or mp["bad"]:
            continue
        lvl = int(mp["level"])
        r = f32(2.5) if float(mp["view_cos"]) > 0.998 else f32(4.0)
        if th != 1.0:
            r = f32(r * f32(th))
        rad = f32(r * F.scale[lvl])
        b1, l1, b2, l2, bi = 256, -1, 256, -1, -1
        for j in py_features_in_area(F, mp["proj_x"], mp["proj_y"], rad, lvl - 1, lvl):
            if assign[j] != -1 and locked[j]:
                continue
            if F.u_right is not None and F.u_right[j] > 0 and abs(f32(mp["proj_xr"] - F.u_right[j])) > rad:
                continue
            d = int(np.unpackbits(mp["desc"] ^ F.desc[j]).sum())
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, int(F.keys_un["octave"][j]), j
            elif d < b2:
                l2, b2 = int(F.keys_un["octave"][j]), d
        if b1 <= 100:
            if l1 == l2 and b1 > f32(ratio) * f32(b2):
                continue
            assign[bi] = base + i; locked[bi] = mp["obs_positive"]; n += 1
    return n, assign, locked


@pytest.mark.parametrize("stereo,th", [(True, 3.0), (False, 1.0)])
def test_search_map_oracle_vs_python(tum_pair, stereo, th):
    p = tum_pair
    F = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"], stereo=stereo, seed=1)
    mps = scenario.map_queries(p["k1"], p["d1"], 8, seed=5, count=900)
    r = np.random.default_rng(3)
    a0 = np.where(r.random(F.n) < 0.15, -2, -1).astype(np.int32); l0 = ((a0 != -1) & (r.random(F.n) < 0.7)).astype(np.uint8)
    got = orc.match_projection_map(F, mps, th, 0.8, a0, l0)
    ref = py_search_map(F, mps, th, 0.8, a0, l0)
    assert got[0] == ref[0] and got[0] > 100 and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def py_search_init(F1, F2, prev, window, ratio, check_ori):
    """Independent Python re-statement of SearchForInitialization, ORBmatcher.cc:562-677."""
    prev = np.asarray(prev, np.float32).copy()
    m12 = np.full(F1.n, -1, np.int32); m21 = np.full(F2.n, -1, np.int32); mdist = np.full(F2.n, 2 ** 31 - 1, np.int64)
    hist = [[] for _ in range(30)]; n = 0
    for i1 in range(F1.n):
        lvl = int(F1.keys_un["octave"][i1])
        if lvl > 0:
            continue
        b1, b2, bi = 2 ** 31 - 1, 2 ** 31 - 1, -1
        for i2 in py_features_in_area(F2, prev[i1, 0], prev[i1, 1], f32(window), lvl, lvl):
            d = int(np.unpackbits(F1.desc[i1] ^ F2.desc[i2]).sum())
            if mdist[i2] <= d:
                continue
            if d < b1:
                b2, b1, bi = b1, d, i2
            elif d < b2:
                b2 = d
        if b1 <= 50 and f32(b1) < f32(f32(b2) * f32(ratio)):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1; n -= 1
            m12[i1] = bi; m21[bi] = i1; mdist[bi] = b1; n += 1
            if check_ori:
                hist[py_rot_bin(F1.keys_un["angle"][i1], F2.keys_un["angle"][bi])].append(i1)
    if check_ori:
        keep = py_three_maxima([len(h) for h in hist])
        for i, h in enumerate(hist):
            if i not in keep:
                for idx in h:
                    if m12[idx] >= 0:
                        m12[idx] = -1; n -= 1
    for i1 in range(F1.n):
        if m12[i1] >= 0:
            prev[i1, 0] = F2.keys_un["x"][m12[i1]]; prev[i1, 1] = F2.keys_un["y"][m12[i1]]
    return n, m12, prev


@pytest.mark.parametrize("check", [True, False])
def test_search_init_oracle_vs_python(tum_pair, check):
    p = tum_pair
    F1 = scenario.frame_view(p["k0"], p["d0"], p["scale"], p["W"], p["H"])
    F2 = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"])
    prev = np.stack([p["k0"]["x"], p["k0"]["y"]], 1).astype(np.float32)
    got = orc.match_init(F1, F2, prev, 40, 0.9, check)
    ref = py_search_init(F1, F2, prev, 40, 0.9, check)
    assert got[0] == ref[0] and got[0] > 50 and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])


def py_search_bow(KF, valid, nodes1, F, nodes2, ratio, check_ori):
    """Independent Python re-statement of SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches), ORBmatcher.cc:159-288."""
    def fmap(nodes):
        m = {}
        for i, nd in enumerate(nodes):
            m.setdefault(int(nd), []).append(i)
        return m
    f1, f2 = fmap(nodes1), fmap(nodes2)
    assign = np.full(F.n, -1, np.int32); hist = [[] for _ in range(30)]; n = 0
    for node in sorted(set(f1) & set(f2)):
        for ik in f1[node]:
            if not valid[ik]:
                continue
            b1, bi, b2 = 256, -1, 256
            for jf in f2[node]:
                if assign[jf] != -1:
                    continue
                d = int(np.unpackbits(KF.desc[ik] ^ F.desc[jf]).sum())
                if d < b1:
                    b2, b1, bi = b1, d, jf
                elif d < b2:
                    b2 = d
            if b1 <= 50 and f32(b1) < f32(ratio) * f32(b2):
                assign[bi] = ik; n += 1
                if check_ori:
                    hist[py_rot_bin(KF.keys_un["angle"][ik], F.keys["angle"][bi])].append(bi)
    if check_ori:
        keep = py_three_maxima([len(h) for h in hist])
        for i, h in enumerate(hist):
            if i not in keep:
                for idx in h:
                    assign[idx] = -1; n -= 1
    return n, assign


@pytest.mark.parametrize("check", [True, False])
def test_search_bow_oracle_vs_python(tum_pair, check):
    p = tum_pair
    KF = scenario.frame_view(p["k0"], p["d0"], p["scale"], p["W"], p["H"])
    F = scenario.frame_view(p["k1"], p["d1"], p["scale"], p["W"], p["H"])
    n1, n2 = scenario.bow_nodes(p["d0"]), scenario.bow_nodes(p["d1"])
    valid = (np.random.default_rng(2).random(KF.n) < 0.8).astype(np.uint8)
    got = orc.match_bow(KF, valid, pysdyn.FeatureVector(n1), F, pysdyn.FeatureVector(n2), 0.7, check)
    ref = py_search_bow(KF, valid, n1, F, n2, 0.7, check)
    assert got[0] == ref[0] and got[0] > 30 and np.array_equal(got[1], ref[1])
