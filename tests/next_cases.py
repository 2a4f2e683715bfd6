"""Seeded cases of the SURVEY §8(f) rows (stereo association, remaining ORBmatcher overloads, ComputeBoW, undistortion)
evaluated through a backend — "oracle" (CPU restatement), "cuda" (the product through the C ABI) or "twin" (independent
Python / cv2 re-statements, where one exists) — and reduced to SHA-256 digests.  tools/gen_golden.py writes the digests
of the twins / oracle to tests/golden/next_rows.json; tests/test_golden.py checks that the oracle (CPU) and the CUDA
path (GPU) reproduce them.  TEST INFRASTRUCTURE."""
import numpy as np

import common
import orc
import pysdyn
import scenario

CAM = scenario.KITTI_CAM
MB, MBF = CAM["bf"] / CAM["fx"], CAM["bf"]
TUM1 = (517.306408, 516.469215, 318.643040, 255.313989, [0.262383, -0.953104, -0.005358, 0.002628, 1.163314])


def rt(R, t):
    return np.concatenate([np.asarray(R, np.float32), np.asarray(t, np.float32).reshape(3, 1)], 1)


class Inputs:
    """Everything is derived from the oracle's extraction of two seeded frames (identical on every backend)."""

    def __init__(self, cfg="tum"):
        self.cfg = cfg
        self.W, self.H, _, self.nf, self.ini, self.mn = common.CONFIGS[cfg]
        E = orc.Extractor(self.nf, 1.2, 8, self.ini, self.mn)
        self.k0, self.d0 = E(common.frame(cfg, 0)); self.k1, self.d1 = E(common.frame(cfg, 1, ox=4, oy=1, t=1))
        self.scale = E.scale
        self.log_sf = np.log(np.float32(1.2))
        self.R, self.tcw, self.ow = scenario.pose_small(seed=3)
        self.pts = scenario.proj_points(self.k1, self.d1, self.scale, self.R, self.tcw, self.ow, seed=11, jitter=1.5)
        self.inv_s2 = (1.0 / (self.scale * self.scale)).astype(np.float32)
        self.sig2 = (self.scale * self.scale).astype(np.float32)

    def view(self, which, stereo=False):
        k, d, seed = (self.k0, self.d0, 0) if which == 0 else (self.k1, self.d1, 1)
        return scenario.frame_view(k, d, self.scale, self.W, self.H, stereo=stereo, seed=seed)


def matcher_cases(inp, backend, ctx=None):
    """-> {name: digest}.  backend "oracle" or "cuda" (ctx = pysdyn.Extractor used as the matcher context)."""
    out = {}
    m = pysdyn.Matcher(ctx, 0.75, True) if backend == "cuda" else None
    kf1, kf2 = inp.view(0), inp.view(1)
    fa = pysdyn.FeatureVector(scenario.bow_nodes(inp.d0, 4)); fb = pysdyn.FeatureVector(scenario.bow_nodes(inp.d1, 4))
    r = np.random.default_rng(5)
    v1 = (r.random(kf1.n) < 0.8).astype(np.uint8); v2 = (r.random(kf2.n) < 0.7).astype(np.uint8)
    n, m12 = m.SearchByBoWKF(kf1, v1, fa, kf2, v2, fb) if m else orc.match_bow_kf(kf1, v1, fa, kf2, v2, fb, 0.75, True)
    out["bow_kf"] = [int(n), common.sha(m12)]
    target = inp.view(1)
    occ = np.where(np.random.default_rng(4).random(target.n) < 0.1, -2, -1).astype(np.int32)
    for variant, th, maxd in ((0, 10.0, 100), (1, 10.0, 50)):
        prm = pysdyn.proj_params(inp.R, inp.tcw, inp.ow, th, maxd, variant, True, inp.log_sf, 8)
        n, a = m.SearchByProjectionPose(target, inp.pts, prm, occ) if m else orc.match_projection_pose(target, inp.pts, prm, occ)
        out["pose%d" % variant] = [int(n), common.sha(a)]
    for which, stereo in ((0, True), (1, False)):
        tgt = inp.view(1, stereo)
        if m:
            prm = pysdyn.best_params(rt(inp.R, inp.tcw), 3.0, inp.log_sf, 8, ow=inp.ow, invz_double=(which == 1), check_normal=True,
                                     chi2_gate=(which == 0), bf=tgt.cam[4], inv_level_sigma2=inp.inv_s2)
            bi, bd = m.ProjectionBest(tgt, inp.pts, prm)
        else:
            bi, bd = orc.fuse_search(which, tgt, inp.inv_s2, inp.pts, inp.R, inp.tcw, inp.ow, 3.0, inp.log_sf, 8)
        out["fuse%d" % which] = [common.sha(bi), common.sha(bd)]
    eye = np.eye(3, dtype=np.float32); zero = np.zeros(3, np.float32)
    R12, t12, _ = scenario.pose_small(seed=6, angle_deg=0.8, t=(0.03, 0.01, -0.05))
    s12 = np.float32(1.03)
    sR12 = (s12 * R12).astype(np.float32); sR21 = ((np.float32(1.0) / s12) * R12.T).astype(np.float32)
    t21 = (-(sR21 @ t12)).astype(np.float32)
    p1 = scenario.proj_points(inp.k0, inp.d0, inp.scale, eye, zero, zero, seed=21, p_valid=0.8)
    p2 = scenario.proj_points(inp.k1, inp.d1, inp.scale, eye, zero, zero, seed=22, p_valid=0.8)
    args = (kf1, kf2, p1, p2, rt(eye, zero), rt(eye, zero), rt(sR12, t12), rt(sR21, t21), 7.5, inp.log_sf, 8)
    n, a = m.SearchBySim3(*args) if m else orc.search_by_sim3(*args)
    out["sim3"] = [int(n), common.sha(a)]
    h1 = (r.random(kf1.n) < 0.3).astype(np.uint8); h2 = (r.random(kf2.n) < 0.3).astype(np.uint8)
    F12 = np.array([[0, 0, 1], [0, 0, -4], [-1, 4, 0]], np.float32) * np.float32(0.01)
    prm = pysdyn.tri_params(F12, (inp.W * 0.4, inp.H * 0.5), False, True, inp.sig2)
    n, a = m.SearchForTriangulation(kf1, h1, fa, kf2, h2, fb, prm) if m else orc.match_triangulation(kf1, h1, fa, kf2, h2, fb, prm)
    out["triangulation"] = [int(n), common.sha(a)]
    return out


def stereo_case(backend, cfg="small", idx=0, disp=(5, 11, 23)):
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    left, right = scenario.stereo_pair(cfg, idx, disp)
    if backend == "cuda":
        L = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H); R = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H)
        kl, _ = L(left); R(right)
        ur, dp, kept = pysdyn.stereo_match(L, R, 1, MB, MBF)
        ur, dp = ur[0, :len(kl)], dp[0, :len(kl)]
        L.close(); R.close()
    else:
        EL, ER = orc.Extractor(nf, 1.2, 8, ini, mn), orc.Extractor(nf, 1.2, 8, ini, mn)
        kl, dl = EL(left); kr, dr = ER(right)
        if backend == "twin":
            import stereo_twin
            pl = [EL.level(l)[19:-19, 19:-19] for l in range(8)]; pr = [ER.level(l)[19:-19, 19:-19] for l in range(8)]
            ur, dp = stereo_twin.compute_stereo_matches(kl, dl, kr, dr, pl, pr, EL.scale, EL.inv_scale, MB, MBF)
        else:
            ur, dp, _ = orc.stereo_matches(EL, ER, kl, dl, kr, dr, MB, MBF)
    return [int((ur >= 0).sum()), common.sha(ur), common.sha(dp)]


def undistort_case(backend):
    fx, fy, cx, cy, d = TUM1
    f = lambda v: float(np.float32(v))
    r = np.random.default_rng(2)
    pts = np.concatenate([r.uniform(-20, 660, (5000, 1)), r.uniform(-20, 500, (5000, 1))], 1).astype(np.float32)
    if backend == "twin":
        import cv2
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32); D = np.array(d, np.float32).reshape(-1, 1)
        out = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    elif backend == "cuda":
        ex = pysdyn.Extractor(500, 1.2, 8, 20, 7, max_width=640, max_height=480)
        ex.set_camera(f(fx), f(fy), f(cx), f(cy), np.array(d, np.float32))
        out = ex.undistort_points(pts)
        ex.close()
    else:
        out = orc.undistort_points(pts, f(fx), f(fy), f(cx), f(cy), np.array(d, np.float32))
    return [common.sha(np.ascontiguousarray(out, np.float32))]


def bow_case(backend, ctx=None):
    inp_desc = orc.Extractor(500, 1.2, 8, 20, 7)(common.frame("small", 0))[1]
    parent, leaf, desc, weight = scenario.synthetic_vocabulary(10, 3, seed=9, ragged=0.1, base_desc=inp_desc[0])
    if backend == "cuda":
        voc = pysdyn.Vocabulary(parent, leaf, desc, weight, 10, 3)
        o = pysdyn.bow_transform(ctx, voc, inp_desc, 2)
        voc.close()
    else:
        o = orc.Vocabulary(parent, leaf, desc, weight, 10, 3).transform(inp_desc, 2)
    return [common.sha(o["word"]), common.sha(o["node"]), common.sha(o["bow_ids"]), common.sha(o["bow_values"]),
            common.sha(o["fv_nodes"]), common.sha(o["fv_offset"]), common.sha(o["fv_index"])]
