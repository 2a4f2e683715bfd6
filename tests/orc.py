"""ctypes binding for the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by the product package.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(_ORACLE_DIR, "liborc.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build():
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])


def _load():
    if not os.path.exists(_LIB_PATH):
        build()
    return C.CDLL(_LIB_PATH)


lib = _load()
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)


def _p(a, t):
    return a.ctypes.data_as(t)


lib.orc_extractor_create.restype = C.c_void_p
lib.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
lib.orc_extractor_destroy.argtypes = [C.c_void_p]
lib.orc_extractor_run.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, C.c_int]
lib.orc_extractor_tables.argtypes = [C.c_void_p] + [_f32p] * 4 + [_i32p] * 2
lib.orc_extractor_level_dims.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
lib.orc_extractor_level_copy.argtypes = [C.c_void_p, C.c_int, _u8p]
lib.orc_extractor_candidates.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int]
lib.orc_extractor_level_count.argtypes = [C.c_void_p, C.c_int]
lib.orc_ic_angle.restype = C.c_float
lib.orc_ic_angle.argtypes = [_u8p, C.c_int]
lib.orc_orb_descriptor.argtypes = [C.c_float, _u8p, C.c_int, _u8p]


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib.orc_resize_linear_u8(_p(src, _u8p), src.shape[1], src.shape[0], src.strides[0], _p(dst, _u8p), dw, dh, dw)
    return dst


def border_reflect101(src, b):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * b, w + 2 * b), np.uint8)
    lib.orc_border_reflect101(_p(src, _u8p), w, h, src.strides[0], _p(dst, _u8p), b, w + 2 * b)
    return dst


def gaussian_blur7(src):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h, w), np.uint8)
    lib.orc_gaussian_blur7(_p(src, _u8p), w, h, src.strides[0], _p(dst, _u8p), w)
    return dst


def fast_nms(img, th):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((w * h, 3), np.int32)
    n = lib.orc_fast_nms(_p(img, _u8p), w, h, img.strides[0], th, _p(out, _i32p), w * h)
    return out[:n].copy()


def fast_score_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w), np.int16)
    lib.orc_fast_score_map(_p(img, _u8p), w, h, img.strides[0], out.ctypes.data_as(C.POINTER(C.c_int16)))
    return out


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib.orc_fast_atan2(_p(y, _f32p), _p(x, _f32p), _p(out, _f32p), y.size)
    return out


def ic_angle(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    ptr = C.cast(img.ctypes.data + y * img.strides[0] + x, _u8p)
    return np.float32(lib.orc_ic_angle(ptr, img.strides[0]))


def orb_descriptor(img, x, y, angle):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(32, np.uint8)
    ptr = C.cast(img.ctypes.data + y * img.strides[0] + x, _u8p)
    lib.orc_orb_descriptor(C.c_float(angle), ptr, img.strides[0], _p(out, _u8p))
    return out


def undistort_points(xy, fx, fy, cx, cy, dist):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2); dist = np.ascontiguousarray(dist, np.float32)
    out = np.empty_like(xy)
    lib.orc_undistort_points.argtypes = [_f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _f32p, C.c_int, _f32p]
    lib.orc_undistort_points(_p(xy, _f32p), len(xy), fx, fy, cx, cy, _p(dist, _f32p), len(dist), _p(out, _f32p))
    return out


class Extractor:
    """Oracle ORBextractor (reference: src/ORBextractor.cc)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = lib.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        sc, inv, sg, isg = (np.empty(nlevels, np.float32) for _ in range(4))
        q = np.empty(nlevels, np.int32)
        um = np.empty(16, np.int32)
        lib.orc_extractor_tables(self.h, _p(sc, _f32p), _p(inv, _f32p), _p(sg, _f32p), _p(isg, _f32p),
                                 _p(q, _i32p), _p(um, _i32p))
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2, self.quota, self.umax = sc, inv, sg, isg, q, um

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels + 4096
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib.orc_extractor_run(self.h, _p(img, _u8p), img.shape[1], img.shape[0], img.strides[0],
                                  kps.ctypes.data, _p(desc, _u8p), cap)
        if n < 0:
            raise ValueError("oracle extractor: input the reference cannot process")
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        """Bordered pyramid level (h+38, w+38)."""
        w, h = C.c_int(), C.c_int()
        lib.orc_extractor_level_dims(self.h, l, C.byref(w), C.byref(h))
        out = np.empty((h.value + 38, w.value + 38), np.uint8)
        lib.orc_extractor_level_copy(self.h, l, _p(out, _u8p))
        return out

    def candidates(self, l):
        n = lib.orc_extractor_candidates(self.h, l, None, 0)
        out = np.empty((max(n, 1), 3), np.int32)
        lib.orc_extractor_candidates(self.h, l, _p(out, _i32p), n)
        return out[:n]

    def level_counts(self):
        return np.array([lib.orc_extractor_level_count(self.h, l) for l in range(self.nlevels)], np.int32)


# ---------------------------------------------------------------------------------------------------
# matcher / dynamic oracle entry points (same POD layouts as include/sdyn.h; structs come from pysdyn's
# ctypes definitions, which only describe the ABI — no product code runs here)
# ---------------------------------------------------------------------------------------------------
def _structs():
    import pysdyn
    return pysdyn


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib.orc_hamming(_p(a, _u8p), _p(b, _u8p))


def features_in_area(F, x, y, r, min_level, max_level):
    out = np.zeros(max(F.n, 1), np.int32)
    lib.orc_features_in_area.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _i32p, C.c_int]
    n = lib.orc_features_in_area(C.byref(F.c), x, y, r, min_level, max_level, _p(out, _i32p), len(out))
    return out[:n].copy()


def match_projection_map(F, mappoints, th, nnratio, assign=None, locked=None, assign_base=0):
    S = _structs()
    mp = np.ascontiguousarray(mappoints, S.MAPPOINT_DTYPE)
    assign = np.full(F.n, -1, np.int32) if assign is None else np.ascontiguousarray(assign, np.int32).copy()
    locked = np.zeros(F.n, np.uint8) if locked is None else np.ascontiguousarray(locked, np.uint8).copy()
    lib.orc_match_projection_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
    n = lib.orc_match_projection_map(C.byref(F.c), mp.ctypes.data, len(mp), th, nnratio, assign.ctypes.data, locked.ctypes.data, assign_base)
    return n, assign, locked


def match_projection_frame(cur, last, last_points, th, mono, check_ori, assign=None, locked=None, want_pairs=False):
    S = _structs()
    lp = np.ascontiguousarray(last_points, S.LASTPOINT_DTYPE)
    assign = np.full(cur.n, -1, np.int32) if assign is None else np.ascontiguousarray(assign, np.int32).copy()
    locked = np.zeros(cur.n, np.uint8) if locked is None else np.ascontiguousarray(locked, np.uint8).copy()
    pairs = np.zeros((max(last.n, 1), 4), np.float32)
    npairs = C.c_int(0)
    lib.orc_match_projection_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    n = lib.orc_match_projection_frame(C.byref(cur.c), C.byref(last.c), lp.ctypes.data, th, int(mono), int(check_ori),
                                       assign.ctypes.data, locked.ctypes.data,
                                       pairs.ctypes.data if want_pairs else None, C.byref(npairs))
    if want_pairs:
        return n, assign, locked, pairs[:npairs.value].copy()
    return n, assign, locked


def match_init(F1, F2, prev_matched, window, nnratio, check_ori):
    prev = np.ascontiguousarray(prev_matched, np.float32).copy()
    m12 = np.full(F1.n, -1, np.int32)
    lib.orc_match_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
    n = lib.orc_match_init(C.byref(F1.c), C.byref(F2.c), prev.ctypes.data, m12.ctypes.data, window, nnratio, int(check_ori))
    return n, m12, prev


def match_bow(KF, kf_valid, fv_kf, F, fv_f, nnratio, check_ori):
    kv = np.ascontiguousarray(kf_valid, np.uint8)
    assign = np.full(F.n, -1, np.int32)
    lib.orc_match_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
    n = lib.orc_match_bow(C.byref(KF.c), kv.ctypes.data, C.byref(fv_kf.c), C.byref(F.c), C.byref(fv_f.c), nnratio,
                          int(check_ori), assign.ctypes.data)
    return n, assign


def match_bow_kf(KF1, valid1, fv1, KF2, valid2, fv2, nnratio, check_ori):
    v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(valid2, np.uint8)
    m12 = np.full(KF1.n, -1, np.int32)
    lib.orc_match_bow_kf.argtypes = [C.c_void_p] * 6 + [C.c_float, C.c_int, C.c_void_p]
    n = lib.orc_match_bow_kf(C.byref(KF1.c), v1.ctypes.data, C.byref(fv1.c), C.byref(KF2.c), v2.ctypes.data, C.byref(fv2.c),
                             nnratio, int(check_ori), m12.ctypes.data)
    return n, m12


def match_projection_pose(target, points, params, assign=None):
    pts = np.ascontiguousarray(points, _structs().PROJPOINT_DTYPE)
    assign = np.full(target.n, -1, np.int32) if assign is None else np.ascontiguousarray(assign, np.int32).copy()
    lib.orc_match_projection_pose.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    n = lib.orc_match_projection_pose(C.byref(target.c), pts.ctypes.data, len(pts), C.byref(params), assign.ctypes.data)
    return n, assign


def fuse_search(which, KF, inv_sigma2, points, rcw, tcw, ow, th, log_sf, nlevels):
    pts = np.ascontiguousarray(points, _structs().PROJPOINT_DTYPE)
    bi = np.full(max(len(pts), 1), -1, np.int32); bd = np.full(max(len(pts), 1), 256, np.int32)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    r_, t_, o_, s_ = f(rcw).reshape(9), f(tcw), f(ow), f(inv_sigma2)
    lib.orc_fuse_search.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    lib.orc_fuse_search(which, C.byref(KF.c), s_.ctypes.data, pts.ctypes.data, len(pts), r_.ctypes.data, t_.ctypes.data, o_.ctypes.data,
                        th, float(np.float32(log_sf)), nlevels, bi.ctypes.data, bd.ctypes.data)
    return bi[:len(pts)], bd[:len(pts)]


def search_by_sim3(KF1, KF2, pts1, pts2, T1w, T2w, S12, S21, th, log_sf, nlevels):
    p1 = np.ascontiguousarray(pts1, _structs().PROJPOINT_DTYPE); p2 = np.ascontiguousarray(pts2, _structs().PROJPOINT_DTYPE)
    f = lambda a: np.ascontiguousarray(a, np.float32).reshape(12)
    a, b, c, d = f(T1w), f(T2w), f(S12), f(S21)
    m12 = np.full(KF1.n, -1, np.int32)
    lib.orc_search_by_sim3.argtypes = [C.c_void_p] * 8 + [C.c_float, C.c_float, C.c_int, C.c_void_p]
    n = lib.orc_search_by_sim3(C.byref(KF1.c), C.byref(KF2.c), p1.ctypes.data, p2.ctypes.data, a.ctypes.data, b.ctypes.data,
                               c.ctypes.data, d.ctypes.data, th, float(np.float32(log_sf)), nlevels, m12.ctypes.data)
    return n, m12


def match_triangulation(KF1, has_mp1, fv1, KF2, has_mp2, fv2, params):
    h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
    m12 = np.full(KF1.n, -1, np.int32)
    lib.orc_match_triangulation.argtypes = [C.c_void_p] * 8
    n = lib.orc_match_triangulation(C.byref(KF1.c), h1.ctypes.data, C.byref(fv1.c), C.byref(KF2.c), h2.ctypes.data, C.byref(fv2.c),
                                    C.byref(params), m12.ctypes.data)
    return n, m12


def in_frustum(F, log_sf, world, normal, min_dist, max_dist, cos_limit=0.5):
    """Frame::isInFrustum for every point -> dict(in_view, proj_x, proj_y, proj_xr, level, view_cos)."""
    w = np.ascontiguousarray(world, np.float32).reshape(-1, 3); nrm = np.ascontiguousarray(normal, np.float32).reshape(-1, 3)
    mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32)
    n = len(w)
    iv = np.zeros(n, np.uint8); px, py, pxr, vc = (np.zeros(n, np.float32) for _ in range(4)); lv = np.zeros(n, np.int32)
    lib.orc_in_frustum.argtypes = [C.c_void_p, C.c_float, C.c_int] + [C.c_void_p] * 4 + [C.c_float] + [C.c_void_p] * 6
    lib.orc_in_frustum(C.byref(F.c), float(np.float32(log_sf)), n, w.ctypes.data, nrm.ctypes.data, mn.ctypes.data, mx.ctypes.data, cos_limit,
                       iv.ctypes.data, px.ctypes.data, py.ctypes.data, pxr.ctypes.data, lv.ctypes.data, vc.ctypes.data)
    return dict(in_view=iv, proj_x=px, proj_y=py, proj_xr=pxr, level=lv, view_cos=vc)


def box_mask(keys, boxes):
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
    mask = np.zeros(len(keys), np.uint64)
    lib.orc_box_mask.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.orc_box_mask(keys.ctypes.data, len(keys), boxes.ctypes.data, len(boxes), mask.ctypes.data)
    return mask


def separate_pairs(pairs, M, mode):
    S = _structs()
    lib.orc_separate_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return S.separate_pairs(None, pairs, M, mode, fn=lambda arr, n, m, md: lib.orc_separate_pairs(arr, n, m, md))


def classify(flag, M, cur_xy, ref_xy, query, train):
    """Tracking::classifyH (flag 1) / classifyF (flag 2) on explicit matches -> falseDyn."""
    c = np.ascontiguousarray(cur_xy, np.float32).reshape(-1, 2); r = np.ascontiguousarray(ref_xy, np.float32).reshape(-1, 2)
    q = np.ascontiguousarray(query, np.int32); t = np.ascontiguousarray(train, np.int32)
    m = np.ascontiguousarray(M, np.float32).reshape(9); out = np.full(len(q), -1, np.int32)
    lib.orc_classify.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    lib.orc_classify(flag, m.ctypes.data, c.ctypes.data, r.ctypes.data, q.ctypes.data, t.ctypes.data, len(q), out.ctypes.data)
    return out


def _box_csr(state):
    """state: dict(ku = mvKeysUn-like keys per ORIGINAL index, d = descriptors, per_box = [[original indices]], box_idx)."""
    off = np.zeros(len(state["per_box"]) + 1, np.int32)
    off[1:] = np.cumsum([len(m) for m in state["per_box"]])
    flat = [i for m in state["per_box"] for i in m]
    xy = np.ascontiguousarray(np.stack([state["ku"]["x"][flat], state["ku"]["y"][flat]], 1), np.float32).reshape(-1, 2)
    desc = np.ascontiguousarray(state["d"][flat], np.uint8).reshape(-1, 32)
    return off, xy, desc


def separate_frames(cur, ref, HorF, mode, last=None, cur_status=None, last_status=None):
    """Tracking::Separate for two frames described by their per-box dynamic lists.  mode 0 = classifyF, 1 = classifyH.
    Returns dict(ret, dyn_status=[array per box], status)."""
    last = ref if last is None else last
    co, cx, cd = _box_csr(cur); ro, rx, rd = _box_csr(ref)
    nb = len(cur["per_box"])
    cbi = np.ascontiguousarray(cur["box_idx"], np.int32); rbi = np.ascontiguousarray(ref["box_idx"], np.int32)
    lbi = np.ascontiguousarray(last["box_idx"], np.int32)
    cbs = np.full(nb, -1, np.int32) if cur_status is None else np.ascontiguousarray(cur_status, np.int32).copy()
    lbs = np.zeros(len(lbi), np.int32) if last_status is None else np.ascontiguousarray(last_status, np.int32)
    m = np.ascontiguousarray(HorF, np.float32).reshape(9)
    cap = max(len(cx), 1) + 16
    off = np.zeros(nb + 1, np.int32); val = np.zeros(cap, np.int32)
    lib.orc_separate.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    ret = lib.orc_separate(nb, co.ctypes.data, cx.ctypes.data, cd.ctypes.data, cbi.ctypes.data, cbs.ctypes.data,
                           len(ref["per_box"]), ro.ctypes.data, rx.ctypes.data, rd.ctypes.data, rbi.ctypes.data,
                           len(lbi), lbi.ctypes.data, lbs.ctypes.data, m.ctypes.data, 1 if mode == 1 else 2, off.ctypes.data, val.ctypes.data, cap)
    return dict(ret=ret, dyn_status=[val[off[b]:off[b + 1]].copy() for b in range(nb)], status=cbs)


def update_frame_list(dyn_status, dyn_class_id):
    """Frame::UpdateFrame: the (box, k) entries appended to the frame, in push order."""
    nb = len(dyn_status)
    so = np.zeros(nb + 1, np.int32); so[1:] = np.cumsum([len(d) for d in dyn_status])
    co = np.zeros(nb + 1, np.int32); co[1:] = np.cumsum([len(c) for c in dyn_class_id])
    sv = np.ascontiguousarray(np.concatenate([np.asarray(d, np.int32) for d in dyn_status] + [np.zeros(0, np.int32)]), np.int32)
    cv = np.ascontiguousarray(np.concatenate([np.asarray(c, np.int32) for c in dyn_class_id] + [np.zeros(0, np.int32)]), np.int32)
    cap = len(sv) + 1
    ob = np.zeros(cap, np.int32); ok = np.zeros(cap, np.int32)
    lib.orc_update_frame.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int]
    n = lib.orc_update_frame(nb, so.ctypes.data, sv.ctypes.data, co.ctypes.data, cv.ctypes.data, ob.ctypes.data, ok.ctypes.data, cap)
    return list(zip(ob[:n].tolist(), ok[:n].tolist()))


def invert3x3(m):
    m = np.ascontiguousarray(m, np.float32).reshape(9)
    out = np.zeros(9, np.float32)
    lib.orc_invert3x3.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_invert3x3(m.ctypes.data, out.ctypes.data)
    return out.reshape(3, 3)


def box_track(boxes, last_objects, last_box_idx, last_omit, last_vel, img_w, img_h):
    """Frame::boxTrack. Returns (boxes, box_idx, omit, velocity)."""
    cap = len(boxes) + len(last_objects) + 1
    b = np.zeros((cap, 4), np.float64); b[:len(boxes)] = np.asarray(boxes, np.float64).reshape(-1, 4)
    lo = np.ascontiguousarray(last_objects, np.float64).reshape(-1, 4)
    li = np.ascontiguousarray(last_box_idx, np.int32); lom = np.ascontiguousarray(last_omit, np.uint8)
    lv = np.ascontiguousarray(last_vel, np.float64).reshape(-1, 2)
    bi = np.zeros(cap, np.int32); om = np.zeros(cap, np.uint8); vel = np.zeros((cap, 2), np.float64)
    lib.orc_box_track.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    n = lib.orc_box_track(b.ctypes.data, len(boxes), cap, lo.ctypes.data, li.ctypes.data, lom.ctypes.data, lv.ctypes.data,
                          len(lo), img_w, img_h, bi.ctypes.data, om.ctypes.data, vel.ctypes.data)
    return b[:n].copy(), bi[:n].copy(), om[:n].copy(), vel[:n].copy()


def first_separate(keys, boxes, box_idx):
    """Frame::firstSeparate + tail split.  Returns dict(order, class_id, n_dyn, boxes, box_idx, dyn=[(box, key)])."""
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4).copy()
    bi = np.ascontiguousarray(box_idx, np.int32).copy()
    n = len(keys)
    order = np.zeros(max(n, 1), np.int32); cid = np.zeros(max(n, 1), np.int32)
    cap = max(n * max(len(b), 1), 1)
    db = np.zeros(cap, np.int32); dk = np.zeros(cap, np.int32)
    nd, npairs = C.c_int(0), C.c_int(0)
    lib.orc_first_separate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    nb = lib.orc_first_separate(keys.ctypes.data, n, b.ctypes.data, bi.ctypes.data, len(b), order.ctypes.data, cid.ctypes.data,
                                C.byref(nd), db.ctypes.data, dk.ctypes.data, cap, C.byref(npairs))
    return {"order": order[:n].copy(), "class_id": cid[:n].copy(), "n_dyn": nd.value, "boxes": b[:nb].copy(),
            "box_idx": bi[:nb].copy(), "dyn": list(zip(db[:npairs.value].tolist(), dk[:npairs.value].tolist()))}


# ---------------------------------------------------------------------------------------------------
# Frame::ComputeStereoMatches (src/Frame.cc:874-1048)
# ---------------------------------------------------------------------------------------------------
lib.orc_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _u8p, C.c_void_p, C.c_int, _u8p,
                                   C.c_float, C.c_float, _f32p, _f32p]


def stereo_matches(ex_left, ex_right, keys_l, desc_l, keys_r, desc_r, mb, mbf):
    """ex_left / ex_right: oracle Extractors whose LAST call produced the keypoints (their pyramids are read).
    Returns (mvuRight, mvDepth, kept)."""
    keys_l = np.ascontiguousarray(keys_l, KP_DTYPE); keys_r = np.ascontiguousarray(keys_r, KP_DTYPE)
    desc_l = np.ascontiguousarray(desc_l, np.uint8); desc_r = np.ascontiguousarray(desc_r, np.uint8)
    n = len(keys_l)
    ur = np.empty(max(n, 1), np.float32); dp = np.empty(max(n, 1), np.float32)
    rc = lib.orc_stereo_matches(ex_left.h, ex_right.h, keys_l.ctypes.data, n, _p(desc_l, _u8p), keys_r.ctypes.data,
                                len(keys_r), _p(desc_r, _u8p), mb, mbf, _p(ur, _f32p), _p(dp, _f32p))
    if rc < 0:
        raise ValueError("oracle stereo: input the reference cannot process (rc=%d)" % rc)
    return ur[:n], dp[:n], rc


# ---------------------------------------------------------------------------------------------------
# Frame::ComputeBoW -> DBoW2 TemplatedVocabulary::transform
# ---------------------------------------------------------------------------------------------------
class Vocabulary:
    def __init__(self, parent, is_leaf, desc, weight, k, L):
        self.parent = np.ascontiguousarray(parent, np.int32); self.is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        self.desc = np.ascontiguousarray(desc, np.uint8); self.weight = np.ascontiguousarray(weight, np.float64)
        lib.orc_vocab_build.restype = C.c_void_p
        lib.orc_vocab_build.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.orc_vocab_free.argtypes = [C.c_void_p]
        self.h = lib.orc_vocab_build(len(self.parent), self.parent.ctypes.data, self.is_leaf.ctypes.data, self.desc.ctypes.data,
                                     self.weight.ctypes.data, k, L)

    def __del__(self):
        if getattr(self, "h", None):
            lib.orc_vocab_free(self.h); self.h = None

    def transform(self, features, levelsup=4):
        """-> dict(word, weight, node per feature; bow_ids, bow_values; fv_nodes, fv_offset, fv_index)."""
        f = np.ascontiguousarray(features, np.uint8); n = len(f)
        m = max(n, 1)
        word = np.zeros(m, np.uint32); w = np.zeros(m, np.float64); node = np.zeros(m, np.uint32)
        bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64); fn = np.zeros(m, np.uint32)
        fo = np.zeros(m + 1, np.int32); fi = np.zeros(m, np.uint32)
        nw, nn = C.c_int(0), C.c_int(0)
        lib.orc_bow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int)] + \
                                        [C.c_void_p] * 3 + [C.POINTER(C.c_int)]
        lib.orc_bow_transform(self.h, f.ctypes.data, n, levelsup, word.ctypes.data, w.ctypes.data, node.ctypes.data, bi.ctypes.data,
                              bv.ctypes.data, C.byref(nw), fn.ctypes.data, fo.ctypes.data, fi.ctypes.data, C.byref(nn))
        return dict(word=word[:n], weight=w[:n], node=node[:n], bow_ids=bi[:nw.value], bow_values=bv[:nw.value],
                    fv_nodes=fn[:nn.value], fv_offset=fo[:nn.value + 1], fv_index=fi[:fo[nn.value]])
