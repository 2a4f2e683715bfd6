"""ctypes binding for oracle/_ref/libref.so — the REFERENCE'S OWN translation units compiled against oracle/ref_shim.
TEST INFRASTRUCTURE ONLY: imported by tests/ (and tools/ that generate fixtures), never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
# SDYN_REF_VARIANT = "dropin" selects oracle/_ref/libdropin.so: the reference's Frame.cc linked against the PRODUCT's host classes
# (same C entry points); see load_variant().
VARIANT = os.environ.get("SDYN_REF_VARIANT", "ref")
LIB_PATH = os.path.join(_ORACLE_DIR, "_ref", "lib%s.so" % VARIANT)
REFERENCE_TREE = "/root/reference"

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int)
_f32p = C.POINTER(C.c_float)


def available():
    """The library exists, or can be built because the reference tree is present in this container."""
    return os.path.exists(LIB_PATH) or os.path.isdir(REFERENCE_TREE)


def _load():
    if os.path.isdir(REFERENCE_TREE):
        subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR, VARIANT])
    return C.CDLL(LIB_PATH)


def load_variant(name):
    """A second, independent copy of this binding over oracle/_ref/lib<name>.so (module object)."""
    import importlib.util
    old = os.environ.get("SDYN_REF_VARIANT")
    os.environ["SDYN_REF_VARIANT"] = name
    try:
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.abspath(__file__))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if old is None:
            os.environ.pop("SDYN_REF_VARIANT", None)
        else:
            os.environ["SDYN_REF_VARIANT"] = old
    return mod


lib = _load()


def _p(a, t):
    return a.ctypes.data_as(t)


lib.ref_extractor_create.restype = C.c_void_p
lib.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
lib.ref_extractor_destroy.argtypes = [C.c_void_p]
lib.ref_extractor_run.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, C.c_int]
lib.ref_extractor_tables.argtypes = [C.c_void_p] + [_f32p] * 4
lib.ref_extractor_level_dims.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
lib.ref_extractor_level_copy.argtypes = [C.c_void_p, C.c_int, _u8p]

ALLOC_MALLOC, ALLOC_BUMP = 0, 1


def set_alloc_mode(mode):
    """ORBextractor.cc:684 orders equal-size octree nodes by heap address: 1 = monotonic arena, 0 = glibc malloc."""
    lib.ref_set_alloc_mode(mode)


class Extractor:
    """ORB_SLAM2::ORBextractor of the reference (src/ORBextractor.cc, compiled unchanged)."""

    def __init__(self, nfeatures, scale, nlevels, ini, mn):
        self.h = lib.ref_extractor_create(nfeatures, scale, nlevels, ini, mn)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def close(self):
        if self.h:
            lib.ref_extractor_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures * 2 + 64
        k = np.zeros(cap, KP_DTYPE)
        d = np.zeros((cap, 32), np.uint8)
        n = lib.ref_extractor_run(self.h, _p(img, _u8p), w, h, img.strides[0], k.ctypes.data, _p(d, _u8p), cap)
        assert 0 <= n <= cap, n
        return k[:n].copy(), d[:n].copy()

    def tables(self):
        out = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        lib.ref_extractor_tables(self.h, *[_p(a, _f32p) for a in out])
        return out

    def level(self, l):
        w, h = C.c_int(), C.c_int()
        if lib.ref_extractor_level_dims(self.h, l, C.byref(w), C.byref(h)) != 0:
            return None
        out = np.empty((h.value + 38, w.value + 38), np.uint8)
        lib.ref_extractor_level_copy(self.h, l, _p(out, _u8p))
        return out


# ---------------------------------------------------------------------------------------------------
# Frame / ORBmatcher / Tracking::Separate of the reference (src/Frame.cc, src/ORBmatcher.cc, src/Tracking.cc:1093-1367)
# ---------------------------------------------------------------------------------------------------
_vp = C.c_void_p
_f64p = C.POINTER(C.c_double)
lib.ref_frame_rgbd_boxes.restype = _vp
lib.ref_frame_rgbd_boxes.argtypes = [_vp, _u8p, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_float, C.c_float]
lib.ref_frame_stereo.restype = _vp
lib.ref_frame_stereo.argtypes = [_vp, _vp, _u8p, _u8p, C.c_int, C.c_int, _vp, C.c_float, C.c_float]
lib.ref_frame_from_arrays.restype = _vp
lib.ref_frame_from_arrays.argtypes = [_vp] * 5 + [C.c_int] + [_vp] * 3
lib.ref_frame_destroy.argtypes = [_vp]
for _n in ("ref_frame_n", "ref_frame_n_right", "ref_frame_n_dyn"):
    getattr(lib, _n).argtypes = [_vp]
lib.ref_frame_keys.argtypes = [_vp, C.c_int, _vp]
lib.ref_frame_descriptors.argtypes = [_vp, C.c_int, _vp]
lib.ref_frame_stereo_values.argtypes = [_vp, _vp, _vp]
lib.ref_frame_grid.argtypes = [_vp, _vp, _vp, C.c_int]
lib.ref_frame_features_in_area.argtypes = [_vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _vp, C.c_int]
lib.ref_frame_boxes.argtypes = [_vp] * 6 + [C.c_int]
lib.ref_frame_set_box_status.argtypes = [_vp, _vp, C.c_int]
lib.ref_frame_dyn_count.argtypes = [_vp, C.c_int]
lib.ref_frame_dyn.argtypes = [_vp, C.c_int] + [_vp] * 5
lib.ref_frame_set_pose.argtypes = [_vp, _vp]
lib.ref_points_create.restype = _vp
lib.ref_points_create.argtypes = [C.c_int] + [_vp] * 8
lib.ref_points_destroy.argtypes = [_vp]
lib.ref_points_set_track.argtypes = [_vp] * 7
lib.ref_points_in_frustum.argtypes = [_vp, _vp, C.c_float] + [_vp] * 6
lib.ref_frame_set_points.argtypes = [_vp, _vp, _vp]
lib.ref_frame_report_against.argtypes = [_vp, _vp]
lib.ref_frame_clear_points.argtypes = [_vp]
lib.ref_frame_preassign.argtypes = [_vp, _vp, _vp]
lib.ref_frame_assignment.argtypes = [_vp, _vp]
lib.ref_frame_set_featvec.argtypes = [_vp, C.c_int, _vp, _vp, _vp]
lib.ref_descriptor_distance.argtypes = [_vp, _vp]
lib.ref_search_by_projection_map.argtypes = [_vp, _vp, C.c_float, C.c_float]
lib.ref_search_by_projection_frame.argtypes = [_vp, _vp, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, _vp, C.c_int, _i32p]
lib.ref_search_for_initialization.argtypes = [_vp, _vp, _vp, _vp, C.c_int, C.c_float, C.c_int]
lib.ref_frame_make_keyframe.argtypes = [_vp]
lib.ref_search_by_bow_frame.argtypes = [_vp, _vp, C.c_float, C.c_int, _vp]
lib.ref_search_by_bow_kf.argtypes = [_vp, _vp, C.c_float, C.c_int, _vp]
lib.ref_search_for_triangulation.argtypes = [_vp, _vp, _vp, C.c_int, C.c_float, C.c_int, _vp]
lib.ref_search_by_projection_reloc.argtypes = [_vp, _vp, C.c_float, C.c_int, C.c_float, C.c_int, _vp]
lib.ref_search_by_projection_sim3.argtypes = [_vp, _vp, _vp, C.c_int, C.c_float, _vp]
lib.ref_fuse.argtypes = [_vp, _vp, C.c_float, C.c_float, _vp]
lib.ref_fuse_sim3.argtypes = [_vp, _vp, _vp, C.c_float, C.c_float, _vp]
lib.ref_search_by_sim3.argtypes = [_vp, _vp, _vp, C.c_float, _vp, _vp, C.c_float, C.c_float]
lib.ref_frame_update.argtypes = [_vp, _vp, _vp, C.c_int]
lib.ref_tracking_separate.argtypes = [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp]
lib.ref_classify.argtypes = [C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, C.c_int, _vp]
lib.ref_cv_gemm.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, C.c_double, _vp, C.c_double, _vp]
lib.ref_cv_invert3x3.argtypes = [_vp, _vp]
lib.ref_cv_norm.restype = C.c_double
lib.ref_cv_norm.argtypes = [_vp, C.c_int]
lib.ref_cv_dot.restype = C.c_double
lib.ref_cv_dot.argtypes = [_vp, _vp, C.c_int]
lib.ref_cv_bfmatch.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def reset_statics():
    """Frame::mbInitialComputations = true (a new calibration / image size follows)."""
    lib.ref_frame_reset_statics()


def statics():
    out = np.zeros(10, np.float32)
    lib.ref_frame_get_statics(out.ctypes.data)
    return dict(bounds=tuple(out[:4].tolist()), grid_inv=tuple(out[4:6].tolist()), cam=tuple(out[6:10].tolist()))


class Points:
    """A list of MapPoint stand-ins (ref_shim/ref_entities.h); index i of the list is how the point is reported back."""

    def __init__(self, world, desc, present=None, normal=None, min_dist=None, max_dist=None, nobs=None, bad=None):
        n = len(desc)
        self.n = n
        self._keep = [_f32(world).reshape(-1, 3), np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)]
        opt = lambda a, t: None if a is None else np.ascontiguousarray(a, t)
        self._opt = [opt(present, np.uint8), opt(normal, np.float32), opt(min_dist, np.float32), opt(max_dist, np.float32),
                     opt(nobs, np.int32), opt(bad, np.uint8)]
        p = lambda a: None if a is None else a.ctypes.data
        self.h = lib.ref_points_create(n, p(self._opt[0]), self._keep[0].ctypes.data, p(self._opt[1]), self._keep[1].ctypes.data,
                                       p(self._opt[2]), p(self._opt[3]), p(self._opt[4]), p(self._opt[5]))

    def set_track(self, in_view, proj_x, proj_y, proj_xr, level, view_cos):
        a = [np.ascontiguousarray(in_view, np.uint8), _f32(proj_x), _f32(proj_y), _f32(proj_xr),
             np.ascontiguousarray(level, np.int32), _f32(view_cos)]
        lib.ref_points_set_track(self.h, *[x.ctypes.data for x in a])

    def __del__(self):
        if getattr(self, "h", None):
            lib.ref_points_destroy(self.h)
            self.h = None


class Frame:
    """ORB_SLAM2::Frame of the reference."""

    def __init__(self, handle, extractor):
        self.h = handle
        self.ex = extractor          # keeps the ORBextractor alive

    @classmethod
    def from_arrays(cls, extractor, keys, desc, bounds, cam, keys_un=None, u_right=None, tcw=None):
        keys = np.ascontiguousarray(keys, KP_DTYPE)
        ku = None if keys_un is None else np.ascontiguousarray(keys_un, KP_DTYPE)
        desc = np.ascontiguousarray(desc, np.uint8)
        ur = None if u_right is None else _f32(u_right)
        b, c = _f32(bounds), _f32(cam)
        t = None if tcw is None else _f32(tcw).reshape(-1)[:12].copy()
        p = lambda a: None if a is None else a.ctypes.data
        return cls(lib.ref_frame_from_arrays(extractor.h, keys.ctypes.data, p(ku), desc.ctypes.data, p(ur), len(keys), b.ctypes.data,
                                             c.ctypes.data, p(t)), extractor)

    @classmethod
    def from_view(cls, extractor, view):
        """A pysdyn.FrameView (what the oracle and the C ABI consume) as a reference Frame."""
        return cls.from_arrays(extractor, view.keys, view.desc, view.bounds, view.cam, keys_un=view.keys_un, u_right=view.u_right,
                               tcw=view.tcw)

    @classmethod
    def rgbd_boxes(cls, extractor, gray, boxes, last=None, depth=None, K=(500.0, 500.0, 320.0, 240.0), dist=(0, 0, 0, 0), bf=40.0,
                   th_depth=40.0):
        gray = np.ascontiguousarray(gray, np.uint8)
        h, w = gray.shape
        bx = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        dep = None if depth is None else _f32(depth)
        K4, D = _f32(K), _f32(dist)
        return cls(lib.ref_frame_rgbd_boxes(extractor.h, _p(gray, _u8p), w, h, None if dep is None else dep.ctypes.data, bx.ctypes.data,
                                            len(bx), last.h if last is not None else None, K4.ctypes.data, D.ctypes.data, len(D), bf,
                                            th_depth), extractor)

    @classmethod
    def stereo(cls, ex_left, ex_right, left, right, K, bf, th_depth=35.0):
        left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
        h, w = left.shape
        K4 = _f32(K)
        f = cls(lib.ref_frame_stereo(ex_left.h, ex_right.h, _p(left, _u8p), _p(right, _u8p), w, h, K4.ctypes.data, bf, th_depth), ex_left)
        f.ex_right = ex_right
        return f

    def close(self):
        if self.h:
            lib.ref_frame_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def n(self):
        return lib.ref_frame_n(self.h)

    def keys(self, which=0):
        n = lib.ref_frame_n_right(self.h) if which == 2 else self.n
        out = np.zeros(n, KP_DTYPE)
        lib.ref_frame_keys(self.h, which, out.ctypes.data)
        return out

    def descriptors(self, right=False):
        n = lib.ref_frame_n_right(self.h) if right else self.n
        out = np.zeros((n, 32), np.uint8)
        lib.ref_frame_descriptors(self.h, int(right), out.ctypes.data)
        return out

    def stereo_values(self):
        u, d = np.zeros(self.n, np.float32), np.zeros(self.n, np.float32)
        lib.ref_frame_stereo_values(self.h, u.ctypes.data, d.ctypes.data)
        return u, d

    def grid(self):
        counts = np.zeros(64 * 48, np.int32); entries = np.zeros(4 * self.n + 16, np.int32)
        n = lib.ref_frame_grid(self.h, counts.ctypes.data, entries.ctypes.data, len(entries))
        return counts, entries[:n].copy()

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(4 * self.n + 16, np.int32)
        n = lib.ref_frame_features_in_area(self.h, x, y, r, min_level, max_level, out.ctypes.data, len(out))
        return out[:n].copy()

    def boxes(self):
        cap = 256
        ob = np.zeros((cap, 4), np.float64); bi = np.zeros(cap, np.int32); om = np.zeros(cap, np.uint8)
        vel = np.zeros((cap, 2), np.float64); st = np.zeros(cap, np.int32)
        n = lib.ref_frame_boxes(self.h, ob.ctypes.data, bi.ctypes.data, om.ctypes.data, vel.ctypes.data, st.ctypes.data, cap)
        return dict(objects=ob[:n].copy(), box_idx=bi[:n].copy(), omit=om[:n].copy(), velocity=vel[:n].copy(), status=st[:n].copy())

    def set_box_status(self, status):
        s = np.ascontiguousarray(status, np.int32)
        lib.ref_frame_set_box_status(self.h, s.ctypes.data, len(s))

    @property
    def n_dyn(self):
        return lib.ref_frame_n_dyn(self.h)

    def dyn(self, box):
        n = lib.ref_frame_dyn_count(self.h, box)
        k, ku = np.zeros(n, KP_DTYPE), np.zeros(n, KP_DTYPE)
        d = np.zeros((n, 32), np.uint8); u, z = np.zeros(n, np.float32), np.zeros(n, np.float32)
        if n:
            lib.ref_frame_dyn(self.h, box, k.ctypes.data, ku.ctypes.data, d.ctypes.data, u.ctypes.data, z.ctypes.data)
        return dict(keys=k, keys_un=ku, desc=d, u_right=u, depth=z)

    def set_pose(self, tcw):
        t = _f32(tcw).reshape(-1)[:12].copy()
        lib.ref_frame_set_pose(self.h, t.ctypes.data)

    def set_points(self, points, outlier=None):
        self._pts = points
        o = None if outlier is None else np.ascontiguousarray(outlier, np.uint8)
        lib.ref_frame_set_points(self.h, points.h, None if o is None else o.ctypes.data)

    def report_against(self, points):
        self._pts = points
        lib.ref_frame_report_against(self.h, points.h)

    def clear_points(self):
        lib.ref_frame_clear_points(self.h)

    def preassign(self, points, assign):
        self._occ = points
        a = np.ascontiguousarray(assign, np.int32)
        lib.ref_frame_preassign(self.h, points.h, a.ctypes.data)

    def assignment(self):
        out = np.zeros(self.n, np.int32)
        lib.ref_frame_assignment(self.h, out.ctypes.data)
        return out

    def set_featvec(self, fv):
        lib.ref_frame_set_featvec(self.h, len(fv.node_id), fv.node_id.ctypes.data, fv.offset.ctypes.data, fv.index.ctypes.data)

    def make_keyframe(self):
        lib.ref_frame_make_keyframe(self.h)

    def update(self, dyn_status):
        off = np.zeros(len(dyn_status) + 1, np.int32)
        off[1:] = np.cumsum([len(d) for d in dyn_status])
        vals = np.ascontiguousarray(np.concatenate([np.asarray(d, np.int32) for d in dyn_status] + [np.zeros(0, np.int32)]), np.int32)
        lib.ref_frame_update(self.h, off.ctypes.data, vals.ctypes.data, len(dyn_status))


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib.ref_descriptor_distance(a.ctypes.data, b.ctypes.data)


def search_by_projection_map(F, points, th, nnratio):
    return lib.ref_search_by_projection_map(F.h, points.h, th, nnratio)


def search_by_projection_frame(cur, last, th, mono, nnratio=0.9, check_ori=True, want_pairs=False):
    pairs = np.zeros((max(last.n, 1), 4), np.float32); npairs = C.c_int(0)
    n = lib.ref_search_by_projection_frame(cur.h, last.h, th, int(mono), nnratio, int(check_ori), int(want_pairs), pairs.ctypes.data,
                                           len(pairs), C.byref(npairs))
    return (n, pairs[:npairs.value].copy()) if want_pairs else n


def search_for_initialization(F1, F2, prev_matched, window, nnratio, check_ori):
    prev = _f32(prev_matched).copy(); m12 = np.full(F1.n, -1, np.int32)
    n = lib.ref_search_for_initialization(F1.h, F2.h, prev.ctypes.data, m12.ctypes.data, window, nnratio, int(check_ori))
    return n, m12, prev


def search_by_bow_frame(KF, F, nnratio, check_ori):
    assign = np.full(F.n, -1, np.int32)
    n = lib.ref_search_by_bow_frame(KF.h, F.h, nnratio, int(check_ori), assign.ctypes.data)
    return n, assign


def search_by_bow_kf(KF1, KF2, nnratio, check_ori):
    m12 = np.full(KF1.n, -1, np.int32)
    n = lib.ref_search_by_bow_kf(KF1.h, KF2.h, nnratio, int(check_ori), m12.ctypes.data)
    return n, m12


def search_for_triangulation(KF1, KF2, F12, only_stereo, nnratio=0.6, check_ori=False):
    f = _f32(F12).reshape(9); m12 = np.full(KF1.n, -1, np.int32)
    n = lib.ref_search_for_triangulation(KF1.h, KF2.h, f.ctypes.data, int(only_stereo), nnratio, int(check_ori), m12.ctypes.data)
    return n, m12


def search_by_projection_reloc(cur, KF, th, orb_dist, nnratio=0.9, check_ori=True):
    assign = np.full(cur.n, -1, np.int32)
    n = lib.ref_search_by_projection_reloc(cur.h, KF.h, th, orb_dist, nnratio, int(check_ori), assign.ctypes.data)
    return n, assign


def search_by_projection_sim3(KF, Scw, points, th, nnratio=0.75, matched=None):
    s = _f32(Scw).reshape(16)
    matched = np.full(KF.n, -1, np.int32) if matched is None else np.ascontiguousarray(matched, np.int32).copy()
    n = lib.ref_search_by_projection_sim3(KF.h, s.ctypes.data, points.h, th, nnratio, matched.ctypes.data)
    return n, matched


def fuse(KF, points, th, nnratio=0.6):
    out = np.full(points.n, -1, np.int32)
    n = lib.ref_fuse(KF.h, points.h, th, nnratio, out.ctypes.data)
    return n, out


def fuse_sim3(KF, Scw, points, th, nnratio=0.6):
    s = _f32(Scw).reshape(16); out = np.full(points.n, -1, np.int32)
    n = lib.ref_fuse_sim3(KF.h, s.ctypes.data, points.h, th, nnratio, out.ctypes.data)
    return n, out


def search_by_sim3(KF1, KF2, matches12, s12, R12, t12, th, nnratio=0.6):
    m = np.ascontiguousarray(matches12, np.int32).copy(); R = _f32(R12).reshape(9); t = _f32(t12).reshape(3)
    n = lib.ref_search_by_sim3(KF1.h, KF2.h, m.ctypes.data, s12, R.ctypes.data, t.ctypes.data, th, nnratio)
    return n, m


def points_in_frustum(F, points, cos_limit=0.5):
    """Frame::isInFrustum of the reference on every point of the list."""
    n = points.n
    iv = np.zeros(n, np.uint8); px, py, pxr, vc = (np.zeros(n, np.float32) for _ in range(4)); lv = np.zeros(n, np.int32)
    lib.ref_points_in_frustum(F.h, points.h, cos_limit, iv.ctypes.data, px.ctypes.data, py.ctypes.data, pxr.ctypes.data, lv.ctypes.data, vc.ctypes.data)
    return dict(in_view=iv, proj_x=px, proj_y=py, proj_xr=pxr, level=lv, view_cos=vc)


def tracking_separate(cur, ref_frame, last, HorF, flag):
    """Tracking::Separate.  Returns (ret, dynStatus per box, box_status of the current frame afterwards)."""
    nb = len(cur.boxes()["box_idx"])
    off = np.zeros(nb + 2, np.int32); cap = 4 * cur.n + 4096
    vals = np.zeros(cap, np.int32); st = np.zeros(nb + 1, np.int32)
    m = _f32(HorF).reshape(9)
    r = lib.ref_tracking_separate(cur.h, ref_frame.h, last.h, m.ctypes.data, flag, off.ctypes.data, vals.ctypes.data, cap, st.ctypes.data)
    return r, [vals[off[b]:off[b + 1]].copy() for b in range(nb)], st[:nb].copy()


def classify(flag, M, cur_xy, ref_xy, query, train):
    c, r = _f32(cur_xy).reshape(-1, 2), _f32(ref_xy).reshape(-1, 2)
    q, t = np.ascontiguousarray(query, np.int32), np.ascontiguousarray(train, np.int32)
    m = _f32(M).reshape(9); out = np.full(len(q), -1, np.int32)
    lib.ref_classify(flag, m.ctypes.data, c.ctypes.data, len(c), r.ctypes.data, len(r), q.ctypes.data, t.ctypes.data, len(q), out.ctypes.data)
    return out


def cv_gemm(A, B, alpha=1.0, Cm=None, beta=0.0, a_t=False):
    A, B = _f32(A), _f32(B)
    rows = A.shape[1] if a_t else A.shape[0]
    out = np.zeros((rows, B.shape[1]), np.float32)
    c = None if Cm is None else _f32(Cm)
    lib.ref_cv_gemm(A.ctypes.data, A.shape[0], A.shape[1], int(a_t), B.ctypes.data, B.shape[0], B.shape[1], alpha,
                    None if c is None else c.ctypes.data, beta, out.ctypes.data)
    return out


def cv_invert3x3(M):
    m = _f32(M).reshape(9); out = np.zeros(9, np.float32)
    lib.ref_cv_invert3x3(m.ctypes.data, out.ctypes.data)
    return out.reshape(3, 3)


def cv_norm(v):
    v = _f32(v).reshape(-1)
    return lib.ref_cv_norm(v.ctypes.data, len(v))


def cv_dot(a, b):
    a, b = _f32(a).reshape(-1), _f32(b).reshape(-1)
    return lib.ref_cv_dot(a.ctypes.data, b.ctypes.data, len(a))


def cv_bfmatch(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    a, b, d = (np.zeros(max(len(q), 1), np.int32) for _ in range(3))
    n = lib.ref_cv_bfmatch(q.ctypes.data, len(q), t.ctypes.data, len(t), a.ctypes.data, b.ctypes.data, d.ctypes.data)
    return list(zip(a[:n].tolist(), b[:n].tolist(), d[:n].tolist()))
