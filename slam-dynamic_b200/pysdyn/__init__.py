"""ctypes binding of the sdyn C-ABI (include/sdyn.h) for tests and bench.py.

This is plumbing: it loads slam-dynamic_b200/libsdyn.so (hand-written sm_100a kernels behind a C ABI)
and raises if the library is missing or a call fails — there is no CPU or eager fallback.
"""
import ctypes as C
import os
import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "libsdyn.so")
SYNTH_PATH = os.path.join(_PKG, "synth", "libsdyn_synth.so")

MAX_LEVELS = 16

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class ScaleInfo(C.Structure):
    _fields_ = [("nlevels", C.c_int32), ("scale", C.c_float * MAX_LEVELS), ("inv_scale", C.c_float * MAX_LEVELS),
                ("sigma2", C.c_float * MAX_LEVELS), ("inv_sigma2", C.c_float * MAX_LEVELS),
                ("features_per_level", C.c_int32 * MAX_LEVELS)]


class LevelInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("pitch", C.c_int32), ("reserved", C.c_int32),
                ("offset", C.c_size_t)]


class DeviceView(C.Structure):
    _fields_ = [("kp", C.c_void_p), ("desc", C.c_void_p), ("count", C.c_void_p), ("level_count", C.c_void_p),
                ("pyramid", C.c_void_p), ("blurred", C.c_void_p), ("pyramid_frame_bytes", C.c_size_t),
                ("cap", C.c_int32), ("nlevels", C.c_int32), ("level", LevelInfo * MAX_LEVELS)]


STAGES = ["pyramid", "fast", "octree", "blur", "describe", "match", "dynamic", "level0", "stereo", "bow", "candidates"]


class StageTimes(C.Structure):
    _fields_ = [("ms", C.c_double * len(STAGES)), ("calls", C.c_longlong * len(STAGES))]


class SdynError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sdyn error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Loads libsdyn.so; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libsdyn.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C slam-dynamic_b200`")
        L = C.CDLL(LIB_PATH)
        L.sdyn_last_error.restype = C.c_char_p
        L.sdyn_last_error.argtypes = [C.c_void_p]
        L.sdyn_create.argtypes = [C.POINTER(OrbParams), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sdyn_destroy.argtypes = [C.c_void_p]
        L.sdyn_scale_info_get.argtypes = [C.c_void_p, C.POINTER(ScaleInfo)]
        L.sdyn_max_keypoints.argtypes = [C.c_void_p]
        L.sdyn_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.sdyn_host_free.argtypes = [C.c_void_p]
        L.sdyn_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.POINTER(C.c_int), C.c_void_p]
        L.sdyn_extract_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sdyn_extract_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                                C.c_int, C.c_void_p]
        L.sdyn_device_results.argtypes = [C.c_void_p, C.POINTER(DeviceView)]
        L.sdyn_fetch_results.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.sdyn_fetch_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.sdyn_fetch_candidates.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.sdyn_sync.argtypes = [C.c_void_p]
        L.sdyn_profile_enable.argtypes = [C.c_void_p, C.c_int]
        L.sdyn_profile_read.argtypes = [C.c_void_p, C.POINTER(StageTimes)]
        L.sdyn_launch_count.restype = C.c_longlong
        L.sdyn_launch_count.argtypes = [C.c_void_p]
        _lib = L
    return _lib


_synth = None


def synth():
    global _synth
    if _synth is None:
        S = C.CDLL(SYNTH_PATH)
        S.sdyn_synth_frame.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int]
        S.sdyn_synth_boxes.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_int]
        S.sdyn_synth_boxes_ids.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_int]
        _synth = S
    return _synth


def synth_frame(seq_seed, frame_seed, w, h, nrect, ox=0, oy=0, t=0, out=None):
    """Seeded integer-only synthetic frame."""
    if out is None:
        out = np.empty((h, w), np.uint8)
    synth().sdyn_synth_frame(seq_seed, frame_seed, w, h, nrect, ox, oy, t, out.ctypes.data, out.strides[0])
    return out


def synth_boxes(seq_seed, w, h, nrect, ox=0, oy=0, t=0, margin=6, cap=64):
    b = np.zeros((cap, 4), np.float64)
    n = synth().sdyn_synth_boxes(seq_seed, w, h, nrect, ox, oy, t, margin, b.ctypes.data, cap)
    return b[:n].copy()


def synth_boxes_ids(seq_seed, w, h, nrect, ox=0, oy=0, t=0, margin=6, cap=64):
    b = np.zeros((cap, 4), np.float64)
    ids = np.zeros(cap, np.int32)
    n = synth().sdyn_synth_boxes_ids(seq_seed, w, h, nrect, ox, oy, t, margin, b.ctypes.data, ids.ctypes.data, cap)
    return b[:n].copy(), ids[:n].copy()


class PinnedArray:
    """numpy view over cudaMallocHost memory (pinned), freed with the object."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        rc = lib().sdyn_host_alloc(C.byref(p), max(nbytes, 1))
        if rc != 0:
            raise SdynError(rc, "pinned allocation of %d bytes failed" % nbytes)
        self.ptr = p
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        if getattr(self, "ptr", None) is not None and self.ptr.value:
            self.array = None
            lib().sdyn_host_free(self.ptr)
            self.ptr = None


class Extractor:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111) over the C ABI.

    __call__(image) -> (keypoints[KP_DTYPE], descriptors[n,32]) like operator()(image, mask, keypoints,
    descriptors); GetLevels()/GetScaleFactors()/... return the same values as the reference getters.
    """

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7,
                 max_width=1241, max_height=376, max_batch=1, device=0):
        self._h = C.c_void_p()
        p = OrbParams(nfeatures, scale_factor, nlevels, ini_th, min_th)
        rc = lib().sdyn_create(C.byref(p), max_width, max_height, max_batch, device, C.byref(self._h))
        if rc != 0:
            raise SdynError(rc, lib().sdyn_last_error(None).decode())
        self.max_batch = max_batch
        self.nlevels = nlevels
        si = ScaleInfo()
        self._check(lib().sdyn_scale_info_get(self._h, C.byref(si)))
        self.mvScaleFactor = np.array(si.scale[:nlevels], np.float32)
        self.mvInvScaleFactor = np.array(si.inv_scale[:nlevels], np.float32)
        self.mvLevelSigma2 = np.array(si.sigma2[:nlevels], np.float32)
        self.mvInvLevelSigma2 = np.array(si.inv_sigma2[:nlevels], np.float32)
        self.mnFeaturesPerLevel = np.array(si.features_per_level[:nlevels], np.int32)
        self.scaleFactor = float(np.float32(scale_factor))
        self.cap = lib().sdyn_max_keypoints(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().sdyn_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise SdynError(rc, lib().sdyn_last_error(self._h).decode())

    # reference getters
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self.mvScaleFactor.copy()
    def GetInverseScaleFactors(self): return self.mvInvScaleFactor.copy()
    def GetScaleSigmaSquares(self): return self.mvLevelSigma2.copy()
    def GetInverseScaleSigmaSquares(self): return self.mvInvLevelSigma2.copy()

    def __call__(self, image, want_pyramid=False):
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected (ORBextractor.cc:1050)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = C.c_int(0)
        self._check(lib().sdyn_extract(self._h, image.ctypes.data, w, h, image.strides[0], kps.ctypes.data,
                                       desc.ctypes.data, self.cap, C.byref(n), None))
        out = kps[:n.value].copy(), desc[:n.value].copy()
        if want_pyramid:
            return out + ([self.level(0, l) for l in range(self.nlevels)],)
        return out

    def extract_batch(self, images, kps=None, desc=None, counts=None):
        """images: [B,H,W] uint8 (pinned or pageable host memory). Returns (kps[B,cap], desc[B,cap,32], n[B])."""
        b, h, w = images.shape
        if kps is None:
            kps = np.zeros((b, self.cap), KP_DTYPE)
            desc = np.zeros((b, self.cap, 32), np.uint8)
            counts = np.zeros(b, np.int32)
        self._check(lib().sdyn_extract_batch(self._h, b, images.ctypes.data, images.strides[0], w, h,
                                             images.strides[1], kps.ctypes.data, desc.ctypes.data, self.cap,
                                             counts.ctypes.data))
        return kps, desc, counts

    def extract_batch_device(self, dptr, nframes, frame_stride, w, h, row_stride, stream=None):
        """Enqueue extraction of frames already in device memory (dptr = integer device address)."""
        self._check(lib().sdyn_extract_batch_device(self._h, nframes, C.c_void_p(dptr), frame_stride, w, h,
                                                    row_stride, C.c_void_p(stream) if stream else None))

    def fetch(self, nframes, stream=None, kps=None, desc=None, counts=None):
        if kps is None:
            kps = np.zeros((nframes, self.cap), KP_DTYPE)
            desc = np.zeros((nframes, self.cap, 32), np.uint8)
            counts = np.zeros(nframes, np.int32)
        self._check(lib().sdyn_fetch_results(self._h, nframes, kps.ctypes.data, desc.ctypes.data, self.cap,
                                             counts.ctypes.data, C.c_void_p(stream) if stream else None))
        return kps, desc, counts

    def device_view(self):
        v = DeviceView()
        self._check(lib().sdyn_device_results(self._h, C.byref(v)))
        return v

    def level(self, frame, l):
        """Bordered pyramid level l of `frame` of the last call: (h+38, w+38) uint8 (mvImagePyramid[l])."""
        v = self.device_view()
        w, h = v.level[l].width, v.level[l].height
        out = np.empty((h + 38, w + 38), np.uint8)
        self._check(lib().sdyn_fetch_level(self._h, frame, l, out.ctypes.data, None, None))
        return out

    def candidates(self, frame, l):
        n = C.c_int(0)
        v = self.device_view()
        cap = v.level[l].width * v.level[l].height // 4 + 16
        out = np.zeros((cap, 3), np.int32)
        self._check(lib().sdyn_fetch_candidates(self._h, frame, l, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    def set_camera(self, fx, fy, cx, cy, dist_coef=()):
        """mK / mDistCoef of the Frames built from this extractor; k1 != 0 turns on device-side mvKeysUn."""
        d = np.ascontiguousarray(dist_coef, np.float32)
        L = lib()
        L.sdyn_set_camera.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int]
        self._check(L.sdyn_set_camera(self._h, fx, fy, cx, cy, d.ctypes.data if len(d) else None, len(d)))

    def fetch_keypoints_un(self, nframes, stream=None):
        out = np.zeros((nframes, self.cap), KP_DTYPE)
        L = lib()
        L.sdyn_fetch_keypoints_un.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self._check(L.sdyn_fetch_keypoints_un(self._h, nframes, out.ctypes.data, self.cap, C.c_void_p(stream) if stream else None))
        return out

    def undistort_points(self, xy):
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.empty_like(xy)
        L = lib()
        L.sdyn_undistort_points.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self._check(L.sdyn_undistort_points(self._h, xy.ctypes.data, len(xy), out.ctypes.data))
        return out

    def image_bounds(self, width, height):
        b = (C.c_float * 4)()
        L = lib()
        L.sdyn_image_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self._check(L.sdyn_image_bounds(self._h, width, height, b))
        return tuple(b)

    def sync(self):
        self._check(lib().sdyn_sync(self._h))

    def profile(self, on):
        self._check(lib().sdyn_profile_enable(self._h, int(on)))

    def profile_read(self):
        """{stage: (ms, calls)} accumulated since the last read (device time, CUDA events on the stream)."""
        t = StageTimes()
        self._check(lib().sdyn_profile_read(self._h, C.byref(t)))
        return {s: (t.ms[i], t.calls[i]) for i, s in enumerate(STAGES)}

    def launch_count(self):
        return lib().sdyn_launch_count(self._h)

    def set_latency_mode(self, on):
        """One-frame contexts: replay the whole extraction call as one CUDA graph (default on)."""
        L = lib()
        L.sdyn_set_latency_mode.argtypes = [C.c_void_p, C.c_int]
        self._check(L.sdyn_set_latency_mode(self._h, int(bool(on))))

    def stream_handle(self):
        """The context's cudaStream_t as an integer (sdyn_stream)."""
        L = lib()
        L.sdyn_stream.restype = C.c_void_p
        L.sdyn_stream.argtypes = [C.c_void_p]
        return L.sdyn_stream(self._h) or 0


# ---------------------------------------------------------------------------------------------------
# ORBmatcher / dynamic-keypoint entry points
# ---------------------------------------------------------------------------------------------------
class FrameViewC(C.Structure):
    _fields_ = [("n", C.c_int32), ("nlevels", C.c_int32), ("keys", C.c_void_p), ("keys_un", C.c_void_p),
                ("desc", C.c_void_p), ("u_right", C.c_void_p), ("scale_factors", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("b", C.c_float), ("tcw", C.c_float * 12)]


class FeatureVectorC(C.Structure):
    _fields_ = [("nnodes", C.c_int32), ("node_id", C.c_void_p), ("offset", C.c_void_p), ("index", C.c_void_p)]


class BoxPairC(C.Structure):
    _fields_ = [("nq", C.c_int32), ("nt", C.c_int32), ("q_desc", C.c_void_p), ("t_desc", C.c_void_p),
                ("q_xy", C.c_void_p), ("t_xy", C.c_void_p), ("match_query", C.c_void_p), ("match_train", C.c_void_p),
                ("match_dist", C.c_void_p), ("false_dyn", C.c_void_p), ("nmatches", C.c_int32)]


MAPPOINT_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                           ("level", "<i4"), ("track_in_view", "u1"), ("bad", "u1"), ("obs_positive", "u1"),
                           ("pad", "u1"), ("desc", "u1", (32,))])
LASTPOINT_DTYPE = np.dtype([("has_mp", "u1"), ("outlier", "u1"), ("obs_positive", "u1"), ("pad", "u1"),
                            ("world", "<f4", (3,)), ("desc", "u1", (32,))])
assert MAPPOINT_DTYPE.itemsize == 56 and LASTPOINT_DTYPE.itemsize == 48
PROJPOINT_DTYPE = np.dtype([("valid", "u1"), ("pad", "u1", (3,)), ("world", "<f4", (3,)), ("normal", "<f4", (3,)),
                            ("min_distance", "<f4"), ("max_distance", "<f4"), ("max_distance_raw", "<f4"), ("angle", "<f4"),
                            ("desc", "u1", (32,))])
assert PROJPOINT_DTYPE.itemsize == 76
PROJ_FRAME_KEYFRAME, PROJ_KEYFRAME_SIM3 = 0, 1


class ProjParamsC(C.Structure):
    _fields_ = [("rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("ow", C.c_float * 3), ("th", C.c_float),
                ("max_descriptor_distance", C.c_int32), ("variant", C.c_int32), ("check_orientation", C.c_int32),
                ("log_scale_factor", C.c_float), ("nlevels", C.c_int32)]


class TriParamsC(C.Structure):
    _fields_ = [("f12", C.c_float * 9), ("epipole_x", C.c_float), ("epipole_y", C.c_float), ("only_stereo", C.c_int32),
                ("check_orientation", C.c_int32), ("level_sigma2", C.c_float * MAX_LEVELS)]


def tri_params(f12, epipole, only_stereo, check_orientation, level_sigma2):
    p = TriParamsC()
    for i, v in enumerate(np.asarray(f12, np.float32).reshape(9)):
        p.f12[i] = float(v)
    p.epipole_x, p.epipole_y = float(np.float32(epipole[0])), float(np.float32(epipole[1]))
    p.only_stereo = int(only_stereo); p.check_orientation = int(check_orientation)
    for i, v in enumerate(level_sigma2):
        p.level_sigma2[i] = float(np.float32(v))
    return p


class BestParamsC(C.Structure):
    _fields_ = [("t1", C.c_float * 12), ("t2", C.c_float * 12), ("use_t2", C.c_int32), ("ow", C.c_float * 3),
                ("invz_double", C.c_int32), ("dist_from_camera", C.c_int32), ("check_normal", C.c_int32), ("chi2_gate", C.c_int32),
                ("bf", C.c_float), ("inv_level_sigma2", C.c_float * MAX_LEVELS), ("th", C.c_float), ("log_scale_factor", C.c_float),
                ("nlevels", C.c_int32)]


def best_params(t1, th, log_scale_factor, nlevels, t2=None, ow=(0, 0, 0), invz_double=False, dist_from_camera=False,
                check_normal=False, chi2_gate=False, bf=0.0, inv_level_sigma2=None):
    p = BestParamsC()
    for i, v in enumerate(np.asarray(t1, np.float32).reshape(12)):
        p.t1[i] = float(v)
    if t2 is not None:
        for i, v in enumerate(np.asarray(t2, np.float32).reshape(12)):
            p.t2[i] = float(v)
        p.use_t2 = 1
    for i in range(3):
        p.ow[i] = float(np.float32(ow[i]))
    p.invz_double = int(invz_double); p.dist_from_camera = int(dist_from_camera); p.check_normal = int(check_normal)
    p.chi2_gate = int(chi2_gate); p.bf = float(np.float32(bf))
    if inv_level_sigma2 is not None:
        for i, v in enumerate(inv_level_sigma2):
            p.inv_level_sigma2[i] = float(np.float32(v))
    p.th = th; p.log_scale_factor = float(np.float32(log_scale_factor)); p.nlevels = nlevels
    return p


def proj_params(rcw, tcw, ow, th, max_dist, variant, check_orientation, log_scale_factor, nlevels):
    p = ProjParamsC()
    for i, v in enumerate(np.asarray(rcw, np.float32).reshape(9)):
        p.rcw[i] = float(v)
    for i in range(3):
        p.tcw[i] = float(np.float32(tcw[i])); p.ow[i] = float(np.float32(ow[i]))
    p.th = th; p.max_descriptor_distance = max_dist; p.variant = variant; p.check_orientation = int(check_orientation)
    p.log_scale_factor = float(np.float32(log_scale_factor)); p.nlevels = nlevels
    return p


class FrameView:
    """The Frame members the searches read (include/Frame.h), kept alive next to their C struct."""

    def __init__(self, keys, desc, scale_factors, bounds, keys_un=None, u_right=None,
                 cam=(0., 0., 0., 0., 0., 0.), tcw=None):
        self.keys = np.ascontiguousarray(keys, KP_DTYPE)
        self.keys_un = self.keys if keys_un is None else np.ascontiguousarray(keys_un, KP_DTYPE)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.u_right = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        self.scale = np.ascontiguousarray(scale_factors, np.float32)
        self.bounds = tuple(float(v) for v in bounds)          # mnMinX, mnMinY, mnMaxX, mnMaxY
        self.cam = tuple(float(v) for v in cam)                # fx, fy, cx, cy, mbf, mb
        self.tcw = np.ascontiguousarray(np.eye(4, dtype=np.float32)[:3] if tcw is None else tcw, np.float32).reshape(12)
        c = FrameViewC()
        c.n = len(self.keys); c.nlevels = len(self.scale)
        c.keys = self.keys.ctypes.data; c.keys_un = self.keys_un.ctypes.data; c.desc = self.desc.ctypes.data
        c.u_right = self.u_right.ctypes.data if self.u_right is not None else None
        c.scale_factors = self.scale.ctypes.data
        c.min_x, c.min_y, c.max_x, c.max_y = self.bounds
        c.fx, c.fy, c.cx, c.cy, c.bf, c.b = self.cam
        for i in range(12):
            c.tcw[i] = float(self.tcw[i])
        self.c = c

    @property
    def n(self):
        return len(self.keys)


class FeatureVector:
    """DBoW2::FeatureVector as CSR (ascending node ids)."""

    def __init__(self, node_of_feature):
        node_of_feature = np.asarray(node_of_feature)
        order = np.argsort(node_of_feature, kind="stable")
        ids, counts = np.unique(node_of_feature, return_counts=True)
        self.node_id = np.ascontiguousarray(ids, np.uint32)
        self.offset = np.ascontiguousarray(np.concatenate([[0], np.cumsum(counts)]), np.int32)
        self.index = np.ascontiguousarray(order, np.uint32)
        c = FeatureVectorC()
        c.nnodes = len(ids); c.node_id = self.node_id.ctypes.data; c.offset = self.offset.ctypes.data
        c.index = self.index.ctypes.data
        self.c = c


def _bind_match(L):
    if getattr(L, "_match_bound", False):
        return L
    vp = C.c_void_p
    L.sdyn_hamming.argtypes = [vp, vp]
    L.sdyn_match_projection_map.argtypes = [vp, C.POINTER(FrameViewC), vp, C.c_int, C.c_float, C.c_float, vp, vp,
                                            C.POINTER(C.c_int)]
    L.sdyn_match_projection_frame.argtypes = [vp, C.POINTER(FrameViewC), C.POINTER(FrameViewC), vp, C.c_float, C.c_int,
                                              C.c_int, vp, vp, C.POINTER(C.c_int), vp, C.POINTER(C.c_int)]
    L.sdyn_match_init.argtypes = [vp, C.POINTER(FrameViewC), C.POINTER(FrameViewC), vp, vp, C.c_int, C.c_float, C.c_int,
                                  C.POINTER(C.c_int)]
    L.sdyn_match_bow.argtypes = [vp, C.POINTER(FrameViewC), vp, C.POINTER(FeatureVectorC), C.POINTER(FrameViewC),
                                 C.POINTER(FeatureVectorC), C.c_float, C.c_int, vp, C.POINTER(C.c_int)]
    L.sdyn_dyn_box_mask.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp]
    L.sdyn_dyn_separate.argtypes = [vp, C.POINTER(BoxPairC), C.c_int, vp, C.c_int]
    L._match_bound = True
    return L


class Matcher:
    """Mirror of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:41-95) over the C ABI.  Pointers become indices:
    mvpMapPoints -> `assign` (-1 = NULL), "occupant has Observations()>0" -> `locked`."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30

    def __init__(self, ctx, nnratio=0.6, check_orientation=True):
        self.ex = ctx                       # an Extractor: owns the sdyn context (device, stream, arenas)
        self.h = ctx._h
        self.nnratio = float(nnratio)
        self.check = bool(check_orientation)
        self.L = _bind_match(lib())

    @staticmethod
    def DescriptorDistance(a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return _bind_match(lib()).sdyn_hamming(a.ctypes.data, b.ctypes.data)

    def SearchByProjectionMap(self, F, mappoints, th, assign=None, locked=None):
        mp = np.ascontiguousarray(mappoints, MAPPOINT_DTYPE)
        assign = np.full(F.n, -1, np.int32) if assign is None else np.ascontiguousarray(assign, np.int32).copy()
        locked = np.zeros(F.n, np.uint8) if locked is None else np.ascontiguousarray(locked, np.uint8).copy()
        n = C.c_int(0)
        self.ex._check(self.L.sdyn_match_projection_map(self.h, C.byref(F.c), mp.ctypes.data, len(mp), th, self.nnratio,
                                                        assign.ctypes.data, locked.ctypes.data, C.byref(n)))
        return n.value, assign, locked

    def SearchByProjectionFrame(self, cur, last, last_points, th, mono, assign=None, locked=None, want_pairs=False):
        lp = np.ascontiguousarray(last_points, LASTPOINT_DTYPE)
        assign = np.full(cur.n, -1, np.int32) if assign is None else np.ascontiguousarray(assign, np.int32).copy()
        locked = np.zeros(cur.n, np.uint8) if locked is None else np.ascontiguousarray(locked, np.uint8).copy()
        n, npairs = C.c_int(0), C.c_int(0)
        pairs = np.zeros((max(last.n, 1), 4), np.float32) if want_pairs else None
        self.ex._check(self.L.sdyn_match_projection_frame(
            self.h, C.byref(cur.c), C.byref(last.c), lp.ctypes.data, th, int(mono), int(self.check), assign.ctypes.data,
            locked.ctypes.data, C.byref(n), pairs.ctypes.data if want_pairs else None, C.byref(npairs) if want_pairs else None))
        if want_pairs:
            return n.value, assign, locked, pairs[:npairs.value].copy()
        return n.value, assign, locked

    def SearchByProjectionPose(self, target, points, params, assign=None):
        """The two pose-projection overloads (ORBmatcher.cc:290-403, 1629-1756): (nmatches, assign)."""
        pts = np.ascontiguousarray(points, PROJPOINT_DTYPE)
        assign = np.full(target.n, -1, np.int32) if assign is None else np.ascontiguousarray(assign, np.int32).copy()
        n = C.c_int(0)
        vp = C.c_void_p
        self.L.sdyn_match_projection_pose.argtypes = [vp, C.POINTER(FrameViewC), vp, C.c_int, C.POINTER(ProjParamsC), vp,
                                                      C.POINTER(C.c_int)]
        self.ex._check(self.L.sdyn_match_projection_pose(self.h, C.byref(target.c), pts.ctypes.data, len(pts), C.byref(params),
                                                         assign.ctypes.data, C.byref(n)))
        return n.value, assign

    def ProjectionBest(self, target, points, params):
        """Independent best keypoint per projected MapPoint (Fuse x2, SearchBySim3 passes): (best_idx, best_dist)."""
        pts = np.ascontiguousarray(points, PROJPOINT_DTYPE)
        bi = np.full(max(len(pts), 1), -1, np.int32); bd = np.full(max(len(pts), 1), 256, np.int32)
        vp = C.c_void_p
        self.L.sdyn_match_projection_best.argtypes = [vp, C.POINTER(FrameViewC), vp, C.c_int, C.POINTER(BestParamsC), vp, vp]
        self.ex._check(self.L.sdyn_match_projection_best(self.h, C.byref(target.c), pts.ctypes.data, len(pts), C.byref(params),
                                                         bi.ctypes.data, bd.ctypes.data))
        return bi[:len(pts)], bd[:len(pts)]

    def SearchBySim3(self, KF1, KF2, pts1, pts2, T1w, T2w, S12, S21, th, log_sf, nlevels):
        """ORBmatcher::SearchBySim3 (:1259-1483): both projection passes on the device, TH_HIGH and the agreement check here."""
        i12, d12 = self.ProjectionBest(KF2, pts1, best_params(T1w, th, log_sf, nlevels, t2=S21, invz_double=True, dist_from_camera=True))
        i21, d21 = self.ProjectionBest(KF1, pts2, best_params(T2w, th, log_sf, nlevels, t2=S12, invz_double=True, dist_from_camera=True))
        m1 = np.where(d12 <= 100, i12, -1); m2 = np.where(d21 <= 100, i21, -1)
        out = np.full(KF1.n, -1, np.int32)
        for i1 in range(KF1.n):
            if m1[i1] >= 0 and m2[m1[i1]] == i1:
                out[i1] = m1[i1]
        return int((out >= 0).sum()), out

    def SearchForTriangulation(self, KF1, has_mp1, fv1, KF2, has_mp2, fv2, params):
        """ORBmatcher::SearchForTriangulation (:814-980): (nmatches, matches12)."""
        h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
        m12 = np.full(KF1.n, -1, np.int32); n = C.c_int(0)
        vp = C.c_void_p
        self.L.sdyn_match_triangulation.argtypes = [vp, C.POINTER(FrameViewC), vp, C.POINTER(FeatureVectorC), C.POINTER(FrameViewC), vp,
                                                    C.POINTER(FeatureVectorC), C.POINTER(TriParamsC), vp, C.POINTER(C.c_int)]
        self.ex._check(self.L.sdyn_match_triangulation(self.h, C.byref(KF1.c), h1.ctypes.data, C.byref(fv1.c), C.byref(KF2.c),
                                                       h2.ctypes.data, C.byref(fv2.c), C.byref(params), m12.ctypes.data, C.byref(n)))
        return n.value, m12

    def SearchForInitialization(self, F1, F2, prev_matched, window=100):
        prev = np.ascontiguousarray(prev_matched, np.float32).copy()
        m12 = np.full(F1.n, -1, np.int32)
        n = C.c_int(0)
        self.ex._check(self.L.sdyn_match_init(self.h, C.byref(F1.c), C.byref(F2.c), prev.ctypes.data, m12.ctypes.data,
                                              window, self.nnratio, int(self.check), C.byref(n)))
        return n.value, m12, prev

    def last_evals(self):
        """Hamming evaluations of the last search call (sdyn_match_last_evals)."""
        self.L.sdyn_match_last_evals.restype = C.c_longlong
        self.L.sdyn_match_last_evals.argtypes = [C.c_void_p]
        return int(self.L.sdyn_match_last_evals(self.h))

    def SearchByBoW(self, KF, kf_valid, fv_kf, F, fv_f):
        kv = np.ascontiguousarray(kf_valid, np.uint8)
        assign = np.full(F.n, -1, np.int32)
        n = C.c_int(0)
        self.ex._check(self.L.sdyn_match_bow(self.h, C.byref(KF.c), kv.ctypes.data, C.byref(fv_kf.c), C.byref(F.c),
                                             C.byref(fv_f.c), self.nnratio, int(self.check), assign.ctypes.data, C.byref(n)))
        return n.value, assign


    def SearchByBoWKF(self, KF1, valid1, fv1, KF2, valid2, fv2):
        """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12): (nmatches, matches12[i1] = KeyFrame-2 index or -1)."""
        v1 = np.ascontiguousarray(valid1, np.uint8); v2 = np.ascontiguousarray(valid2, np.uint8)
        m12 = np.full(KF1.n, -1, np.int32)
        n = C.c_int(0)
        vp = C.c_void_p
        self.L.sdyn_match_bow_kf.argtypes = [vp, C.POINTER(FrameViewC), vp, C.POINTER(FeatureVectorC), C.POINTER(FrameViewC), vp,
                                             C.POINTER(FeatureVectorC), C.c_float, C.c_int, vp, C.POINTER(C.c_int)]
        self.ex._check(self.L.sdyn_match_bow_kf(self.h, C.byref(KF1.c), v1.ctypes.data, C.byref(fv1.c), C.byref(KF2.c),
                                                v2.ctypes.data, C.byref(fv2.c), self.nnratio, int(self.check), m12.ctypes.data,
                                                C.byref(n)))
        return n.value, m12


def box_mask(ctx, keys, boxes):
    """Frame::firstSeparate's keypoint-in-box test: uint64 bitmask per keypoint."""
    L = _bind_match(lib())
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
    mask = np.zeros(len(keys), np.uint64)
    ctx._check(L.sdyn_dyn_box_mask(ctx._h, keys.ctypes.data, len(keys), boxes.ctypes.data, len(boxes), mask.ctypes.data))
    return mask


def separate_pairs(ctx, pairs, M, mode, fn=None):
    """pairs: list of (q_desc[nq,32], q_xy[nq,2], t_desc[nt,32], t_xy[nt,2]).  Returns per pair
    (query, train, dist, false_dyn) arrays — BFMatcher(crossCheck) + classifyF (mode 0) / classifyH (mode 1)."""
    L = _bind_match(lib())
    arr = (BoxPairC * max(len(pairs), 1))()
    keep = []
    for i, (qd, qx, td, tx) in enumerate(pairs):
        qd = np.ascontiguousarray(qd, np.uint8).reshape(-1, 32); td = np.ascontiguousarray(td, np.uint8).reshape(-1, 32)
        qx = np.ascontiguousarray(qx, np.float32).reshape(-1, 2); tx = np.ascontiguousarray(tx, np.float32).reshape(-1, 2)
        outs = [np.full(max(len(qd), 1), -9, np.int32) for _ in range(4)]
        keep.append((qd, qx, td, tx, outs))
        p = arr[i]
        p.nq, p.nt = len(qd), len(td)
        p.q_desc, p.t_desc, p.q_xy, p.t_xy = qd.ctypes.data, td.ctypes.data, qx.ctypes.data, tx.ctypes.data
        p.match_query, p.match_train, p.match_dist, p.false_dyn = (o.ctypes.data for o in outs)
    M = np.ascontiguousarray(M, np.float32).reshape(9)
    if fn is None:
        ctx._check(L.sdyn_dyn_separate(ctx._h, arr, len(pairs), M.ctypes.data, mode))
    else:
        fn(arr, len(pairs), M.ctypes.data, mode)
    res = []
    for i, (_, _, _, _, outs) in enumerate(keep):
        n = arr[i].nmatches
        res.append(tuple(o[:n].copy() for o in outs))
    return res


# ---------------------------------------------------------------------------------------------------
# batched, device-resident front end
# ---------------------------------------------------------------------------------------------------
class TrackInputsC(C.Structure):
    _fields_ = [("last_points", C.c_void_p), ("last_keys", C.c_void_p), ("last_keys_un", C.c_void_p),
                ("n_last", C.c_void_p), ("last_stride", C.c_int32),
                ("map_points", C.c_void_p), ("n_map", C.c_void_p), ("map_stride", C.c_int32),
                ("boxes", C.c_void_p), ("n_boxes", C.c_void_p), ("ref_box", C.c_void_p),
                ("ref_desc", C.c_void_p), ("ref_xy", C.c_void_p), ("ref_off", C.c_void_p), ("ref_stride", C.c_int32),
                ("fmat", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("b", C.c_float),
                ("th_frame", C.c_float), ("th_map", C.c_float), ("nnratio_map", C.c_float),
                ("mono", C.c_int32), ("check_orientation", C.c_int32),
                ("poses", C.c_void_p),
                ("map", C.c_void_p), ("last_ids", C.c_void_p), ("last_flags", C.c_void_p),
                ("map_ids", C.c_void_p), ("map_proj", C.c_void_p),
                ("map_flags", C.c_void_p), ("viewing_cos_limit", C.c_float),
                ("rgbd_split", C.c_int32), ("frame_pitch", C.c_int64)]


TRACK_ARRAYS = ["last_points", "last_keys", "last_keys_un", "n_last", "map_points", "n_map", "boxes", "n_boxes",
                "ref_box", "ref_desc", "ref_xy", "ref_off", "fmat", "poses", "last_ids", "last_flags", "map_ids", "map_proj", "map_flags"]
FORM_SEPARATE_KEYS_UN, FORM_RESIDENT_LAST, FORM_RESIDENT_MAP, FORM_DEVICE_FRUSTUM = 1, 2, 4, 8
LP_OUTLIER, LP_OBS_POSITIVE = 1, 2
MP_BAD, MP_OBS_POSITIVE, MP_SKIP = 1, 2, 4
MAP_POINT_DTYPE = np.dtype([("world", "<f4", (3,)), ("normal", "<f4", (3,)), ("min_distance", "<f4"), ("max_distance", "<f4"),
                            ("desc", "u1", (32,))])
MAP_PROJ_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                           ("track_in_view", "u1"), ("bad", "u1"), ("obs_positive", "u1"), ("pad", "u1")])
assert MAP_POINT_DTYPE.itemsize == 64 and MAP_PROJ_DTYPE.itemsize == 24


def track_inputs(ptrs, frame0, strides, params, map_table=None, rgbd_split=False, frame_pitch=0):
    """ptrs: {name: (base address, bytes per frame)}; frame0: first frame of the batch in those arrays.  Arrays missing from
    `ptrs` stay NULL (that is how the resident forms are selected: no last_points / map_points, but last_ids / map_ids).
    frame_pitch > 0: frame-major records (every array's frames are frame_pitch bytes apart)."""
    t = TrackInputsC()
    for name in TRACK_ARRAYS:
        if name in ptrs and ptrs[name] is not None:
            base, per_frame = ptrs[name]
            setattr(t, name, base + frame0 * (frame_pitch if frame_pitch else per_frame))
    t.frame_pitch = int(frame_pitch)
    t.viewing_cos_limit = 0.5            # Tracking::SearchLocalPoints: isInFrustum(pMP, 0.5)
    t.last_stride, t.map_stride, t.ref_stride = strides
    for k, v in params.items():
        if k in ("tcw_cur", "tcw_last"):
            continue                         # per-frame now: the "poses" array
        setattr(t, k, v)
    if map_table is not None:
        t.map = map_table._h
    t.rgbd_split = int(bool(rgbd_split))
    return t


def track_input_layout(nframes, strides, forms=0):
    """sdyn_track_input_layout: ({array name: offset}, block bytes) of the single-copy host staging block."""
    L = lib()
    L.sdyn_track_input_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t * len(TRACK_ARRAYS)),
                                          C.POINTER(C.c_size_t)]
    offs, total = (C.c_size_t * len(TRACK_ARRAYS))(), C.c_size_t()
    rc = L.sdyn_track_input_layout(nframes, strides[0], strides[1], strides[2], int(forms), C.byref(offs), C.byref(total))
    if rc != 0:
        raise SdynError(rc, "sdyn_track_input_layout: bad argument")
    return dict(zip(TRACK_ARRAYS, [int(o) for o in offs])), int(total.value)


def track_record_layout(strides, forms=0):
    """sdyn_track_record_layout: ({array name: offset inside one frame's record}, record bytes)."""
    L = lib()
    L.sdyn_track_record_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t * len(TRACK_ARRAYS)), C.POINTER(C.c_size_t)]
    offs, pitch = (C.c_size_t * len(TRACK_ARRAYS))(), C.c_size_t()
    rc = L.sdyn_track_record_layout(strides[0], strides[1], strides[2], int(forms), C.byref(offs), C.byref(pitch))
    if rc != 0:
        raise SdynError(rc, "sdyn_track_record_layout: bad argument")
    return dict(zip(TRACK_ARRAYS, [int(o) for o in offs])), int(pitch.value)


def pack_records(arrays, strides, forms, out=None):
    """Frame-major record pool from per-array host arrays ([nframes, ...] each): -> (uint8 [nframes, pitch], layout, pitch)."""
    layout, pitch = track_record_layout(strides, forms)
    n = len(next(iter(arrays.values())))
    pool = np.zeros((n, pitch), np.uint8) if out is None else out
    for name in TRACK_ARRAYS:
        if name not in arrays:
            continue
        rows = arrays[name].view(np.uint8).reshape(n, -1)
        pool[:, layout[name]:layout[name] + rows.shape[1]] = rows
    return pool, layout, pitch


class MapTable:
    """Device-resident MapPoint table (sdyn_map_*): id -> world position, normal, scale-invariance distances, descriptor."""

    def __init__(self, capacity, device=0):
        L = lib()
        L.sdyn_map_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.sdyn_map_destroy.argtypes = [C.c_void_p]
        L.sdyn_map_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        h = C.c_void_p()
        rc = L.sdyn_map_create(device, capacity, C.byref(h))
        if rc != 0:
            raise SdynError(rc, "sdyn_map_create")
        self._h, self.capacity = h, capacity

    def update(self, first_id, points, stream=None):
        pts = np.ascontiguousarray(points, MAP_POINT_DTYPE)
        rc = lib().sdyn_map_update(self._h, first_id, len(pts), pts.ctypes.data, C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise SdynError(rc, "sdyn_map_update")
        self._keep = pts                     # the copy is asynchronous on pinned memory

    def close(self):
        if self._h:
            lib().sdyn_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def track_frame_order(ex, nframes):
    """rgbd_split steps: (order [nframes, cap], N [nframes], N_s [nframes]) of the tracked keypoint lists."""
    L = lib()
    L.sdyn_track_frame_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    order = np.zeros((nframes, ex.cap), np.int32); n = np.zeros(nframes, np.int32); ns = np.zeros(nframes, np.int32)
    ex._check(L.sdyn_track_frame_order(ex._h, nframes, order.ctypes.data, n.ctypes.data, ns.ctypes.data, ex.cap))
    return order, n, ns


def _bind_track(L):
    if getattr(L, "_track_bound", False):
        return L
    vp = C.c_void_p
    L.sdyn_track_batch_device.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(TrackInputsC), vp]
    L.sdyn_track_fetch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp]
    L.sdyn_track_batch.argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(TrackInputsC),
                                   vp, vp, vp, vp, vp, vp, vp, C.c_int]
    L.sdyn_track_batch_async.argtypes = L.sdyn_track_batch.argtypes
    L.sdyn_track_wait.argtypes = [vp]
    L.sdyn_track_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_longlong * 2)]
    L._track_bound = True
    return L


def track_batch_device(ex, nframes, dptr, frame_stride, w, h, row_stride, tin, stream=None):
    L = _bind_track(lib())
    ex._check(L.sdyn_track_batch_device(ex._h, nframes, C.c_void_p(dptr), frame_stride, w, h, row_stride, C.byref(tin),
                                        C.c_void_p(stream) if stream else None))


def track_batch_stereo_device(left, right, nframes, dptr_left, dptr_right, frame_stride, w, h, row_stride, tin, mb, mbf):
    L = _bind_track(lib())
    vp = C.c_void_p
    L.sdyn_track_batch_stereo_device.argtypes = [vp, vp, C.c_int, vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                                 C.POINTER(TrackInputsC), C.c_float, C.c_float]
    left._check(L.sdyn_track_batch_stereo_device(left._h, right._h, nframes, C.c_void_p(dptr_left), C.c_void_p(dptr_right),
                                                 frame_stride, w, h, row_stride, C.byref(tin), mb, mbf))


def track_fetch(ex, nframes, stream=None, out=None):
    L = _bind_track(lib())
    if out is None:
        out = (np.zeros((nframes, ex.cap), np.int32), np.zeros((nframes, ex.cap), np.uint8),
               np.zeros((nframes, ex.cap), np.uint8), np.zeros((nframes, 4), np.int32))
    a, l, m, c = out
    ex._check(L.sdyn_track_fetch(ex._h, nframes, a.ctypes.data, l.ctypes.data, m.ctypes.data, c.ctypes.data, ex.cap,
                                 C.c_void_p(stream) if stream else None))
    return a, l, m, c


def track_batch_host(ex, images, tin, outs):
    """Host-buffer step: images [B,H,W] uint8 and `tin` pointing at HOST arrays (pinned for async copies).
    outs = (kps, desc, n, assign, locked, dyn_mask, counts) host arrays, filled in place."""
    L = _bind_track(lib())
    b, h, w = images.shape
    kps, desc, n, assign, locked, mask, counts = outs
    ex._check(L.sdyn_track_batch(ex._h, b, images.ctypes.data, images.strides[0], w, h, images.strides[1], C.byref(tin),
                                 kps.ctypes.data, desc.ctypes.data, n.ctypes.data, assign.ctypes.data, locked.ctypes.data,
                                 mask.ctypes.data, counts.ctypes.data, ex.cap))
    return outs


def track_batch_host_async(ex, images, tin, outs):
    """Asynchronous host-buffer step: enqueues H2D + kernels + D2H on the context's stream and returns;
    `outs` are valid after track_wait(ex)."""
    L = _bind_track(lib())
    b, h, w = images.shape
    kps, desc, n, assign, locked, mask, counts = outs
    ex._check(L.sdyn_track_batch_async(ex._h, b, images.ctypes.data, images.strides[0], w, h, images.strides[1], C.byref(tin),
                                       kps.ctypes.data, desc.ctypes.data, n.ctypes.data, assign.ctypes.data,
                                       locked.ctypes.data, mask.ctypes.data, counts.ctypes.data, ex.cap))


def track_wait(ex):
    ex._check(_bind_track(lib()).sdyn_track_wait(ex._h))


def track_stats(ex, nframes):
    """(frame-search, map-search) Hamming evaluations of the last fetched step, summed over its frames."""
    ev = (C.c_longlong * 2)()
    ex._check(_bind_track(lib()).sdyn_track_stats(ex._h, nframes, C.byref(ev)))
    return int(ev[0]), int(ev[1])


# ---------------------------------------------------------------------------------------------------
# Frame::ComputeStereoMatches (src/Frame.cc:874-1048)
# ---------------------------------------------------------------------------------------------------
class StereoViewC(C.Structure):
    _fields_ = [("u_right", C.c_void_p), ("depth", C.c_void_p), ("kept", C.c_void_p), ("cap", C.c_int32)]


def _bind_stereo(L):
    if getattr(L, "_stereo_bound", False):
        return L
    L.sdyn_stereo_match_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
    L.sdyn_stereo_results.argtypes = [C.c_void_p, C.POINTER(StereoViewC)]
    L.sdyn_stereo_fetch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.sdyn_stereo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p]
    L._stereo_bound = True
    return L


def stereo_match_device(left, right, nframes, mb, mbf, stream=None):
    """Enqueues ComputeStereoMatches for the frames the two Extractors processed last; results stay on the device."""
    left._check(_bind_stereo(lib()).sdyn_stereo_match_device(left._h, right._h, nframes, mb, mbf, stream))


def stereo_view(left):
    v = StereoViewC()
    left._check(_bind_stereo(lib()).sdyn_stereo_results(left._h, C.byref(v)))
    return v


def stereo_fetch(left, nframes, stream=None):
    cap = left.cap
    ur = np.empty((nframes, cap), np.float32); dp = np.empty((nframes, cap), np.float32)
    kept = np.empty(nframes, np.int32)
    left._check(_bind_stereo(lib()).sdyn_stereo_fetch(left._h, nframes, ur.ctypes.data, dp.ctypes.data, cap,
                                                      kept.ctypes.data, stream))
    return ur, dp, kept


def stereo_match(left, right, nframes, mb, mbf):
    """Host-output form: (mvuRight [nframes][cap], mvDepth [nframes][cap], kept [nframes])."""
    cap = left.cap
    ur = np.empty((nframes, cap), np.float32); dp = np.empty((nframes, cap), np.float32)
    kept = np.empty(nframes, np.int32)
    left._check(_bind_stereo(lib()).sdyn_stereo_match(left._h, right._h, nframes, mb, mbf, ur.ctypes.data,
                                                      dp.ctypes.data, cap, kept.ctypes.data))
    return ur, dp, kept


# ---------------------------------------------------------------------------------------------------
# detection boxes: file format and Frame::boxTrack (host code of the C ABI)
# ---------------------------------------------------------------------------------------------------
def boxes_parse(text, cap=64):
    """'id cx cy w h' lines -> [n,4] float64 cv::Rect2d rows (Examples/RGB-D/rgbd_my.cc:232-252)."""
    raw = text.encode() if isinstance(text, str) else bytes(text)
    out = np.zeros((cap, 4), np.float64); n = C.c_int(0)
    L = lib()
    L.sdyn_boxes_parse.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    rc = L.sdyn_boxes_parse(raw, len(raw), out.ctypes.data, cap, C.byref(n))
    if rc != 0:
        raise SdynError(rc, "sdyn_boxes_parse")
    return out[:n.value].copy()


def boxes_read(path, cap=64):
    out = np.zeros((cap, 4), np.float64); n = C.c_int(0)
    L = lib()
    L.sdyn_boxes_read.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    rc = L.sdyn_boxes_read(str(path).encode(), out.ctypes.data, cap, C.byref(n))
    if rc != 0:
        raise SdynError(rc, "sdyn_boxes_read")
    return out[:n.value].copy()


def box_track(boxes, last_objects, last_box_idx, last_omit, last_vel, img_w, img_h):
    """Frame::boxTrack. Returns (boxes, box_idx, omit, velocity)."""
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    lo = np.ascontiguousarray(last_objects, np.float64).reshape(-1, 4)
    cap = len(boxes) + len(lo) + 1
    b = np.zeros((cap, 4), np.float64); b[:len(boxes)] = boxes
    li = np.ascontiguousarray(last_box_idx, np.int32); lom = np.ascontiguousarray(last_omit, np.uint8)
    lv = np.ascontiguousarray(last_vel, np.float64).reshape(-1, 2)
    bi = np.zeros(cap, np.int32); om = np.zeros(cap, np.uint8); vel = np.zeros((cap, 2), np.float64); n = C.c_int(0)
    L = lib()
    vp = C.c_void_p
    L.sdyn_box_track.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.POINTER(C.c_int)]
    rc = L.sdyn_box_track(b.ctypes.data, len(boxes), cap, lo.ctypes.data, li.ctypes.data, lom.ctypes.data, lv.ctypes.data, len(lo),
                          img_w, img_h, bi.ctypes.data, om.ctypes.data, vel.ctypes.data, C.byref(n))
    if rc != 0:
        raise SdynError(rc, "sdyn_box_track")
    return b[:n.value].copy(), bi[:n.value].copy(), om[:n.value].copy(), vel[:n.value].copy()


# ---------------------------------------------------------------------------------------------------
# Frame::ComputeBoW -> DBoW2 TemplatedVocabulary::transform
# ---------------------------------------------------------------------------------------------------
def _bind_bow(L):
    if getattr(L, "_bow_bound", False):
        return L
    vp = C.c_void_p
    L.sdyn_vocab_create.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.sdyn_vocab_load_text.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.sdyn_vocab_destroy.argtypes = [vp]
    L.sdyn_vocab_info.argtypes = [vp, C.POINTER(C.c_int32 * 4)]
    L.sdyn_bow_transform.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp]
    L.sdyn_bow_transform_device.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.sdyn_bow_fetch.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp]
    L.sdyn_bow_assemble.argtypes = [vp, vp, vp, C.c_int, vp, vp, C.POINTER(C.c_int), vp, vp, vp, C.POINTER(C.c_int)]
    L._bow_bound = True
    return L


class Vocabulary:
    """A DBoW2 vocabulary tree resident on one GPU (ORBvocabulary)."""

    def __init__(self, parent=None, is_leaf=None, desc=None, weight=None, k=10, L=6, path=None, device=0):
        lb = _bind_bow(lib())
        self._h = C.c_void_p()
        if path is not None:
            rc = lb.sdyn_vocab_load_text(device, str(path).encode(), C.byref(self._h))
        else:
            p = np.ascontiguousarray(parent, np.int32); lf = np.ascontiguousarray(is_leaf, np.uint8)
            d = np.ascontiguousarray(desc, np.uint8); w = np.ascontiguousarray(weight, np.float64)
            rc = lb.sdyn_vocab_create(device, len(p), p.ctypes.data, lf.ctypes.data, d.ctypes.data, w.ctypes.data, k, L, C.byref(self._h))
        if rc != 0:
            raise SdynError(rc, "vocabulary could not be created")
        info = (C.c_int32 * 4)()
        lb.sdyn_vocab_info(self._h, C.byref(info))
        self.nnodes, self.nwords, self.k, self.L = (int(v) for v in info)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().sdyn_vocab_destroy(self._h); self._h = C.c_void_p()

    __del__ = close


def bow_assemble(word, weight, node):
    """transform(features, BowVector&, FeatureVector&, levelsup)'s containers from the per-feature arrays (host code)."""
    lb = _bind_bow(lib())
    word = np.ascontiguousarray(word, np.uint32); weight = np.ascontiguousarray(weight, np.float64); node = np.ascontiguousarray(node, np.uint32)
    n = len(word); m = max(n, 1)
    bi = np.zeros(m, np.uint32); bv = np.zeros(m, np.float64); fn = np.zeros(m, np.uint32)
    fo = np.zeros(m + 1, np.int32); fi = np.zeros(m, np.uint32); nw, nn = C.c_int(0), C.c_int(0)
    rc = lb.sdyn_bow_assemble(word.ctypes.data, weight.ctypes.data, node.ctypes.data, n, bi.ctypes.data, bv.ctypes.data, C.byref(nw),
                              fn.ctypes.data, fo.ctypes.data, fi.ctypes.data, C.byref(nn))
    if rc != 0:
        raise SdynError(rc, "sdyn_bow_assemble")
    return dict(bow_ids=bi[:nw.value], bow_values=bv[:nw.value], fv_nodes=fn[:nn.value], fv_offset=fo[:nn.value + 1],
                fv_index=fi[:fo[nn.value] if n else 0])


def bow_transform(ex, voc, desc, levelsup=4):
    """Frame::ComputeBoW for host descriptors: per-feature (word, weight, node) from the device + host-assembled containers."""
    lb = _bind_bow(lib())
    d = np.ascontiguousarray(desc, np.uint8); n = len(d); m = max(n, 1)
    word = np.zeros(m, np.uint32); w = np.zeros(m, np.float64); node = np.zeros(m, np.uint32)
    ex._check(lb.sdyn_bow_transform(ex._h, voc._h, d.ctypes.data, n, levelsup, word.ctypes.data, w.ctypes.data, node.ctypes.data))
    out = dict(word=word[:n], weight=w[:n], node=node[:n])
    out.update(bow_assemble(word[:n], w[:n], node[:n]))
    return out


def bow_transform_device(ex, voc, nframes, levelsup=4, stream=None):
    ex._check(_bind_bow(lib()).sdyn_bow_transform_device(ex._h, voc._h, nframes, levelsup, stream))


def bow_fetch(ex, nframes, stream=None):
    cap = ex.cap
    word = np.zeros((nframes, cap), np.uint32); w = np.zeros((nframes, cap), np.float64); node = np.zeros((nframes, cap), np.uint32)
    ex._check(_bind_bow(lib()).sdyn_bow_fetch(ex._h, nframes, word.ctypes.data, w.ctypes.data, node.ctypes.data, cap, stream))
    return word, w, node
