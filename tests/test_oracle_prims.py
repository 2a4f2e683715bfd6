"""Known-answer tests: the oracle's restated OpenCV primitives vs cv2 4.13.0.
These pin the oracle; the reference ships no tests of its own."""
import numpy as np
import pytest

import orc

cv2 = pytest.importorskip("cv2")
rng = np.random.default_rng(20261018)


def noise(h, w, smooth=False):
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    return cv2.GaussianBlur(a, (5, 5), 1.5) if smooth else a


@pytest.mark.parametrize("shape", [(376, 1241, 313, 1034), (480, 640, 400, 533), (313, 1034, 261, 862),
                                   (100, 120, 50, 60), (64, 64, 32, 32), (37, 53, 80, 91), (105, 346, 88, 288)])
def test_resize_linear(shape):
    h, w, dh, dw = shape
    a = noise(h, w)
    assert np.array_equal(cv2.resize(a, (dw, dh), interpolation=cv2.INTER_LINEAR), orc.resize_linear(a, dw, dh))


def test_resize_linear_4k():
    a = noise(2160, 3840)
    assert np.array_equal(cv2.resize(a, (3200, 1800), interpolation=cv2.INTER_LINEAR), orc.resize_linear(a, 3200, 1800))


@pytest.mark.parametrize("shape,b", [((50, 70), 19), ((5, 7), 4), ((376, 1241), 19), ((20, 20), 19)])
def test_border_reflect101(shape, b):
    a = noise(*shape)
    assert np.array_equal(cv2.copyMakeBorder(a, b, b, b, b, cv2.BORDER_REFLECT_101), orc.border_reflect101(a, b))


@pytest.mark.parametrize("shape", [(376, 1241), (480, 640), (50, 7), (9, 300), (105, 346)])
def test_gaussian_blur(shape):
    a = noise(*shape)
    ref = cv2.GaussianBlur(a.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, orc.gaussian_blur7(a))


@pytest.mark.parametrize("th", [7, 12, 20, 40])
@pytest.mark.parametrize("smooth", [True, False])
def test_fast_nms(th, smooth):
    a = noise(160, 230, smooth)
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True)
    ref = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(a, None)], np.int32).reshape(-1, 3)
    assert np.array_equal(ref, orc.fast_nms(a, th))          # same set, same raster order, same response


def test_fast_small_cells():
    """cv::FAST on the 36x36-ish cell views the extractor uses, incl. degenerate ones (< 7 px)."""
    a = noise(120, 120, True)
    for (h, w) in [(36, 36), (7, 7), (6, 30), (30, 6), (8, 41)]:
        cell = np.ascontiguousarray(a[3:3 + h, 5:5 + w])
        det = cv2.FastFeatureDetector_create(threshold=7, nonmaxSuppression=True)
        ref = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in det.detect(cell, None)], np.int32).reshape(-1, 3)
        assert np.array_equal(ref, orc.fast_nms(cell, 7))


def test_fast_score_is_threshold_independent():
    a = noise(90, 90, True)
    s = orc.fast_score_map(a)
    for th in (7, 20):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=False)
        pts = {(int(p.pt[0]), int(p.pt[1])) for p in det.detect(a, None)}
        mine = {(x, y) for y, x in zip(*np.nonzero(s >= th))}
        assert pts == mine


def test_fast_atan2():
    y = rng.integers(-200000, 200000, 20000).astype(np.float32)
    x = rng.integers(-200000, 200000, 20000).astype(np.float32)
    y[:50] = 0; x[25:75] = 0
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(ref, orc.fast_atan2(y, x))
    assert float(orc.fast_atan2(np.float32([1]), np.float32([1]))[0]) == pytest.approx(44.990455627, abs=1e-6)


def test_orientation_against_cv2_orb():
    """SURVEY A-8: cv2.ORB's IC angle (same orb.cpp ancestry) equals the oracle's on identical pixels."""
    import common
    img = common.frame("kitti", 0)
    orb = cv2.ORB_create(nfeatures=800, nlevels=1, edgeThreshold=19, patchSize=31, fastThreshold=20)
    kps = orb.detect(img, None)
    assert len(kps) > 100
    for p in kps:
        x, y = int(round(p.pt[0])), int(round(p.pt[1]))
        assert np.float32(p.angle) == orc.ic_angle(img, x, y)


def test_descriptor_sampling_against_cv2_orb():
    """SURVEY A-8: with the blur swapped for cv2.ORB's own (float sepFilter2D) the oracle's rotated-BRIEF
    sampling reproduces cv2.ORB.compute bit for bit — pins pattern, rotation and rounding."""
    import common
    img = common.frame("kitti", 0)
    orb = cv2.ORB_create(nfeatures=500, nlevels=1, edgeThreshold=31, patchSize=31, fastThreshold=20)
    kps = orb.detect(img, None)
    kps, ref = orb.compute(img, kps)
    g = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    blur = cv2.sepFilter2D(img, -1, g, g, borderType=cv2.BORDER_REFLECT_101)
    same = 0
    for p, d in zip(kps, ref):
        x, y = int(round(p.pt[0])), int(round(p.pt[1]))
        same += np.array_equal(d, orc.orb_descriptor(blur, x, y, p.angle))
    assert same == len(kps) and same > 100


CAMERAS = {   # Examples/RGB-D/TUM1.yaml, TUM2.yaml (5 coefficients), a 4-coefficient camera
    "tum1": (517.306408, 516.469215, 318.643040, 255.313989, [0.262383, -0.953104, -0.005358, 0.002628, 1.163314]),
    "tum2": (520.908620, 521.007327, 325.141442, 249.701764, [0.231222, -0.784899, -0.003257, -0.000105, 0.917205]),
    "k4": (458.654, 457.296, 367.215, 248.375, [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]),
}


@pytest.mark.parametrize("cam", sorted(CAMERAS))
def test_undistort_points_bitexact_vs_cv2(cam):
    """Frame::UndistortKeyPoints / ComputeImageBounds call cv::undistortPoints(mat, mat, mK, mDistCoef, Mat(), mK)
    (src/Frame.cc:812-872); the oracle's restatement must reproduce cv2 bit for bit."""
    fx, fy, cx, cy, d = CAMERAS[cam]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32); D = np.array(d, np.float32).reshape(-1, 1)
    r = np.random.default_rng(1)
    pts = np.concatenate([r.uniform(0, 640, (20000, 1)), r.uniform(0, 480, (20000, 1))], 1).astype(np.float32)
    pts = np.concatenate([pts, np.array([[0, 0], [640, 0], [0, 480], [640, 480], [-50, 900]], np.float32)])
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
    got = orc.undistort_points(pts, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), D.reshape(-1))
    assert np.array_equal(ref.view(np.uint32), got.view(np.uint32))
