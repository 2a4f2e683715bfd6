"""Parity of the CUDA extractor (through the C ABI) against the CPU oracle — stage by stage.
Bars (BASELINE.json north_star): pyramid bytes, keypoints (x, y, octave, response) bit-exact; angles within
1e-3 degrees; >= 99.9 % of descriptors bit-identical."""
import numpy as np
import pytest

import common
import orc
import pysdyn

pytestmark = pytest.mark.gpu

ANGLE_TOL_DEG = 1e-3
DESC_MIN_IDENTICAL = 0.999


def make(cfg, batch=1):
    w, h, _, nf, ini, mn = common.CONFIGS[cfg]
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=w, max_height=h, max_batch=batch)
    cpu = orc.Extractor(nf, 1.2, 8, ini, mn)
    return gpu, cpu


def check_frame(k, d, ok, od, tag=""):
    assert len(k) == len(ok), (tag, len(k), len(ok))
    for name in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(k[name], ok[name]), (tag, name)
    assert np.max(np.abs(k["angle"].astype(np.float64) - ok["angle"])) <= ANGLE_TOL_DEG, tag
    same = (d == od).all(1)
    assert same.mean() >= DESC_MIN_IDENTICAL, (tag, same.mean())
    return same


@pytest.mark.parametrize("cfg", ["small", "tum", "kitti", "kitti_mono"])
def test_extract_matches_oracle(cfg):
    gpu, cpu = make(cfg)
    assert np.array_equal(gpu.mvScaleFactor, cpu.scale) and np.array_equal(gpu.mnFeaturesPerLevel, cpu.quota)
    assert np.array_equal(gpu.mvInvScaleFactor, cpu.inv_scale) and np.array_equal(gpu.mvLevelSigma2, cpu.sigma2)
    for idx in range(3):
        img = common.frame(cfg, idx)
        k, d = gpu(img)
        ok, od = cpu(img)
        for l in range(8):                                   # pyramid incl. the 19-px frame: bit-exact
            assert np.array_equal(gpu.level(0, l), cpu.level(l)), (cfg, idx, l)
        for l in range(8):                                   # FAST candidates after the per-cell fallback: same set
            a = gpu.candidates(0, l)
            b = cpu.candidates(l)
            a = a[np.lexsort((a[:, 0], a[:, 1]))]
            b = b[np.lexsort((b[:, 0], b[:, 1]))]
            assert np.array_equal(a, b), (cfg, idx, l, len(a), len(b))
        check_frame(k, d, ok, od, (cfg, idx))


def test_extract_batch_equals_single():
    gpu, cpu = make("tum", batch=4)
    imgs = np.stack([common.frame("tum", i) for i in range(4)])
    kb, db, nb = gpu.extract_batch(imgs)
    for i in range(4):
        ok, od = cpu(imgs[i])
        check_frame(kb[i, :nb[i]], db[i, :nb[i]], ok, od, i)


def test_strided_input_and_size_change():
    gpu, cpu = make("tum")
    big = np.zeros((480, 700), np.uint8)
    big[:, :640] = common.frame("tum", 5)
    view = big[:, :640]                                      # stride 700
    k, d = gpu(view)
    ok, od = cpu(np.ascontiguousarray(view))
    check_frame(k, d, ok, od)
    small = np.ascontiguousarray(common.frame("tum", 6)[:300, :400])   # geometry recomputed on the fly
    k, d = gpu(small)
    ok, od = cpu(small)
    check_frame(k, d, ok, od)


@pytest.mark.parametrize("w,h", [(333, 251), (517, 243), (1023, 301), (641, 479)])
def test_odd_sizes_tile_edges(w, h):
    """Sizes that are multiples of nothing: the last tile of every level hangs over the bordered extent, so the TMA box
    loads of k_fast / k_blur / k_resize / k_orient_describe run into their zero-filled out-of-range parts, and the pyramid
    (incl. frames) must still equal the oracle's byte for byte."""
    gpu = pysdyn.Extractor(700, 1.2, 8, 15, 7, max_width=w, max_height=h)
    cpu = orc.Extractor(700, 1.2, 8, 15, 7)
    img = np.ascontiguousarray(common.frame("kitti", 11)[:h, :w])
    k, d = gpu(img)
    ok, od = cpu(img)
    check_frame(k, d, ok, od, (w, h))
    for level in (0, 3, 7):
        assert np.array_equal(gpu.level(0, level), cpu.level(level)), (w, h, level)
    gpu.close()


def test_edge_cases():
    gpu, cpu = make("small")
    k, d = gpu(np.zeros((0, 0), np.uint8))                   # empty image: silent no-op (ORBextractor.cc:1046)
    assert len(k) == 0
    flat = np.full((240, 320), 128, np.uint8)                # no corners anywhere
    k, d = gpu(flat)
    assert len(k) == 0 and len(cpu(flat)[0]) == 0
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (240, 320), dtype=np.uint8) # corners everywhere: stresses candidate capacity
    k, d = gpu(noise)
    ok, od = cpu(noise)
    check_frame(k, d, ok, od)
    with pytest.raises(pysdyn.SdynError):                    # level 7 too small for the 30-px grid: error, not UB
        gpu(np.zeros((100, 100), np.uint8))


def test_4k_stress_frame():
    """3840x2160 / 8000 features (BASELINE.json config 5): octree with ~1700-leaf levels, 26k FAST cells."""
    gpu, cpu = make("4k")
    img = common.frame("4k", 0)
    k, d = gpu(img)
    ok, od = cpu(img)
    assert len(ok) > 7900
    for l in (0, 3, 7):
        assert np.array_equal(gpu.level(0, l), cpu.level(l))
    check_frame(k, d, ok, od, "4k")


def test_stereo_pair_two_contexts_concurrently():
    """Left and right extractors run on two threads in the reference (src/Frame.cc:151-154): two contexts, two
    streams, concurrent calls."""
    import threading
    w, h, _, nf, ini, mn = common.CONFIGS["kitti"]
    left, right = common.frame("kitti", 3), common.frame("kitti", 3, ox=11)      # disparity-like shift
    exs = [pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=w, max_height=h) for _ in range(2)]
    out = [None, None]

    def run(i, img):
        for _ in range(3):
            out[i] = exs[i](img)

    th = [threading.Thread(target=run, args=(i, im)) for i, im in enumerate((left, right))]
    [t.start() for t in th]
    [t.join() for t in th]
    cpu = orc.Extractor(nf, 1.2, 8, ini, mn)
    for i, im in enumerate((left, right)):
        ok, od = cpu(im)
        check_frame(out[i][0], out[i][1], ok, od, ("stereo", i))
    [e.close() for e in exs]


@pytest.mark.parametrize("nfeatures,scale,nlevels,ini,mn", [(800, 1.5, 5, 20, 7), (600, 2.0, 3, 15, 5), (1200, 1.1, 8, 20, 7),
                                                            (300, 1.2, 1, 20, 7)])
def test_other_orb_parameters(nfeatures, scale, nlevels, ini, mn):
    """ORBextractor settings other than the shipped YAMLs: coarse pyramids (scale 1.5 / 2.0: the resize kernel's
    run-time staging pitch), a fine one (1.1) and a single level."""
    W, H = 640, 480
    gpu = pysdyn.Extractor(nfeatures, scale, nlevels, ini, mn, max_width=W, max_height=H)
    cpu = orc.Extractor(nfeatures, scale, nlevels, ini, mn)
    assert np.array_equal(gpu.mvScaleFactor, cpu.scale) and np.array_equal(gpu.mnFeaturesPerLevel, cpu.quota)
    for idx in range(2):
        img = common.frame("tum", idx)
        k, d = gpu(img)
        ok, od = cpu(img)
        for l in range(nlevels):
            assert np.array_equal(gpu.level(0, l), cpu.level(l)), (scale, idx, l)
        check_frame(k, d, ok, od, (scale, nlevels, idx))
        assert len(ok) > 100
    gpu.close()


def test_latency_mode_graph_is_bit_identical():
    """One-frame contexts replay the call as a CUDA graph (sdyn_set_latency_mode): same outputs as the stream path, across image
    sizes (re-capture), strided inputs and a distorted camera (mvKeysUn kernel inside the graph)."""
    import time
    cfg = "tum"
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    g = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=1)
    s = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=1)
    s.set_latency_mode(False)
    for idx, crop in [(0, None), (1, None), (2, (400, 600)), (3, None), (4, (333, 301))]:
        img = common.frame(cfg, idx)
        if crop:
            img = img[:crop[0], :crop[1]]          # a view with a row stride: the staging copy handles it
        for rep in range(2):
            kg, dg = g(img); ks, ds = s(img)
            assert len(kg) > 100 and kg.tobytes() == ks.tobytes() and np.array_equal(dg, ds)
    ok, od = orc.Extractor(nf, 1.2, 8, ini, mn)(common.frame(cfg, 3))
    kg, dg = g(common.frame(cfg, 3))
    assert kg.tobytes() == ok.tobytes() and np.array_equal(dg, od)
    img = common.frame(cfg, 0)
    t = []
    for e in (g, s):
        for _ in range(5):
            e(img)
        t0 = time.perf_counter()
        for _ in range(30):
            e(img)
        t.append((time.perf_counter() - t0) / 30)
    print("one-frame extraction: graph %.3f ms, stream launches %.3f ms" % (1e3 * t[0], 1e3 * t[1]))
    g.close(); s.close()
