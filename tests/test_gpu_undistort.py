"""Parity of the device-side keypoint undistortion (Frame::UndistortKeyPoints / ComputeImageBounds,
src/Frame.cc:812-872 -> cv::undistortPoints) against the oracle, which tests/test_oracle_prims.py pins to cv2."""
import numpy as np
import pytest

import common
import oracle_track
import orc
import pysdyn
import scenario
from test_oracle_prims import CAMERAS

pytestmark = pytest.mark.gpu


def f32cam(cam):
    fx, fy, cx, cy, d = CAMERAS[cam]
    return float(np.float32(fx)), float(np.float32(fy)), float(np.float32(cx)), float(np.float32(cy)), np.array(d, np.float32)


@pytest.mark.parametrize("cam", sorted(CAMERAS))
def test_points_and_bounds(cam):
    fx, fy, cx, cy, d = f32cam(cam)
    ex = pysdyn.Extractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    ex.set_camera(fx, fy, cx, cy, d)
    r = np.random.default_rng(2)
    pts = np.concatenate([r.uniform(-20, 660, (50000, 1)), r.uniform(-20, 500, (50000, 1))], 1).astype(np.float32)
    got = ex.undistort_points(pts)
    ref = orc.undistort_points(pts, fx, fy, cx, cy, d)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    c = orc.undistort_points(np.array([[0, 0], [640, 0], [0, 480], [640, 480]], np.float32), fx, fy, cx, cy, d)
    expect = (min(c[0, 0], c[2, 0]), min(c[0, 1], c[1, 1]), max(c[1, 0], c[3, 0]), max(c[2, 1], c[3, 1]))
    assert ex.image_bounds(640, 480) == tuple(float(v) for v in expect)
    ex.set_camera(fx, fy, cx, cy, [])                     # no distortion: identity and the plain image rectangle
    assert np.array_equal(ex.undistort_points(pts[:100]), pts[:100]) and ex.image_bounds(640, 480) == (0.0, 0.0, 640.0, 480.0)
    ex.close()


def test_extraction_produces_keys_un_and_track_searches_them():
    import torch
    cfg, B, cam = "tum", 2, "tum1"
    fx, fy, cx, cy, d = f32cam(cam)
    W, H, nrect, nf, ini, mn = common.CONFIGS[cfg]
    cid = common.CONFIG_ID[cfg]
    seq_seed = 1000 * cid + 7
    frames = np.stack([pysdyn.synth_frame(seq_seed, 1000 * cid + i, W, H, nrect, *scenario.sequence_offsets(i),
                                          scenario.sequence_time(i)) for i in range(0, 1 + B)])
    cpu = orc.Extractor(nf, 1.2, 8, ini, mn)
    kd = [cpu(im) for im in frames]
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    gpu.set_camera(fx, fy, cx, cy, d)
    bounds = gpu.image_bounds(W, H)
    last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
    arrays = scenario.build_track_batch(kd, seq_seed, 1, W, H, nrect, 8, last_stride, map_stride, ref_stride, n_map=1500, seed=3)
    for f in range(B):                                     # LastFrame.mvKeysUn of a distorted camera
        n0 = int(arrays["n_last"][f])
        un = orc.undistort_points(np.stack([arrays["last_keys"][f, :n0]["x"], arrays["last_keys"][f, :n0]["y"]], 1), fx, fy, cx, cy, d)
        arrays["last_keys_un"][f, :n0]["x"] = un[:, 0]; arrays["last_keys_un"][f, :n0]["y"] = un[:, 1]
    params = scenario.track_params(W, H)
    params.update(min_x=bounds[0], min_y=bounds[1], max_x=bounds[2], max_y=bounds[3])
    dev = {k: torch.from_numpy(v.view(np.uint8).reshape(v.shape[0], -1)).cuda() for k, v in arrays.items()}
    ptrs = {k: (t.data_ptr(), t.shape[1]) for k, t in dev.items()}
    tin = pysdyn.track_inputs(ptrs, 0, (last_stride, map_stride, ref_stride), params)
    dframes = torch.from_numpy(frames[1:]).cuda()
    pysdyn.track_batch_device(gpu, B, dframes.data_ptr(), W * H, W, H, W, tin)
    kps, desc, counts = gpu.fetch(B)
    kun = gpu.fetch_keypoints_un(B)
    assign, locked, mask, cnt = pysdyn.track_fetch(gpu, B)
    for f in range(B):
        k, dsc = kd[f + 1]
        n = counts[f]
        assert n == len(k) and np.array_equal(kps[f, :n]["x"], k["x"])          # mvKeys stay raw
        un = orc.undistort_points(np.stack([k["x"], k["y"]], 1), fx, fy, cx, cy, d)
        ku = k.copy(); ku["x"] = un[:, 0]; ku["y"] = un[:, 1]
        assert np.array_equal(kun[f, :n], ku) and np.abs(ku["x"] - k["x"]).max() > 1.0
        ea, el, em, ec = oracle_track.track_frame(k, dsc, cpu.scale, W, H, arrays, f, params, last_stride, keys_un=ku, bounds=bounds)
        assert np.array_equal(desc[f, :n], dsc)
        assert np.array_equal(cnt[f], ec) and np.array_equal(assign[f, :n], ea) and np.array_equal(locked[f, :n], el)
        assert np.array_equal(mask[f, :n], em)
    gpu.close()
