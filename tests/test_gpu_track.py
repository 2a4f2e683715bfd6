"""Parity of the batched device-resident front end (extract + 2 searches + dynamic mask) against the
oracle composition, frame by frame."""
import numpy as np
import pytest

import common
import oracle_track
import orc
import pysdyn
import scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,B", [("tum", 4), ("kitti", 3)])
def test_track_batch_matches_oracle(cfg, B):
    import torch
    W, H, nrect, nf, ini, mn = common.CONFIGS[cfg]
    cid = common.CONFIG_ID[cfg]
    seq_seed = 1000 * cid + 7
    first = 1
    frames = np.stack([pysdyn.synth_frame(seq_seed, 1000 * cid + i, W, H, nrect, *scenario.sequence_offsets(i),
                                          scenario.sequence_time(i))
                       for i in range(first - 1, first + B)])
    cpu = orc.Extractor(nf, 1.2, 8, ini, mn)
    kd = [cpu(im) for im in frames]
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
    arrays = scenario.build_track_batch(kd, seq_seed, first, W, H, nrect, 8, last_stride, map_stride, ref_stride,
                                        n_map=1500, seed=3)
    assert arrays["n_boxes"].max() >= 3
    params = scenario.track_params(W, H)
    dev = {k: torch.from_numpy(v.view(np.uint8).reshape(v.shape[0], -1)).cuda() for k, v in arrays.items()}
    ptrs = {k: (t.data_ptr(), t.shape[1]) for k, t in dev.items()}
    tin = pysdyn.track_inputs(ptrs, 0, (last_stride, map_stride, ref_stride), params)
    dframes = torch.from_numpy(frames[1:]).cuda()
    pysdyn.track_batch_device(gpu, B, dframes.data_ptr(), W * H, W, H, W, tin)
    kps, desc, counts = gpu.fetch(B)
    assign, locked, mask, cnt = pysdyn.track_fetch(gpu, B)
    total_masked = 0
    for f in range(B):
        k, d = kd[f + 1]
        n = counts[f]
        assert n == len(k) and np.array_equal(kps[f, :n]["x"], k["x"]) and np.array_equal(kps[f, :n]["octave"], k["octave"])
        ea, el, em, ec = oracle_track.track_frame(k, d, cpu.scale, W, H, arrays, f, params, last_stride)
        # the GPU run uses its own descriptors (>= 99.9 % identical); compare the match arrays only when the
        # frame's descriptors are bit-identical, which holds for these seeds
        assert np.array_equal(desc[f, :n], d), "descriptor mismatch would propagate into match indices"
        assert np.array_equal(cnt[f], ec), (f, cnt[f], ec)
        assert np.array_equal(assign[f, :n], ea) and np.array_equal(locked[f, :n], el)
        assert np.array_equal(mask[f, :n], em)
        total_masked += int(em.sum())
        assert ec[0] > 100 and ec[1] > 50
    assert total_masked > 0

    # host-buffer forms of the same step (sdyn_track_batch and its async + wait split) give identical results
    hptrs = {k: (v.ctypes.data, int(np.prod(v.shape[1:])) * v.dtype.itemsize) for k, v in arrays.items()}
    hin = pysdyn.track_inputs(hptrs, 0, (last_stride, map_stride, ref_stride), params)
    # ... and so does the single-copy form: every array inside one block at sdyn_track_input_layout's offsets
    separate = not np.array_equal(arrays["last_keys"], arrays["last_keys_un"])
    layout, total = pysdyn.track_input_layout(B, (last_stride, map_stride, ref_stride), pysdyn.FORM_SEPARATE_KEYS_UN if separate else 0)
    block = np.full(total, 0xA5, np.uint8)
    pptrs = {}
    for k, v in arrays.items():
        rows = v.view(np.uint8).reshape(v.shape[0], -1)[:B]
        if separate or k != "last_keys_un":
            block[layout[k]:layout[k] + rows.size] = rows.reshape(-1)
        pptrs[k] = (block.ctypes.data + layout[k], rows.shape[1])
    if not separate:
        pptrs["last_keys_un"] = pptrs["last_keys"]
    pin = pysdyn.track_inputs(pptrs, 0, (last_stride, map_stride, ref_stride), params)
    for fn in ("sync", "async", "packed"):
        outs = (np.zeros((B, gpu.cap), pysdyn.KP_DTYPE), np.zeros((B, gpu.cap, 32), np.uint8), np.zeros(B, np.int32),
                np.zeros((B, gpu.cap), np.int32), np.zeros((B, gpu.cap), np.uint8), np.zeros((B, gpu.cap), np.uint8),
                np.zeros((B, 4), np.int32))
        imgs = np.ascontiguousarray(frames[1:])
        if fn == "sync":
            pysdyn.track_batch_host(gpu, imgs, hin, outs)
        elif fn == "packed":
            pysdyn.track_batch_host(gpu, imgs, pin, outs)
        else:
            pysdyn.track_batch_host_async(gpu, imgs, hin, outs)
            pysdyn.track_wait(gpu)
        assert np.array_equal(outs[2], counts) and np.array_equal(outs[6], cnt)
        for f in range(B):
            n = counts[f]
            assert np.array_equal(outs[0][f, :n], kps[f, :n]) and np.array_equal(outs[1][f, :n], desc[f, :n])
            assert np.array_equal(outs[3][f, :n], assign[f, :n]) and np.array_equal(outs[4][f, :n], locked[f, :n])
            assert np.array_equal(outs[5][f, :n], mask[f, :n])
    gpu.close()


def test_track_batch_stereo_matches_oracle():
    """BASELINE config 3: stereo pairs — both extractions, ComputeStereoMatches and the two searches with their mvuRight
    gates in one device-resident step (sdyn_track_batch_stereo_device) against the oracle composition."""
    import torch
    cfg, B = "kitti", 2
    W, H, nrect, nf, ini, mn = common.CONFIGS[cfg]
    cid = common.CONFIG_ID[cfg]
    seq_seed = 1000 * cid + 7
    idx = list(range(0, 1 + B))
    left = np.stack([pysdyn.synth_frame(seq_seed, 1000 * cid + i, W, H, nrect, *scenario.sequence_offsets(i), scenario.sequence_time(i))
                     for i in idx])
    right = np.stack([pysdyn.synth_frame(seq_seed, 1000 * cid + 5000 + i, W, H, nrect, scenario.sequence_offsets(i)[0] + 9 + 2 * i,
                                         scenario.sequence_offsets(i)[1], scenario.sequence_time(i)) for i in idx[1:]])
    cpuL, cpuR = orc.Extractor(nf, 1.2, 8, ini, mn), orc.Extractor(nf, 1.2, 8, ini, mn)
    kd = [cpuL(im) for im in left]
    cam = scenario.KITTI_CAM
    mb, mbf = cam["bf"] / cam["fx"], cam["bf"]
    L = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    R = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    last_stride, map_stride, ref_stride = L.cap, 1500, 512
    arrays = scenario.build_track_batch(kd, seq_seed, 1, W, H, nrect, 8, last_stride, map_stride, ref_stride, n_map=1500, seed=3)
    params = scenario.track_params(W, H)
    dev = {k: torch.from_numpy(v.view(np.uint8).reshape(v.shape[0], -1)).cuda() for k, v in arrays.items()}
    ptrs = {k: (t.data_ptr(), t.shape[1]) for k, t in dev.items()}
    tin = pysdyn.track_inputs(ptrs, 0, (last_stride, map_stride, ref_stride), params)
    dl = torch.from_numpy(left[1:]).cuda(); dr = torch.from_numpy(right).cuda()
    for _ in range(2):
        pysdyn.track_batch_stereo_device(L, R, B, dl.data_ptr(), dr.data_ptr(), W * H, W, H, W, tin, mb, mbf)
    kps, desc, counts = L.fetch(B)
    ur, dp, kept = pysdyn.stereo_fetch(L, B)
    assign, locked, mask, cnt = pysdyn.track_fetch(L, B)
    for f in range(B):
        k, d = kd[f + 1]
        n = counts[f]
        kr, dr_ = cpuR(right[f])
        cpuL(left[f + 1])                                   # the oracle's left pyramid of THIS frame
        our, odp, okept = orc.stereo_matches(cpuL, cpuR, k, d, kr, dr_, mb, mbf)
        assert n == len(k) and np.array_equal(desc[f, :n], d)
        assert np.array_equal(ur[f, :n].view(np.uint32), our.view(np.uint32)) and kept[f] == okept and okept > 300
        ea, el, em, ec = oracle_track.track_frame(k, d, cpuL.scale, W, H, arrays, f, params, last_stride, u_right=our)
        assert np.array_equal(cnt[f], ec), (f, cnt[f], ec)
        assert np.array_equal(assign[f, :n], ea) and np.array_equal(locked[f, :n], el) and np.array_equal(mask[f, :n], em)
        # the gates must have mattered: the monocular result differs
        ma, _, _, mc = oracle_track.track_frame(k, d, cpuL.scale, W, H, arrays, f, params, last_stride)
        assert not np.array_equal(ma, ea)
    L.close(); R.close()


def test_track_partial_batch_and_single_search():
    """nframes < max_batch (the two searches go out as separate candidate launches) and steps with only one of the
    two searches (last_stride = 0 or map_stride = 0) give the same per-frame results as the full step."""
    import torch
    cfg, B, Bmax = "tum", 2, 4
    W, H, nrect, nf, ini, mn = common.CONFIGS[cfg]
    cid = common.CONFIG_ID[cfg]
    seq_seed = 1000 * cid + 7
    frames = np.stack([pysdyn.synth_frame(seq_seed, 1000 * cid + i, W, H, nrect, *scenario.sequence_offsets(i), scenario.sequence_time(i))
                       for i in range(0, 1 + B)])
    cpu = orc.Extractor(nf, 1.2, 8, ini, mn)
    kd = [cpu(im) for im in frames]
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=Bmax)
    last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
    arrays = scenario.build_track_batch(kd, seq_seed, 1, W, H, nrect, 8, last_stride, map_stride, ref_stride, n_map=1500, seed=3)
    params = scenario.track_params(W, H)
    dev = {k: torch.from_numpy(v.view(np.uint8).reshape(v.shape[0], -1)).cuda() for k, v in arrays.items()}
    ptrs = {k: (t.data_ptr(), t.shape[1]) for k, t in dev.items()}
    dframes = torch.from_numpy(frames[1:]).cuda()
    tin = pysdyn.track_inputs(ptrs, 0, (last_stride, map_stride, ref_stride), params)
    pysdyn.track_batch_device(gpu, B, dframes.data_ptr(), W * H, W, H, W, tin)
    kps, desc, counts = gpu.fetch(B)
    assign, locked, mask, cnt = pysdyn.track_fetch(gpu, B)
    for f in range(B):
        k, d = kd[f + 1]
        n = counts[f]
        ea, el, em, ec = oracle_track.track_frame(k, d, cpu.scale, W, H, arrays, f, params, last_stride)
        assert np.array_equal(cnt[f], ec) and np.array_equal(assign[f, :n], ea) and np.array_equal(locked[f, :n], el)
        assert np.array_equal(mask[f, :n], em)
    # map search only: the frame search is switched off by last_stride = 0
    tin2 = pysdyn.track_inputs(ptrs, 0, (0, map_stride, ref_stride), params)
    pysdyn.track_batch_device(gpu, B, dframes.data_ptr(), W * H, W, H, W, tin2)
    a2, l2, m2, c2 = pysdyn.track_fetch(gpu, B)
    for f in range(B):
        k, d = kd[f + 1]
        F = scenario.frame_view(k, d, cpu.scale, W, H)
        F = pysdyn.FrameView(k, d, cpu.scale, (0.0, 0.0, float(W), float(H)),
                             cam=(scenario.KITTI_CAM["fx"], scenario.KITTI_CAM["fy"], scenario.KITTI_CAM["cx"], scenario.KITTI_CAM["cy"],
                                  scenario.KITTI_CAM["bf"], scenario.KITTI_CAM["bf"] / scenario.KITTI_CAM["fx"]), tcw=arrays["poses"][f, :12])
        nm = int(arrays["n_map"][f])
        n2, ea, el = orc.match_projection_map(F, arrays["map_points"][f, :nm], params["th_map"], params["nnratio_map"], assign_base=0)
        assert c2[f, 0] == 0 and c2[f, 1] == n2 and np.array_equal(a2[f, :len(k)], ea) and np.array_equal(l2[f, :len(k)], el)
    gpu.close()


def _sequence(cfg, count):
    W, H, nrect, nf, ini, mn = common.CONFIGS[cfg]
    cid = common.CONFIG_ID[cfg]
    seq_seed = 1000 * cid + 7
    frames = np.stack([pysdyn.synth_frame(seq_seed, 1000 * cid + i, W, H, nrect, *scenario.sequence_offsets(i), scenario.sequence_time(i))
                       for i in range(count)])
    cpu = orc.Extractor(nf, 1.2, 8, ini, mn)
    return W, H, nrect, nf, ini, mn, seq_seed, frames, cpu, [cpu(im) for im in frames]


def _dev(arrays):
    import torch
    dev = {k: torch.from_numpy(v.view(np.uint8).reshape(v.shape[0], -1)).cuda() for k, v in arrays.items()}
    return dev, {k: (t.data_ptr(), t.shape[1]) for k, t in dev.items()}


def test_per_frame_poses_change_the_level_rule():
    """Every frame of a batch carries its own pose pair: the forward / backward / neutral branch of ORBmatcher.cc:1505-1506 is
    taken per frame on the device (one pose per batch cannot express independent sequences)."""
    poses = np.stack([scenario.frame_pose(i) for i in range(1, 5)])
    tz = poses[:, 11]
    assert (tz > 1).any() and (tz < -1).any() and (np.abs(tz) < 0.5).any()


@pytest.mark.parametrize("cfg,B", [("tum", 3)])
def test_track_resident_forms_match_explicit(cfg, B):
    """Resident LastFrame (this slot's previous step) + resident MapPoint table + per-frame ids / flags / projection records
    give the results of the explicit arrays — device pointers and the host-buffer single-block upload."""
    import torch
    W, H, nrect, nf, ini, mn, seq_seed, frames, cpu, kd = _sequence(cfg, B + 2)
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
    strides = (last_stride, map_stride, ref_stride)
    params = scenario.track_params(W, H)
    a1 = scenario.build_track_batch(kd[:B + 1], seq_seed, 1, W, H, nrect, 8, *strides, n_map=1500, seed=3)
    a2 = scenario.build_track_batch(kd[1:B + 2], seq_seed, 2, W, H, nrect, 8, *strides, n_map=1500, seed=3)
    d1, p1 = _dev(a1); d2, p2 = _dev(a2)
    f1 = torch.from_numpy(frames[1:B + 1]).cuda(); f2 = torch.from_numpy(frames[2:B + 2]).cuda()
    # explicit reference run of step 2
    pysdyn.track_batch_device(gpu, B, f2.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p2, 0, strides, params))
    want = [x.copy() for x in pysdyn.track_fetch(gpu, B)]
    for f in range(B):                    # ... which is the oracle's result
        k, d = kd[f + 2]
        ea, el, em, ec = oracle_track.track_frame(k, d, cpu.scale, W, H, a2, f, params, last_stride)
        assert np.array_equal(want[3][f], ec) and np.array_equal(want[0][f, :len(k)], ea) and np.array_equal(want[2][f, :len(k)], em)
    # step 1 (explicit) leaves every slot's keypoints resident; step 2 then names its LastFrame points by id
    table, res = scenario.resident_forms(a2)
    mt = pysdyn.MapTable(len(table)); mt.update(0, table)
    torch.cuda.synchronize()
    dr, pr = _dev(res)
    rp = {k: v for k, v in p2.items() if k not in ("last_points", "last_keys", "last_keys_un", "n_last", "map_points")}
    rp.update(pr)
    pysdyn.track_batch_device(gpu, B, f1.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p1, 0, strides, params))
    pysdyn.track_batch_device(gpu, B, f2.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(rp, 0, strides, params, map_table=mt))
    got = pysdyn.track_fetch(gpu, B)
    assert np.array_equal(got[3], want[3])
    for f in range(B):                    # (entries past a frame's keypoint count are unspecified)
        n = len(kd[f + 2][0])
        for a, b in zip(got[:3], want[:3]):
            assert np.array_equal(a[f, :n], b[f, :n])
    assert want[3][:, 0].min() > 100 and want[3][:, 1].min() > 50
    # host-buffer form: one pinned block with the resident layout (ids, flags, projection records, poses, boxes ...)
    forms = pysdyn.FORM_RESIDENT_LAST | pysdyn.FORM_RESIDENT_MAP
    layout, total = pysdyn.track_input_layout(B, strides, forms)
    block = np.full(total, 0xA5, np.uint8)
    hp = {}
    host = dict(a2); host.update(res)
    for name in pysdyn.TRACK_ARRAYS:
        if name in ("last_points", "last_keys", "last_keys_un", "n_last", "map_points") or name not in host:
            continue
        rows = host[name].view(np.uint8).reshape(host[name].shape[0], -1)[:B]
        block[layout[name]:layout[name] + rows.size] = rows.reshape(-1)
        hp[name] = (block.ctypes.data + layout[name], rows.shape[1])
    outs = (np.zeros((B, gpu.cap), pysdyn.KP_DTYPE), np.zeros((B, gpu.cap, 32), np.uint8), np.zeros(B, np.int32),
            np.zeros((B, gpu.cap), np.int32), np.zeros((B, gpu.cap), np.uint8), np.zeros((B, gpu.cap), np.uint8), np.zeros((B, 4), np.int32))
    pysdyn.track_batch_device(gpu, B, f1.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p1, 0, strides, params))
    pysdyn.track_batch_host(gpu, np.ascontiguousarray(frames[2:B + 2]), pysdyn.track_inputs(hp, 0, strides, params, map_table=mt), outs)
    assert np.array_equal(outs[6], want[3])
    for f in range(B):
        n = outs[2][f]
        assert np.array_equal(outs[3][f, :n], want[0][f, :n]) and np.array_equal(outs[5][f, :n], want[2][f, :n])
    gpu.close()
    # the resident LastFrame survives a re-allocation of the track state: step 1 runs without a map search (state sized for
    # `cap` queries), step 2 brings 1500 map points (> cap: the state grows) and still finds its LastFrame on the device
    gpu2 = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    assert map_stride > gpu2.cap
    pysdyn.track_batch_device(gpu2, B, f1.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p1, 0, (last_stride, 0, ref_stride), params))
    pysdyn.track_batch_device(gpu2, B, f2.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(rp, 0, strides, params, map_table=mt))
    got2 = pysdyn.track_fetch(gpu2, B)
    assert np.array_equal(got2[3], want[3])
    for f in range(B):
        n = len(kd[f + 2][0])
        for a, b in zip(got2[:3], want[:3]):
            assert np.array_equal(a[f, :n], b[f, :n])
    mt.close(); gpu2.close()


@pytest.mark.parametrize("cfg,B,cam", [("tum", 3, None), ("tum", 2, "tum1")])
def test_track_rgbd_split_matches_oracle(cfg, B, cam):
    """BASELINE config 2 as the reference runs it (RGB-D constructor): firstSeparate moves the in-box keypoints out of the frame,
    Separate + UpdateFrame re-admit the static ones, and the searches run on that list."""
    import torch
    W, H, nrect, nf, ini, mn, seq_seed, frames, cpu, kd = _sequence(cfg, B + 1)
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    bounds = (0.0, 0.0, float(W), float(H))
    und = lambda k: k
    if cam:
        fx, fy, cx, cy, dist = np.float32(517.306408), np.float32(516.469215), np.float32(318.643040), np.float32(255.313989), \
            np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)
        gpu.set_camera(fx, fy, cx, cy, dist)
        bounds = gpu.image_bounds(W, H)

        def und(k):
            xy = orc.undistort_points(np.stack([k["x"], k["y"]], 1), fx, fy, cx, cy, dist)
            ku = k.copy(); ku["x"] = xy[:, 0]; ku["y"] = xy[:, 1]
            return ku
    last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
    strides = (last_stride, map_stride, ref_stride)
    arrays = scenario.build_track_batch(kd, seq_seed, 1, W, H, nrect, 8, *strides, n_map=1500, seed=3)
    for f in range(B):
        n0 = int(arrays["n_last"][f])
        arrays["last_keys_un"][f, :n0] = und(arrays["last_keys"][f, :n0])
    params = scenario.track_params(W, H)
    params.update(min_x=bounds[0], min_y=bounds[1], max_x=bounds[2], max_y=bounds[3])
    dev, ptrs = _dev(arrays)
    dframes = torch.from_numpy(frames[1:]).cuda()
    pysdyn.track_batch_device(gpu, B, dframes.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(ptrs, 0, strides, params, rgbd_split=True))
    kps, desc, counts = gpu.fetch(B)
    assign, locked, mask, cnt = pysdyn.track_fetch(gpu, B)
    order, n_all, n_static = pysdyn.track_frame_order(gpu, B)
    readmitted = 0
    for f in range(B):
        k, d = kd[f + 1]
        eo, ens, ea, el, em, ec, _ = oracle_track.track_frame_rgbd(k, d, cpu.scale, W, H, arrays, f, params, last_stride, keys_un=und(k),
                                                                   bounds=bounds)
        assert n_all[f] == len(eo) and n_static[f] == ens and np.array_equal(order[f, :len(eo)], eo)
        assert np.array_equal(cnt[f], ec), (f, cnt[f], ec)
        assert np.array_equal(assign[f, :len(eo)], ea) and np.array_equal(locked[f, :len(eo)], el)
        assert np.array_equal(mask[f, :len(k)], em)
        assert ens < len(k) and ec[0] > 80
        readmitted += len(eo) - ens
        # the split must matter: the stereo-constructor result differs
        sa, _, _, sc = oracle_track.track_frame(k, d, cpu.scale, W, H, arrays, f, params, last_stride, keys_un=und(k), bounds=bounds)
        assert not np.array_equal(sc[:2], ec[:2]) or len(sa) != len(ea)
    assert readmitted > 0
    gpu.close()


def test_track_rgbd_split_with_resident_last():
    """Two consecutive rgbd_split steps of the same slots: the list step 1 tracked with (static + re-admitted keypoints) IS
    step 2's LastFrame, resident on the device, and step 2 names its MapPoints by id in that list's order."""
    import torch
    cfg, B = "tum", 3
    W, H, nrect, nf, ini, mn, seq_seed, frames, cpu, kd = _sequence(cfg, B + 2)
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
    strides = (last_stride, map_stride, ref_stride)
    params = scenario.track_params(W, H)
    a1 = scenario.build_track_batch(kd[:B + 1], seq_seed, 1, W, H, nrect, 8, *strides, n_map=1500, seed=3)
    a2 = scenario.build_track_batch(kd[1:B + 2], seq_seed, 2, W, H, nrect, 8, *strides, n_map=1500, seed=3)
    d1, p1 = _dev(a1)
    f1 = torch.from_numpy(frames[1:B + 1]).cuda(); f2 = torch.from_numpy(frames[2:B + 2]).cuda()
    pysdyn.track_batch_device(gpu, B, f1.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p1, 0, strides, params, rgbd_split=True))
    order1, n_all1, n_static1 = pysdyn.track_frame_order(gpu, B)
    shorter = 0
    for f in range(B):
        eo = oracle_track.track_frame_rgbd(kd[f + 1][0], kd[f + 1][1], cpu.scale, W, H, a1, f, params, last_stride)[0]
        assert np.array_equal(order1[f, :n_all1[f]], eo)
        scenario.reorder_last(a2, f, order1[f, :n_all1[f]])
        shorter += int(n_all1[f] < len(kd[f + 1][0]))
    assert shorter > 0                        # the tracked list really differs from the extraction list
    table, res = scenario.resident_forms(a2)
    mt = pysdyn.MapTable(len(table)); mt.update(0, table)
    torch.cuda.synchronize()
    d2, p2 = _dev(a2); dr, pr = _dev(res)
    rp = {k: v for k, v in p2.items() if k not in ("last_points", "last_keys", "last_keys_un", "n_last", "map_points")}
    rp.update(pr)
    pysdyn.track_batch_device(gpu, B, f2.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(rp, 0, strides, params, map_table=mt, rgbd_split=True))
    assign, locked, mask, cnt = pysdyn.track_fetch(gpu, B)
    order2, n_all2, n_static2 = pysdyn.track_frame_order(gpu, B)
    for f in range(B):
        k, d = kd[f + 2]
        eo, ens, ea, el, em, ec, _ = oracle_track.track_frame_rgbd(k, d, cpu.scale, W, H, a2, f, params, last_stride)
        assert n_all2[f] == len(eo) and n_static2[f] == ens and np.array_equal(order2[f, :len(eo)], eo)
        assert np.array_equal(cnt[f], ec), (f, cnt[f], ec)
        assert np.array_equal(assign[f, :len(eo)], ea) and np.array_equal(locked[f, :len(eo)], el)
        assert np.array_equal(mask[f, :len(k)], em) and ec[0] > 80
    mt.close(); gpu.close()


def test_track_pool_overflow_keeps_last_frame_and_repeat_succeeds():
    """A search window wide enough to exhaust the candidate pool: the step reports SDYN_ERR_CAPACITY, the pool is doubled, NO
    slot's resident LastFrame has advanced, and repeating the step gives the oracle's result."""
    import torch
    cfg, B = "tum", 2
    W, H, nrect, nf, ini, mn, seq_seed, frames, cpu, kd = _sequence(cfg, B + 2)
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    last_stride, map_stride, ref_stride = gpu.cap, 1500, 512
    strides = (last_stride, map_stride, ref_stride)
    params = scenario.track_params(W, H)
    wide = dict(params); wide["th_frame"] = 120.0           # hundreds of un-gated candidates per query against a pool of ~96 per query
    a1 = scenario.build_track_batch(kd[:B + 1], seq_seed, 1, W, H, nrect, 8, *strides, n_map=1500, seed=3)
    a2 = scenario.build_track_batch(kd[1:B + 2], seq_seed, 2, W, H, nrect, 8, *strides, n_map=1500, seed=3)
    d1, p1 = _dev(a1); d2, p2 = _dev(a2)
    f1 = torch.from_numpy(frames[1:B + 1]).cuda(); f2 = torch.from_numpy(frames[2:B + 2]).cuda()
    table, res = scenario.resident_forms(a2)
    mt = pysdyn.MapTable(len(table)); mt.update(0, table)
    torch.cuda.synchronize()
    dr, pr = _dev(res)
    rp = {k: v for k, v in p2.items() if k not in ("last_points", "last_keys", "last_keys_un", "n_last", "map_points")}
    rp.update(pr)
    pysdyn.track_batch_device(gpu, B, f1.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(p1, 0, strides, params))   # primes LastFrame
    pysdyn.track_fetch(gpu, B)
    failures = 0
    for attempt in range(6):
        pysdyn.track_batch_device(gpu, B, f2.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(rp, 0, strides, wide, map_table=mt))
        try:
            got = pysdyn.track_fetch(gpu, B)
            break
        except pysdyn.SdynError as e:
            assert e.code == -3 and "pool" in str(e), str(e)             # SDYN_ERR_CAPACITY
            failures += 1
    else:
        raise AssertionError("the pool never became large enough")
    assert failures >= 1                                                  # the overflow path really ran
    for f in range(B):
        k, d = kd[f + 2]
        ea, el, em, ec = oracle_track.track_frame(k, d, cpu.scale, W, H, a2, f, wide, last_stride)
        assert np.array_equal(got[3][f], ec), (f, got[3][f], ec)
        assert np.array_equal(got[0][f, :len(k)], ea) and np.array_equal(got[1][f, :len(k)], el) and np.array_equal(got[2][f, :len(k)], em)
    mt.close(); gpu.close()


def test_track_device_frustum_matches_explicit():
    """Frame::isInFrustum on the device (map_flags form): local-map ids + one state byte per point; the projection records
    the explicit form uploads are derived from the resident MapPoint table and the frame's pose."""
    import torch
    cfg, B = "kitti", 2
    W, H, nrect, nf, ini, mn, seq_seed, frames, cpu, kd = _sequence(cfg, B + 1)
    gpu = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    last_stride, map_stride, ref_stride = gpu.cap, 2000, 512
    strides = (last_stride, map_stride, ref_stride)
    params = scenario.track_params(W, H)
    arrays = scenario.build_track_batch(kd, seq_seed, 1, W, H, nrect, 8, *strides, n_map=2000, seed=3, frustum=True, scale=cpu.scale)
    assert 0.3 < arrays["map_points"]["track_in_view"].mean() < 0.95
    step = {k: v for k, v in arrays.items() if k not in ("map_table", "map_flags")}
    dev, ptrs = _dev(step)
    dframes = torch.from_numpy(frames[1:]).cuda()
    pysdyn.track_batch_device(gpu, B, dframes.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(ptrs, 0, strides, params))
    want = [x.copy() for x in pysdyn.track_fetch(gpu, B)]
    for f in range(B):                    # the explicit form is the oracle's result
        k, d = kd[f + 1]
        ea, el, em, ec = oracle_track.track_frame(k, d, cpu.scale, W, H, arrays, f, params, last_stride)
        assert np.array_equal(want[3][f], ec) and np.array_equal(want[0][f, :len(k)], ea) and ec[1] > 50
    table, res = scenario.resident_forms(arrays)
    mt = pysdyn.MapTable(len(table)); mt.update(0, table)
    torch.cuda.synchronize()
    dr, pr = _dev({"map_ids": res["map_ids"], "map_flags": res["map_flags"]})
    rp = {k: v for k, v in ptrs.items() if k != "map_points"}
    rp.update(pr)
    pysdyn.track_batch_device(gpu, B, dframes.data_ptr(), W * H, W, H, W, pysdyn.track_inputs(rp, 0, strides, params, map_table=mt))
    got = pysdyn.track_fetch(gpu, B)
    assert np.array_equal(got[3], want[3])
    for f in range(B):
        n = len(kd[f + 1][0])
        for a, b in zip(got[:3], want[:3]):
            assert np.array_equal(a[f, :n], b[f, :n])
    mt.close(); gpu.close()
