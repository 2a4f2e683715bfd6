"""Oracle composition of one frame of the batched front end (TEST INFRASTRUCTURE ONLY): the same four steps
as sdyn_track_batch_device, built from the oracle's restated reference functions."""
import numpy as np

import orc
import pysdyn
import scenario


def track_frame(keys, desc, scale, W, H, arrays, f, params, last_stride, keys_un=None, bounds=None, u_right=None):
    """Returns (assign, locked, dyn_mask, counts[4]) for frame f of `arrays` (scenario.build_track_batch)."""
    cam = scenario.KITTI_CAM
    bounds = (0.0, 0.0, float(W), float(H)) if bounds is None else tuple(float(b) for b in bounds)
    cur = pysdyn.FrameView(keys, desc, scale, bounds, keys_un=keys_un, u_right=u_right,
                           cam=(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"], cam["bf"] / cam["fx"]),
                           tcw=arrays["poses"][f, :12])
    n0 = int(arrays["n_last"][f])
    last = pysdyn.FrameView(arrays["last_keys"][f, :n0], np.zeros((n0, 32), np.uint8), scale,
                            bounds, keys_un=arrays["last_keys_un"][f, :n0],
                            cam=cur.cam, tcw=arrays["poses"][f, 12:])
    n1, assign, locked = orc.match_projection_frame(cur, last, arrays["last_points"][f, :n0], params["th_frame"],
                                                    bool(params["mono"]), bool(params["check_orientation"]))
    nm = int(arrays["n_map"][f])
    n2, assign, locked = orc.match_projection_map(cur, arrays["map_points"][f, :nm], params["th_map"],
                                                  params["nnratio_map"], assign, locked, assign_base=last_stride)
    in_box, dyn = dyn_mask(keys, desc, arrays, f, keys_un)
    return assign, locked, dyn.astype(np.uint8), np.array([n1, n2, int(in_box.sum()), int(dyn.sum())], np.int32)


def rgbd_order(keys, desc, arrays, f, keys_un=None):
    """The list the RGB-D constructor path tracks with: firstSeparate + tail split, Separate per surviving box, UpdateFrame.
    -> (order [extraction indices: static keypoints, then the re-admitted ones], N_s, in_box, readmit)."""
    ku = keys if keys_un is None else keys_un
    nb = int(arrays["n_boxes"][f])
    boxes = arrays["boxes"][f, :nb]
    in_box = orc.box_mask(keys, boxes) != 0
    sep = orc.first_separate(keys, boxes, np.arange(nb))
    ns = len(keys) - sep["n_dyn"]
    slots = {}
    for s, k in sep["dyn"]:
        slots.setdefault(s, []).append(k)
    nslots = len(sep["boxes"])
    dyn_status = [np.zeros(0, np.int32) for _ in range(nslots)]
    static_exit = False
    for s in range(nslots):
        surv = int(sep["box_idx"][s]); r = int(arrays["ref_box"][f, surv]); ks = slots.get(s, [])
        if r < 0 or not ks:
            continue
        o0, o1 = int(arrays["ref_off"][f, r]), int(arrays["ref_off"][f, r + 1])
        if o1 == o0:
            continue
        qd = desc[ks]; qx = np.stack([ku["x"][ks], ku["y"][ks]], 1)
        mq, mt, md, fd = orc.separate_pairs([(qd, qx, arrays["ref_desc"][f, o0:o1], arrays["ref_xy"][f, o0:o1])], arrays["fmat"][f], 0)[0]
        good = len(mq)
        if good < 3 or good < 0.2 * len(ks):
            continue
        dyn_status[s] = fd
        if int((fd != -1).sum()) > max(1.0, 0.2 * good):
            static_exit = True
    readmit = np.zeros(len(keys), bool)
    order = list(sep["order"][:ns])
    if static_exit:                       # if (Separate(...) == 1) mCurrentFrame.UpdateFrame(dynSatus)
        for s, q in orc.update_frame_list(dyn_status, [slots.get(s, []) for s in range(nslots)]):
            order.append(slots[s][q]); readmit[slots[s][q]] = True
    return np.asarray(order, np.int64), ns, in_box, readmit


def track_frame_rgbd(keys, desc, scale, W, H, arrays, f, params, last_stride, last_view=None, keys_un=None, bounds=None):
    """The same frame with RGB-D-constructor semantics (rgbd_split): Frame::firstSeparate + tail split (src/Frame.cc:555-604,
    337-367), Tracking::Separate (Tracking.cc:1093-1239), Frame::UpdateFrame (Frame.cc:607-653), THEN the two searches on the
    frame the reference tracks with (static keypoints + re-admitted ones).
    Returns (order, N_s, assign, locked, dyn_mask, counts): assign / locked index the tracked list, order[j] = extraction index."""
    cam = scenario.KITTI_CAM
    bounds = (0.0, 0.0, float(W), float(H)) if bounds is None else tuple(float(b) for b in bounds)
    ku = keys if keys_un is None else keys_un
    order, ns, in_box, readmit = rgbd_order(keys, desc, arrays, f, keys_un)
    cid = np.where(in_box, np.arange(len(keys)), -1).astype(np.int32)
    fk = keys[order].copy(); fk["class_id"] = cid[order]
    fku = ku[order].copy(); fku["class_id"] = cid[order]
    cur = pysdyn.FrameView(fk, desc[order], scale, bounds, keys_un=fku,
                           cam=(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"], cam["bf"] / cam["fx"]), tcw=arrays["poses"][f, :12])
    if last_view is None:
        n0 = int(arrays["n_last"][f])
        last_view = pysdyn.FrameView(arrays["last_keys"][f, :n0], np.zeros((n0, 32), np.uint8), scale, bounds,
                                     keys_un=arrays["last_keys_un"][f, :n0], cam=cur.cam, tcw=arrays["poses"][f, 12:])
    n1, assign, locked = orc.match_projection_frame(cur, last_view, arrays["last_points"][f, :last_view.n], params["th_frame"],
                                                    bool(params["mono"]), bool(params["check_orientation"]))
    nm = int(arrays["n_map"][f])
    n2, assign, locked = orc.match_projection_map(cur, arrays["map_points"][f, :nm], params["th_map"], params["nnratio_map"], assign, locked,
                                                  assign_base=last_stride)
    dyn = in_box & ~readmit
    return order, ns, assign, locked, dyn.astype(np.uint8), np.array([n1, n2, int(in_box.sum()), int(dyn.sum())], np.int32), cur


def dyn_mask(keys, desc, arrays, f, keys_un=None):
    """(in_box, masked) per extracted keypoint: firstSeparate's box test, Separate's per-box BFMatcher + classifyF with its gates,
    UpdateFrame's re-admission (stereo-constructor bookkeeping: the frame keeps every keypoint)."""
    # dynamic mask: firstSeparate (+ erase bug) -> Separate (BF + classifyF + gates) -> UpdateFrame
    nb = int(arrays["n_boxes"][f])
    boxes = arrays["boxes"][f, :nb]
    in_box = orc.box_mask(keys, boxes) != 0
    sep = orc.first_separate(keys, boxes, np.arange(nb))
    slots = {}
    for s, k in sep["dyn"]:
        slots.setdefault(s, []).append(k)
    readmit = np.zeros(len(keys), bool)
    static_exit = False
    for s in range(len(sep["boxes"])):
        surv = int(sep["box_idx"][s])                  # original id of the box now in slot s
        r = int(arrays["ref_box"][f, surv])
        ks = slots.get(s, [])
        if r < 0 or not ks:
            continue
        o0, o1 = int(arrays["ref_off"][f, r]), int(arrays["ref_off"][f, r + 1])
        if o1 == o0:
            continue
        ku = keys if keys_un is None else keys_un      # classifyF reads mvdynKeysUn (Tracking.cc:1129-1131)
        qd = desc[ks]; qx = np.stack([ku["x"][ks], ku["y"][ks]], 1)
        mq, mt, md, fd = orc.separate_pairs([(qd, qx, arrays["ref_desc"][f, o0:o1], arrays["ref_xy"][f, o0:o1])],
                                            arrays["fmat"][f], 0)[0]
        good = len(mq)
        if good < 3 or good < 0.2 * len(ks):
            continue
        num0 = int((fd != -1).sum())
        for q in fd[fd != -1]:
            readmit[ks[int(q)]] = True
        if num0 > max(1.0, 0.2 * good):
            static_exit = True
    dyn = in_box & ~(readmit & static_exit)
    return in_box, dyn
