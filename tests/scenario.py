"""Seeded synthetic tracking scenarios (SURVEY §8d "synthetic motion / map") for the matcher and dynamic-mask
parity tests and for bench.py.  Everything derives from oracle- or GPU-extracted keypoints of consecutive
synthetic frames, so real correspondences exist."""
import numpy as np

import pysdyn

# KITTI 04-12 stereo intrinsics (Examples/Stereo/KITTI04-12.yaml:8-11,25)
KITTI_CAM = dict(fx=707.0912, fy=707.0912, cx=601.8873, cy=183.1104, bf=379.8145)


def rng_for(seed):
    return np.random.default_rng(seed)


def frame_view(keys, desc, scale, W, H, cam=KITTI_CAM, stereo=False, tcw=None, seed=0):
    """Frame of an undistorted camera: mvKeysUn == mvKeys, bounds = image (Frame::ComputeImageBounds, k1 == 0)."""
    u_right = None
    if stereo:
        r = rng_for(seed + 17)
        z = r.uniform(4.0, 40.0, len(keys)).astype(np.float32)
        u_right = (keys["x"] - np.float32(cam["bf"]) / z).astype(np.float32)
        u_right[r.random(len(keys)) < 0.3] = -1.0           # no stereo match for 30 % of the keypoints
    b = cam["bf"] / cam["fx"]
    return pysdyn.FrameView(keys, desc, scale, (0.0, 0.0, float(W), float(H)), u_right=u_right,
                            cam=(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"], b), tcw=tcw)


def flip_bits(desc, r, nbits):
    """Flip nbits[i] random bits of descriptor i (vectorised)."""
    d = desc.copy()
    nbits = np.asarray(nbits)
    rows = np.arange(len(d))
    for j in range(int(nbits.max()) if len(nbits) else 0):
        sel = rows[nbits > j]
        pos = r.integers(0, 256, len(sel))
        np.bitwise_xor.at(d, (sel, pos >> 3), (1 << (pos & 7)).astype(np.uint8))
    return d


def last_points(last_keys, last_desc, shift, cam=KITTI_CAM, seed=0, p_mp=0.85, p_outlier=0.05, p_obs=0.9, noise_bits=6, tcw=None):
    """LastFrame map points: back-project every last-frame keypoint at a seeded depth so that, with the
    current pose `tcw` (identity rotation + translation; default identity), it projects onto its position in the current
    frame (content shifted by `shift`)."""
    r = rng_for(seed + 101)
    n = len(last_keys)
    lp = np.zeros(n, pysdyn.LASTPOINT_DTYPE)
    lp["has_mp"] = r.random(n) < p_mp
    lp["outlier"] = r.random(n) < p_outlier
    lp["obs_positive"] = r.random(n) < p_obs
    z = r.uniform(4.0, 40.0, n).astype(np.float32)
    u = last_keys["x"] - np.float32(shift[0]) + r.normal(0, 1.0, n).astype(np.float32)
    v = last_keys["y"] - np.float32(shift[1]) + r.normal(0, 1.0, n).astype(np.float32)
    lp["world"][:, 0] = (u - np.float32(cam["cx"])) * z / np.float32(cam["fx"])
    lp["world"][:, 1] = (v - np.float32(cam["cy"])) * z / np.float32(cam["fy"])
    lp["world"][:, 2] = z
    behind = r.random(n) < 0.02
    lp["world"][behind, 2] *= -1                              # a few points behind the camera (invzc < 0)
    if tcw is not None:                                       # x3Dc = x3Dw + t  ->  x3Dw = x3Dc - t
        lp["world"] -= np.asarray(tcw, np.float32).reshape(3, 4)[:, 3]
    lp["desc"] = flip_bits(last_desc, r, r.integers(0, noise_bits + 1, n))
    return lp


def map_queries(keys, desc, nlevels, seed=0, count=None, jitter=2.0):
    """Local-map points as SearchByProjection(F, vpMapPoints) sees them: projections near existing keypoints
    of the frame (so candidates exist), predicted level around the keypoint's octave."""
    r = rng_for(seed + 202)
    n = len(keys) if count is None else count
    src = r.integers(0, len(keys), n)
    mp = np.zeros(n, pysdyn.MAPPOINT_DTYPE)
    mp["proj_x"] = keys["x"][src] + r.normal(0, jitter, n).astype(np.float32)
    mp["proj_y"] = keys["y"][src] + r.normal(0, jitter, n).astype(np.float32)
    z = r.uniform(4.0, 40.0, n).astype(np.float32)
    mp["proj_xr"] = mp["proj_x"] - np.float32(KITTI_CAM["bf"]) / z
    mp["view_cos"] = np.where(r.random(n) < 0.5, 0.9995, 0.99).astype(np.float32)
    mp["level"] = np.clip(keys["octave"][src] + r.integers(0, 2, n), 0, nlevels - 1)
    mp["track_in_view"] = r.random(n) < 0.9
    mp["bad"] = r.random(n) < 0.03
    mp["obs_positive"] = r.random(n) < 0.9
    mp["desc"] = flip_bits(desc[src], r, r.integers(0, 9, n))
    return mp


def map_points_3d(keys, desc, scale, tcw, cam=KITTI_CAM, seed=0, count=None, jitter=2.0):
    """Local-map MapPoints as the reference holds them (world position, mean viewing direction, scale-invariance distances,
    descriptor) placed so that, under pose `tcw` (rows 0..2 of mTcw), they project next to keypoints of the frame — plus the
    per-frame state bits.  Frame::isInFrustum then derives the projection records.  -> (MAP_POINT_DTYPE[n], flags uint8[n])."""
    r = rng_for(seed + 606)
    n = len(keys) if count is None else count
    src = r.integers(0, len(keys), n)
    T = np.asarray(tcw, np.float64).reshape(3, 4)
    R, t = T[:, :3], T[:, 3]
    z = r.uniform(4.0, 40.0, n)
    u = keys["x"][src].astype(np.float64) + r.normal(0, jitter, n); v = keys["y"][src].astype(np.float64) + r.normal(0, jitter, n)
    pc = np.stack([(u - cam["cx"]) * z / cam["fx"], (v - cam["cy"]) * z / cam["fy"], z], 1)
    pc[r.random(n) < 0.03, 2] *= -1                              # behind the camera
    pw = (pc - t) @ R                                           # Rcw^T (pc - tcw)
    pts = np.zeros(n, pysdyn.MAP_POINT_DTYPE)
    pts["world"] = pw.astype(np.float32)
    ow = -(R.T @ t)
    po = pts["world"].astype(np.float64) - ow
    dist = np.linalg.norm(po, axis=1)
    nrm = po / dist[:, None] + r.normal(0, 0.45, (n, 3))        # spread: some fail the 60 degree test
    pts["normal"] = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    lvl = np.clip(keys["octave"][src] + r.integers(0, 2, n), 0, len(scale) - 1)
    ref_dist = dist * r.uniform(0.6, 1.5, n)                    # distance at which the point was created: some fall outside [0.8 min, 1.2 max]
    pts["max_distance"] = (ref_dist * scale[lvl]).astype(np.float32)
    pts["min_distance"] = (pts["max_distance"] / scale[-1]).astype(np.float32)
    pts["desc"] = flip_bits(desc[src], r, r.integers(0, 9, n))
    flags = ((r.random(n) < 0.03) * pysdyn.MP_BAD + (r.random(n) < 0.9) * pysdyn.MP_OBS_POSITIVE + (r.random(n) < 0.05) * pysdyn.MP_SKIP).astype(np.uint8)
    return pts, flags


def map_queries_from_frustum(view, pts, flags, log_sf, cos_limit=0.5):
    """SearchByProjection(F, vpMapPoints)'s view of those MapPoints: Frame::isInFrustum (oracle) fills the tracking fields."""
    import orc
    fr = orc.in_frustum(view, log_sf, pts["world"], pts["normal"], pts["min_distance"], pts["max_distance"], cos_limit)
    mp = np.zeros(len(pts), pysdyn.MAPPOINT_DTYPE)
    for name in ("proj_x", "proj_y", "proj_xr", "view_cos", "level"):
        mp[name] = fr[name]
    mp["track_in_view"] = (fr["in_view"] != 0) & ((flags & pysdyn.MP_SKIP) == 0)
    mp["bad"] = (flags & pysdyn.MP_BAD) != 0
    mp["obs_positive"] = (flags & pysdyn.MP_OBS_POSITIVE) != 0
    mp["desc"] = pts["desc"]
    return mp


def bow_nodes(desc, bits=6):
    """Stand-in for DBoW2's node assignment (the vocabulary file is absent): a node id from descriptor bits,
    so that similar descriptors share a node often, as in a real vocabulary."""
    return (desc[:, 0] & ((1 << bits) - 1)).astype(np.uint32) * 7 + 3


def degenerate_descriptors(n, seed, distinct=12):
    """Tie-heavy descriptors: only `distinct` different values, so distance ties and claim conflicts abound."""
    r = rng_for(seed + 303)
    base = r.integers(0, 256, (distinct, 32), dtype=np.uint8)
    return base[r.integers(0, distinct, n)]


# ---------------------------------------------------------------------------------------------------
# batched front-end inputs (sdyn_track_inputs): a sequence of frames, each tracked against its predecessor
# ---------------------------------------------------------------------------------------------------
def _tri(i, period):
    """Triangle wave 0..period..0: keeps the camera (and the moving objects) inside the populated part of the
    synthetic world however long the sequence is, while consecutive frames still differ by a small shift."""
    m = i % (2 * period)
    return m if m <= period else 2 * period - m


def sequence_offsets(i):
    """Camera offset of frame i of a synthetic sequence (integer shifts <= 8 px between frames)."""
    return 3 * _tri(i, 16), (i * 5) % 7


def sequence_time(i):
    """Time step of the independently moving rectangles at frame i (they oscillate, so they stay in view)."""
    return _tri(i, 12)


def translation_fmat(dx, dy):
    """F21 with x2^T F x1 = 0 for x2 = x1 - (dx, dy): pure image translation (static scene)."""
    return np.float32([[0, 0, dy], [0, 0, -dx], [-dy, dx, 0]])


def frame_pose(i):
    """Pose pair of frame i of a synthetic sequence: the camera advances, retreats or stands still along z in turn, so the
    forward / backward / neutral level rules of SearchByProjection(cur, last) (ORBmatcher.cc:1505-1506) all occur in a batch
    and every frame of a batch has its own pose."""
    cur = np.eye(4, dtype=np.float32)[:3].copy(); last = np.eye(4, dtype=np.float32)[:3].copy()
    cur[:, 3] = [0.01 * (i % 5), -0.02 * (i % 3), (0.0, 1.5, -1.5, 0.25)[i % 4]]
    return np.concatenate([cur.reshape(12), last.reshape(12)])


def build_track_batch(kd, seq_seed, first_index, W, H, nrect, nlevels, last_stride, map_stride, ref_stride,
                      n_map=3000, seed=0, offsets=None, time=None, frustum=False, scale=None, cam=KITTI_CAM):
    """kd: list of (keys, desc) for frames first_index-1 .. first_index+B-1 of one sequence (B = len(kd)-1).
    Returns dict of numpy arrays laid out as sdyn_track_inputs expects ([B, stride, ...])."""
    B = len(kd) - 1
    out = {
        "last_points": np.zeros((B, last_stride), pysdyn.LASTPOINT_DTYPE),
        "last_keys": np.zeros((B, last_stride), pysdyn.KP_DTYPE),
        "last_keys_un": np.zeros((B, last_stride), pysdyn.KP_DTYPE),
        "n_last": np.zeros(B, np.int32),
        "map_points": np.zeros((B, map_stride), pysdyn.MAPPOINT_DTYPE),
        "n_map": np.zeros(B, np.int32),
        "boxes": np.zeros((B, 64, 4), np.float64),
        "n_boxes": np.zeros(B, np.int32),
        "ref_box": np.full((B, 64), -1, np.int32),
        "ref_desc": np.zeros((B, ref_stride, 32), np.uint8),
        "ref_xy": np.zeros((B, ref_stride, 2), np.float32),
        "ref_off": np.zeros((B, 65), np.int32),
        "fmat": np.zeros((B, 9), np.float32),
        "poses": np.zeros((B, 24), np.float32),
    }
    offsets = offsets or sequence_offsets
    time = time or sequence_time
    for f in range(B):
        i = first_index + f
        (k0, d0), (k1, d1) = kd[f], kd[f + 1]
        out["poses"][f] = frame_pose(i)
        ox0, oy0 = offsets(i - 1); ox1, oy1 = offsets(i)
        shift = (ox1 - ox0, oy1 - oy0)
        n0 = min(len(k0), last_stride)
        out["last_points"][f, :n0] = last_points(k0[:n0], d0[:n0], shift, seed=seed + 31 * i, tcw=out["poses"][f, :12])
        out["last_keys"][f, :n0] = k0[:n0]; out["last_keys_un"][f, :n0] = k0[:n0]
        out["n_last"][f] = n0
        nm = min(n_map, map_stride)
        if frustum:
            # 3-D MapPoints + Frame::isInFrustum: the projection records are derived, not drawn
            pts, fl = map_points_3d(k1, d1, scale, out["poses"][f, :12], cam=cam, seed=seed + 57 * i, count=nm)
            view = pysdyn.FrameView(k1, d1, scale, (0.0, 0.0, float(W), float(H)),
                                    cam=(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["bf"], cam["bf"] / cam["fx"]), tcw=out["poses"][f, :12])
            out["map_points"][f, :nm] = map_queries_from_frustum(view, pts, fl, np.log(np.float32(1.2)))
            out.setdefault("map_table", np.zeros((B, map_stride), pysdyn.MAP_POINT_DTYPE))[f, :nm] = pts
            out.setdefault("map_flags", np.zeros((B, map_stride), np.uint8))[f, :nm] = fl
        else:
            out["map_points"][f, :nm] = map_queries(k1, d1, nlevels, seed=seed + 57 * i, count=nm)
        out["n_map"][f] = nm
        # detection boxes of the current frame and of the reference (= previous) frame, joined on rectangle id
        b1, id1 = pysdyn.synth_boxes_ids(seq_seed, W, H, nrect, ox1, oy1, time(i), margin=8)
        b0, id0 = pysdyn.synth_boxes_ids(seq_seed, W, H, nrect, ox0, oy0, time(i - 1), margin=8)
        nb = min(len(b1), 62)
        out["boxes"][f, :nb] = b1[:nb]
        out["boxes"][f, nb] = [W + 50.0, H + 50.0, 10.0, 10.0]      # a detection without keypoints (gets erased)
        out["n_boxes"][f] = nb + 1
        for b in range(nb):
            hit = np.nonzero(id0 == id1[b])[0]
            out["ref_box"][f, b] = int(hit[0]) if len(hit) else -1
        off = 0
        for r in range(min(len(b0), 64)):
            x, y, w, h = b0[r]
            inside = np.nonzero((k0["x"] >= x) & (k0["x"] < x + w) & (k0["y"] >= y) & (k0["y"] < y + h))[0]
            inside = inside[: max(0, ref_stride - off)]
            out["ref_off"][f, r] = off
            out["ref_desc"][f, off:off + len(inside)] = d0[inside]
            out["ref_xy"][f, off:off + len(inside), 0] = k0["x"][inside]
            out["ref_xy"][f, off:off + len(inside), 1] = k0["y"][inside]
            off += len(inside)
        out["ref_off"][f, min(len(b0), 64):] = off
        out["fmat"][f] = translation_fmat(*shift).reshape(9)
    return out


def track_params(W, H, cam=KITTI_CAM, th_frame=7.0, th_map=3.0, nnratio_map=0.8, mono=0, check_orientation=1):
    return dict(min_x=0.0, min_y=0.0, max_x=float(W), max_y=float(H), fx=cam["fx"], fy=cam["fy"], cx=cam["cx"],
                cy=cam["cy"], bf=cam["bf"], b=cam["bf"] / cam["fx"], th_frame=th_frame,
                th_map=th_map, nnratio_map=nnratio_map, mono=mono, check_orientation=check_orientation)


def reorder_last(arrays, f, order):
    """LastFrame of frame input f becomes the list `order` of its keypoints (what an rgbd_split step leaves behind: the static
    keypoints followed by the re-admitted ones, src/Frame.cc:337-367,607-653) — per-keypoint rows move with their keypoint."""
    order = np.asarray(order, np.int64)
    n = len(order)
    for name in ("last_keys", "last_keys_un", "last_points"):
        if name in arrays:
            rows = arrays[name][f, order].copy()
            arrays[name][f] = np.zeros((), arrays[name].dtype)
            arrays[name][f, :n] = rows
    arrays["n_last"][f] = n


def resident_forms(arrays):
    """The same step inputs in the resident forms of sdyn_track_inputs: one MapPoint table (every LastFrame point and every
    local-map point of every frame gets an id) plus, per frame, ids / flag bytes / projection records.
    -> (table [MAP_POINT_DTYPE], {last_ids, last_flags, map_ids, map_proj})."""
    lp, mp = arrays["last_points"], arrays["map_points"]
    B, ls = lp.shape; ms = mp.shape[1]
    table = np.zeros(B * (ls + ms), pysdyn.MAP_POINT_DTYPE)
    out = {"last_ids": np.full((B, ls), -1, np.int32), "last_flags": np.zeros((B, ls), np.uint8),
           "map_ids": np.full((B, ms), -1, np.int32), "map_proj": np.zeros((B, ms), pysdyn.MAP_PROJ_DTYPE)}
    if "map_flags" in arrays:              # frustum scenarios: the table carries the full MapPoint, the device derives map_proj
        out["map_flags"] = arrays["map_flags"].copy()
    for f in range(B):
        base = f * (ls + ms)
        table["world"][base:base + ls] = lp["world"][f]; table["desc"][base:base + ls] = lp["desc"][f]
        out["last_ids"][f] = np.where(lp["has_mp"][f] != 0, base + np.arange(ls), -1)
        out["last_flags"][f] = (lp["outlier"][f] != 0) * pysdyn.LP_OUTLIER + (lp["obs_positive"][f] != 0) * pysdyn.LP_OBS_POSITIVE
        if "map_table" in arrays:
            table[base + ls:base + ls + ms] = arrays["map_table"][f]
        table["desc"][base + ls:base + ls + ms] = mp["desc"][f]
        out["map_ids"][f] = base + ls + np.arange(ms)
        for name in ("proj_x", "proj_y", "proj_xr", "view_cos", "level", "track_in_view", "bad", "obs_positive"):
            out["map_proj"][name][f] = mp[name][f]
    return table, out


def stereo_pair(cfg, idx, disparities=(5, 11, 23), seq=0):
    """Rectified left / right frames: the right view shows the left content shifted by a per-band disparity (the
    image is cut into len(disparities) horizontal bands), plus its own sensor noise."""
    import common
    left = common.frame(cfg, idx, seq=seq)
    right = np.empty_like(left)
    h = left.shape[0]
    edges = np.linspace(0, h, len(disparities) + 1).astype(int)
    for d, y0, y1 in zip(disparities, edges[:-1], edges[1:]):
        right[y0:y1] = common.frame(cfg, idx + 1000, seq=seq, ox=d)[y0:y1]
    return left, right


def pose_small(seed=0, angle_deg=1.5, t=(0.05, -0.02, 0.1)):
    """A small camera motion as the float32 (Rcw, tcw, Ow = -Rcw^T tcw) triple the pose searches take."""
    r = rng_for(seed + 303)
    ax = r.normal(size=3); ax /= np.linalg.norm(ax)
    a = np.deg2rad(angle_deg)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = (np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K).astype(np.float32)
    tcw = np.asarray(t, np.float32)
    ow = (-(R.T.astype(np.float32) @ tcw)).astype(np.float32)
    return R, tcw, ow


def proj_points(keys, desc, scale, R, tcw, ow, cam=KITTI_CAM, seed=0, p_valid=0.85, noise_bits=6, jitter=1.0):
    """Candidate MapPoints of the pose-projection searches: every keypoint of the searched frame is back-projected at
    a seeded depth through the given pose, so the point projects next to it; distances / normals follow
    MapPoint::UpdateNormalAndDepth (mfMaxDistance = dist * scale[octave], mfMinDistance = mfMaxDistance / scale[last])."""
    r = rng_for(seed + 404)
    n = len(keys)
    pp = np.zeros(n, pysdyn.PROJPOINT_DTYPE)
    pp["valid"] = r.random(n) < p_valid
    z = r.uniform(4.0, 40.0, n)
    u = keys["x"].astype(np.float64) + r.normal(0, jitter, n); v = keys["y"].astype(np.float64) + r.normal(0, jitter, n)
    pc = np.stack([(u - cam["cx"]) * z / cam["fx"], (v - cam["cy"]) * z / cam["fy"], z], 1)
    behind = r.random(n) < 0.03
    pc[behind, 2] *= -1                                        # some points behind the camera
    pw = (pc - tcw.astype(np.float64)) @ R.astype(np.float64)  # Rcw^T (pc - tcw)
    pp["world"] = pw.astype(np.float32)
    po = pp["world"].astype(np.float64) - ow.astype(np.float64)
    dist = np.linalg.norm(po, axis=1)
    nrm = po / dist[:, None] + r.normal(0, 0.5, (n, 3))        # viewing direction with a spread: some fail the 60 degree test
    pp["normal"] = (nrm / np.linalg.norm(nrm, axis=1)[:, None]).astype(np.float32)
    lvl = np.clip(keys["octave"] + r.integers(-1, 2, n), 0, len(scale) - 1)
    ref_dist = dist * r.uniform(0.7, 1.4, n)                   # distance at which the point was created
    raw = (ref_dist * scale[lvl]).astype(np.float32)
    pp["max_distance_raw"] = raw
    pp["max_distance"] = np.float32(1.2) * raw
    pp["min_distance"] = np.float32(0.8) * (raw / scale[-1])
    pp["angle"] = np.mod(keys["angle"] + r.normal(0, 8.0, n), 360).astype(np.float32)
    pp["desc"] = flip_bits(desc, r, r.integers(0, noise_bits + 1, n))
    return pp


def synthetic_vocabulary(k=10, L=3, seed=0, ragged=0.0, stop=0.02, base_desc=None):
    """A DBoW2-shaped vocabulary tree in node-id (file) order: (parent, is_leaf, desc, weight).  Children descriptors
    are the parent's with a few flipped bits (so descents are decisive but ties occur), leaves get IDF-like weights,
    `stop` of them weight 0 (stopped words) and `ragged` of the inner nodes end early (leaves above level L)."""
    r = rng_for(seed + 505)
    parent, leaf, desc, weight, level = [0], [0], [np.zeros(32, np.uint8) if base_desc is None else base_desc], [0.0], [0]
    frontier = [0]
    for lvl in range(1, L + 1):
        nxt = []
        for p in frontier:
            for _ in range(k):
                d = desc[p].copy()
                for b in r.integers(0, 256, max(1, 48 >> lvl)):
                    d[b >> 3] ^= np.uint8(1 << (b & 7))
                nid = len(parent)
                is_leaf = lvl == L or (lvl >= 2 and r.random() < ragged)
                parent.append(p); leaf.append(int(is_leaf)); desc.append(d); level.append(lvl)
                weight.append(0.0 if (is_leaf and r.random() < stop) else (float(r.uniform(0.5, 9.0)) if is_leaf else 0.0))
                if not is_leaf:
                    nxt.append(nid)
        frontier = nxt
    # DBoW2 writes nodes level by level, which is the order produced here (parents precede children)
    return (np.array(parent, np.int32), np.array(leaf, np.uint8), np.stack(desc).astype(np.uint8), np.array(weight, np.float64))


def write_vocabulary_text(path, parent, leaf, desc, weight, k, L):
    """ORBvoc.txt layout (TemplatedVocabulary::saveToTextFile): header 'k L scoring weighting', then per node
    'parent isLeaf d0 .. d31 weight'."""
    with open(path, "w") as f:
        f.write("%d %d 0 0\n" % (k, L))
        for i in range(1, len(parent)):
            f.write("%d %d %s %.17g\n" % (parent[i], leaf[i], " ".join(str(int(b)) for b in desc[i]), weight[i]))
