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
    d = desc.copy()
    for i in range(len(d)):
        for _ in range(int(nbits[i])):
            k = int(r.integers(0, 256))
            d[i, k >> 3] ^= np.uint8(1 << (k & 7))
    return d


def last_points(last_keys, last_desc, shift, cam=KITTI_CAM, seed=0, p_mp=0.85, p_outlier=0.05, p_obs=0.9, noise_bits=6):
    """LastFrame map points: back-project every last-frame keypoint at a seeded depth so that, with the
    current pose = identity, it projects onto its position in the current frame (content shifted by `shift`)."""
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


def bow_nodes(desc, bits=6):
    """Stand-in for DBoW2's node assignment (the vocabulary file is absent): a node id from descriptor bits,
    so that similar descriptors share a node often, as in a real vocabulary."""
    return (desc[:, 0] & ((1 << bits) - 1)).astype(np.uint32) * 7 + 3


def degenerate_descriptors(n, seed, distinct=12):
    """Tie-heavy descriptors: only `distinct` different values, so distance ties and claim conflicts abound."""
    r = rng_for(seed + 303)
    base = r.integers(0, 256, (distinct, 32), dtype=np.uint8)
    return base[r.integers(0, distinct, n)]
