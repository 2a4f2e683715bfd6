"""Shared helpers for the parity tests."""
import hashlib
import numpy as np
import pysdyn

CONFIGS = {
    # name: (W, H, nrect, nfeatures, iniTh, minTh)  — BASELINE.json configs / SURVEY §8
    "kitti": (1241, 376, 160, 2000, 12, 7),      # Examples/Stereo/KITTI04-12.yaml
    "kitti_mono": (1241, 376, 160, 2000, 20, 7), # Examples/Monocular/KITTI04-12.yaml
    "tum": (640, 480, 120, 1000, 20, 7),         # Examples/RGB-D/TUM3.yaml
    "4k": (3840, 2160, 2800, 8000, 20, 7),       # stress config
    "small": (320, 240, 40, 500, 20, 7),
}
CONFIG_ID = {"kitti": 0, "tum": 1, "kitti_mono": 2, "4k": 4, "small": 5}


def frame(cfg, idx, seq=0, ox=0, oy=0, t=0):
    w, h, r = CONFIGS[cfg][:3]
    seed = 1000 * CONFIG_ID[cfg] + 100000 * seq
    return pysdyn.synth_frame(seed + 7, seed + idx, w, h, r, ox, oy, t)


def kp_tuple_array(k):
    return np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"],
                     k["octave"].astype(np.float32), k["class_id"].astype(np.float32)], 1)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
