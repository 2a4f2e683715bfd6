"""Independent Python twin of Frame::ComputeStereoMatches (reference: src/Frame.cc:874-1048) on REAL OpenCV
primitives: the window arithmetic runs through cv2 (convertTo / subtract / cv2.norm(NORM_L1)) exactly as the
reference's cv::Mat expressions do, and all scalar arithmetic is numpy float32.  Used by
tests/test_oracle_stereo.py to pin oracle/orc_stereo.cpp.  TEST INFRASTRUCTURE ONLY."""
import math

import cv2
import numpy as np

F = np.float32
TH_HIGH, TH_LOW = 100, 50


def _round(x):
    """C round(): half away from zero, on a float32 value."""
    x = float(x)
    return F(math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5))


def compute_stereo_matches(keys_l, desc_l, keys_r, desc_r, pyr_l, pyr_r, scale, inv_scale, mb, mbf):
    """pyr_l / pyr_r: lists of un-bordered level images (numpy uint8).  Returns (mvuRight, mvDepth)."""
    n = len(keys_l)
    u_right = np.full(n, -1.0, F); depth = np.full(n, -1.0, F)
    th_orb = (TH_HIGH + TH_LOW) // 2
    n_rows = pyr_l[0].shape[0]
    rows = [[] for _ in range(n_rows)]
    for ir, kp in enumerate(keys_r):
        r = F(2.0) * scale[kp["octave"]]
        maxr = int(math.ceil(F(kp["y"] + r))); minr = int(math.floor(F(kp["y"] - r)))
        for yi in range(minr, maxr + 1):
            rows[yi].append(ir)
    min_z = F(mb); min_d = F(0); max_d = F(mbf) / min_z
    bits = np.unpackbits(desc_r, axis=1) if len(desc_r) else np.zeros((0, 256), np.uint8)
    dist_idx = []
    for il, kp in enumerate(keys_l):
        level = int(kp["octave"]); v_l = F(kp["y"]); u_l = F(kp["x"])
        cands = rows[int(v_l)]
        if not cands:
            continue
        min_u = F(u_l - max_d); max_u = F(u_l - min_d)
        if max_u < 0:
            continue
        best, best_r = TH_HIGH, 0
        dl = np.unpackbits(desc_l[il])
        for ir in cands:
            kr = keys_r[ir]
            if kr["octave"] < level - 1 or kr["octave"] > level + 1:
                continue
            if kr["x"] >= min_u and kr["x"] <= max_u:
                d = int(np.count_nonzero(dl != bits[ir]))
                if d < best:
                    best, best_r = d, ir
        if best >= th_orb:
            continue
        u_r0 = F(keys_r[best_r]["x"])
        sf = inv_scale[level]
        su_l = _round(F(u_l * sf)); sv_l = _round(F(v_l * sf)); su_r0 = _round(F(u_r0 * sf))
        w, L = 5, 5
        il_img = pyr_l[level][int(sv_l - w):int(sv_l + w + 1), int(su_l - w):int(su_l + w + 1)].astype(np.float32)
        il_img = cv2.subtract(il_img, il_img[w, w] * np.ones(il_img.shape, np.float32))
        ini_u = F(su_r0 + L - w); end_u = F(su_r0 + L + w + 1)
        if ini_u < 0 or end_u >= pyr_r[level].shape[1]:
            continue
        best_d, best_inc = 2 ** 31 - 1, 0
        dists = np.zeros(2 * L + 1, F)
        for inc in range(-L, L + 1):
            ir_img = pyr_r[level][int(sv_l - w):int(sv_l + w + 1),
                                  int(su_r0 + inc - w):int(su_r0 + inc + w + 1)].astype(np.float32)
            ir_img = cv2.subtract(ir_img, ir_img[w, w] * np.ones(ir_img.shape, np.float32))
            dist = F(cv2.norm(il_img, ir_img, cv2.NORM_L1))
            if dist < F(best_d):
                best_d, best_inc = int(dist), inc
            dists[L + inc] = dist
        if best_inc == -L or best_inc == L:
            continue
        d1, d2, d3 = dists[L + best_inc - 1], dists[L + best_inc], dists[L + best_inc + 1]
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = F(F(d1 - d3) / F(F(2.0) * F(F(d1 + d3) - F(F(2.0) * d2))))
        if delta < -1 or delta > 1:
            continue
        best_ur = F(scale[level] * F(F(su_r0 + F(best_inc)) + delta))
        disparity = F(u_l - best_ur)
        if disparity >= min_d and disparity < max_d:
            if disparity <= 0:
                disparity = F(0.01); best_ur = F(float(u_l) - 0.01)
            depth[il] = F(mbf) / disparity
            u_right[il] = best_ur
            dist_idx.append((best_d, il))
    if not dist_idx:
        return u_right, depth
    dist_idx.sort()
    median = F(dist_idx[len(dist_idx) // 2][0])
    th = F(F(F(1.5) * F(1.4)) * median)
    for d, il in reversed(dist_idx):
        if F(d) < th:
            break
        u_right[il] = -1; depth[il] = -1
    return u_right, depth
