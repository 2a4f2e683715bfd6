#!/usr/bin/env python3
"""Generates tests/golden/extract_cv2.json: known-answer vectors for the extraction path produced with the REAL
OpenCV primitives (cv2 4.13.0) through tests/cv2_twin.py on seeded synthetic frames.  The reference ships no
golden vectors, and its own binary cannot be built here, so OpenCV-backed outputs are the pin:
the oracle (CPU tests) and the CUDA path (GPU tests) must both reproduce these hashes.

Also writes tests/golden/match_oracle.json: regression hashes of the oracle's matcher / dynamic-mask outputs
on seeded scenarios (these have no independent OpenCV counterpart; they guard against drift).

Also writes tests/golden/next_rows.json (tests/next_cases.py): digests of the SURVEY 8(f) rows from the independent
twins (cv2-backed) where they exist and from the oracle otherwise.

Run in the build container:  python tools/gen_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "slam-dynamic_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
import cv2_twin  # noqa: E402
import orc  # noqa: E402
import pysdyn  # noqa: E402
import scenario  # noqa: E402


def kp_array(tk):
    a = np.zeros(len(tk), pysdyn.KP_DTYPE)
    for i, (x, y, size, ang, resp, octv, cid) in enumerate(tk):
        a[i] = (x, y, size, ang, resp, octv, cid)
    return a


def main():
    out = {"source": "cv2 %s via tests/cv2_twin.py" % __import__("cv2").__version__, "frames": []}
    for cfg, idxs in [("small", [0, 1, 2]), ("tum", [0, 1]), ("kitti", [0, 1]), ("kitti_mono", [0])]:
        W, H, _, nf, ini, mn = common.CONFIGS[cfg]
        T = cv2_twin.Twin(nf, 1.2, 8, ini, mn)
        for idx in idxs:
            img = common.frame(cfg, idx)
            tk, td = T(img)
            k = kp_array(tk)
            out["frames"].append({
                "config": cfg, "index": idx, "image_sha256": common.sha(img), "n": len(k),
                "pyramid_sha256": [common.sha(T.pyr[l]) for l in range(8)],
                "keypoints_sha256": common.sha(k), "descriptors_sha256": common.sha(td),
                "per_level": np.bincount(k["octave"], minlength=8).tolist(),
            })
            print(cfg, idx, len(k))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "extract_cv2.json"), "w"), indent=1)

    m = {"source": "oracle regression vectors", "cases": []}
    for cfg in ("tum", "kitti"):
        W, H, _, nf, ini, mn = common.CONFIGS[cfg]
        E = orc.Extractor(nf, 1.2, 8, ini, mn)
        k0, d0 = E(common.frame(cfg, 0)); k1, d1 = E(common.frame(cfg, 1, ox=4, oy=1, t=1))
        cur = scenario.frame_view(k1, d1, E.scale, W, H, stereo=True, seed=1)
        last = scenario.frame_view(k0, d0, E.scale, W, H, stereo=True, seed=0)
        lp = scenario.last_points(k0, d0, (4, 1), seed=7)
        n1, a1, l1, pr = orc.match_projection_frame(cur, last, lp, 7.0, False, True, want_pairs=True)
        mp = scenario.map_queries(k1, d1, 8, seed=3, count=3000)
        n2, a2, l2 = orc.match_projection_map(cur, mp, 3.0, 0.8, a1, l1)
        n3, m12, prev = orc.match_init(last, cur, np.stack([k0["x"], k0["y"]], 1), 100, 0.9, True)
        fa = pysdyn.FeatureVector(scenario.bow_nodes(d0)); fb = pysdyn.FeatureVector(scenario.bow_nodes(d1))
        n4, ab = orc.match_bow(last, np.ones(len(k0), np.uint8), fa, cur, fb, 0.7, True)
        m["cases"].append({"config": cfg, "frame": [n1, common.sha(a1), common.sha(l1), common.sha(pr)],
                           "map": [n2, common.sha(a2), common.sha(l2)], "init": [n3, common.sha(m12), common.sha(prev)],
                           "bow": [n4, common.sha(ab)]})
        print(cfg, n1, n2, n3, n4)
    json.dump(m, open(os.path.join(ROOT, "tests", "golden", "match_oracle.json"), "w"), indent=1)

    # independent twins where they exist (cv2-backed stereo windows, cv2.undistortPoints), oracle otherwise
    import next_cases
    nr = {"source": "stereo: tests/stereo_twin.py (cv2 window arithmetic); undistort: cv2.undistortPoints; others: oracle regression",
          "stereo_small": next_cases.stereo_case("twin"), "undistort_tum1": next_cases.undistort_case("twin"),
          "bow": next_cases.bow_case("oracle"), "matcher_tum": next_cases.matcher_cases(next_cases.Inputs("tum"), "oracle")}
    json.dump(nr, open(os.path.join(ROOT, "tests", "golden", "next_rows.json"), "w"), indent=1)
    print("next rows:", {k: (v if not isinstance(v, dict) else "...") for k, v in nr.items() if k != "source"})


if __name__ == "__main__":
    main()
