#!/usr/bin/env python3
"""Golden digests of the extraction of every frame of bench.py's frame pools (KITTI and TUM, 257 frames each: the pool the
headline number cycles through), produced by THE REFERENCE'S OWN ORBextractor.cc (oracle/_ref/libref.so, monotonic
allocation order = the B-1 pin) and cross-checked against the oracle on every frame.  Run in the build container (needs
/root/reference); writes tests/golden/sweep_ref.json, which tests/test_gpu_sweep.py holds the CUDA path to on the B200.

Per frame: n keypoints, sha256(keypoint records)[:16], sha256(descriptors)[:16]."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("slam-dynamic_b200", "tests", ""):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import bench
import orc
import ref


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    out = {}
    ref.set_alloc_mode(ref.ALLOC_BUMP)
    for cfg in ("kitti", "tum"):
        W, H, nrect, nf, ini, mn, _ = bench.WORKLOADS[cfg]
        frames = bench.make_frames(cfg, 0, 0, bench.POOL + 1)
        R = ref.Extractor(nf, bench.SCALE, bench.NLEVELS, ini, mn)
        O = orc.Extractor(nf, bench.SCALE, bench.NLEVELS, ini, mn)
        rows = []
        for i, img in enumerate(frames):
            k, d = R(img)
            ok, od = O(img)
            assert k.tobytes() == ok.tobytes() and np.array_equal(d, od), (cfg, i)
            rows.append([len(k), digest(k), digest(d)])
            if i % 32 == 0:
                print(cfg, i, rows[-1], flush=True)
        out[cfg] = {"frames": len(frames), "pool_digest": digest(frames), "rows": rows}
    with open(os.path.join(ROOT, "tests", "golden", "sweep_ref.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote tests/golden/sweep_ref.json")


if __name__ == "__main__":
    main()
