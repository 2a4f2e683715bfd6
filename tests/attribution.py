"""Attribution of descriptor mismatches to rounding boundaries (TEST INFRASTRUCTURE ONLY).

north_star: "at least 99.9 % of descriptors bit-identical, with every mismatch traced to an angle-quantisation boundary".
computeOrbDescriptor (src/ORBextractor.cc:108-147) samples the blurred level at cvRound(x*b + y*a), cvRound(x*a - y*b) with
a = cos(angle), b = sin(angle) in float: the only step of the path that is not integer-exact is cos/sin, so two correct
implementations can only disagree on a bit whose sample coordinate sits within a few float ulps of a half-integer (where
cvRound flips).  `attribute` checks exactly that for every differing bit."""
import os
import re
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32
FACTOR_PI = f32(np.pi / 180.0)          # ORBextractor.cc:107


def pattern():
    """256 x (x0, y0, x1, y1) from include/sdyn_brief_pattern.inc (bit k of the descriptor compares sample 0 < sample 1)."""
    txt = open(os.path.join(os.path.dirname(_HERE), "include", "sdyn_brief_pattern.inc")).read()
    body = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    v = np.array([int(x) for x in re.findall(r"-?\d+", body)], np.int32)
    assert len(v) == 1024
    return v.reshape(256, 4)


def sample_coordinates(angle_deg):
    """Unrounded rotated coordinates of the 512 pattern points for one keypoint, float32 arithmetic without contraction
    (SURVEY Appendix B-3): returns (col, row) arrays of shape (256, 2)."""
    ang = f32(f32(angle_deg) * FACTOR_PI)
    a, b = f32(np.cos(np.float64(ang))), f32(np.sin(np.float64(ang)))
    p = pattern().astype(np.float32)
    xs, ys = p[:, [0, 2]], p[:, [1, 3]]
    row = (xs * b).astype(f32) + (ys * a).astype(f32)
    col = (xs * a).astype(f32) - (ys * b).astype(f32)
    return col.astype(f32), row.astype(f32)


def attribute(angle_deg, desc_a, desc_b, ulps=8):
    """For every bit on which two descriptors of the same keypoint differ: is one of the four rounded coordinates of that
    bit's point pair within `ulps` float ulps of a half-integer?  Returns [(bit, explained, margin)]."""
    da, db = np.asarray(desc_a, np.uint8), np.asarray(desc_b, np.uint8)
    diff = np.nonzero(np.unpackbits(da ^ db, bitorder="little"))[0]
    col, row = sample_coordinates(angle_deg)
    out = []
    for bit in diff:
        v = np.concatenate([col[bit], row[bit]]).astype(np.float64)
        dist = np.abs(np.abs(v - np.floor(v)) - 0.5)                     # distance to the nearest half-integer
        tol = ulps * np.spacing(np.maximum(np.abs(v), 1.0).astype(f32)).astype(np.float64)
        out.append((int(bit), bool((dist <= tol).any()), float((dist / tol).min())))
    return out


def all_explained(angles, desc_a, desc_b, ulps=8):
    """-> (n_mismatching_descriptors, n_unexplained_bits, detail) over arrays of keypoints."""
    rows = np.nonzero((np.asarray(desc_a) != np.asarray(desc_b)).any(1))[0]
    bad = []
    for i in rows:
        for bit, ok, margin in attribute(angles[i], desc_a[i], desc_b[i], ulps):
            if not ok:
                bad.append((int(i), bit, margin))
    return len(rows), len(bad), bad
