"""Host-side edges of the dynamic-keypoint path through the C ABI (no GPU needed): the YOLO detection file format
(Examples/RGB-D/rgbd_my.cc:232-252) and Frame::boxTrack (src/Frame.cc:481-552) against the oracle."""
import numpy as np

import orc
import pysdyn


def test_parse_detection_file(tmp_path):
    text = "0 100.5 50.25 40 20\n\n2 10 5 40 30\r\n7 300 200 0 0\nbroken line\n1 639.5 479.5 1 1"
    b = pysdyn.boxes_parse(text)
    expect = np.array([[80.5, 40.25, 40, 20], [0, 0, 40, 30], [300, 200, 0, 0], [639.0, 479.0, 1, 1]])
    assert np.array_equal(b, expect)                        # MAX(cx - w/2, 0) clamps, w / h pass through
    p = tmp_path / "000001.txt"
    p.write_text(text)
    assert np.array_equal(pysdyn.boxes_read(p), expect)
    assert len(pysdyn.boxes_read(tmp_path / "missing.txt")) == 0      # no file = no boxes for that frame
    assert len(pysdyn.boxes_parse("")) == 0
    try:
        pysdyn.boxes_parse(text, cap=2)
        raise AssertionError("capacity overflow not reported")
    except pysdyn.SdynError as e:
        assert e.code == -3


def random_boxes(r, n, W, H):
    x = r.uniform(0, W - 60, n); y = r.uniform(0, H - 60, n)
    w = r.uniform(20, 200, n); h = r.uniform(20, 150, n)
    return np.stack([x, y, w, h], 1)


def test_box_track_matches_oracle_over_sequences():
    r = np.random.default_rng(11)
    W, H = 640, 480
    for trial in range(60):
        objs = np.zeros((0, 4)); idx = np.zeros(0, np.int32); omit = np.zeros(0, np.uint8); vel = np.zeros((0, 2))
        oobjs, oidx, oomit, ovel = objs, idx, omit, vel
        truth = random_boxes(r, int(r.integers(0, 9)), W, H)
        for t in range(8):
            truth = truth + r.normal(0, 6, truth.shape) * np.array([1, 1, 0.2, 0.2])
            keep = r.random(len(truth)) > 0.2                # detector misses some boxes in some frames
            det = truth[keep]
            if r.random() < 0.3:
                det = np.concatenate([det, random_boxes(r, int(r.integers(1, 3)), W, H)])     # new objects
            if r.random() < 0.1:
                det = np.concatenate([det, [[50, 50, 0, 0]]])                                   # degenerate box
            got = pysdyn.box_track(det, objs, idx, omit, vel, W, H)
            ref = orc.box_track(det, oobjs, oidx, oomit, ovel, W, H)
            for g, e in zip(got, ref):
                assert np.array_equal(g, e), (trial, t)
            objs, idx, omit, vel = got
            oobjs, oidx, oomit, ovel = ref
            if r.random() < 0.1:                             # tracking lost: next frame renumbers from zero
                objs = np.zeros((0, 4)); idx = np.zeros(0, np.int32); omit = np.zeros(0, np.uint8); vel = np.zeros((0, 2))
                oobjs, oidx, oomit, ovel = objs, idx, omit, vel
