"""The drop-in C++ adapter classes (slam-dynamic_b200/host/: ORB_SLAM2::ORBextractor with the reference's
signature, and the ORBmatcher glue templates) run on the GPU through a small C++ program; its dumped inputs are
replayed through the CPU oracle here."""
import os
import struct
import subprocess

import numpy as np
import pytest

import orc
import pysdyn
import scenario

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_adapter")


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if len(h) < 4:
                break
            name = f.read(struct.unpack("<I", h)[0]).decode()
            nbytes = struct.unpack("<Q", f.read(8))[0]
            out[name] = f.read(nbytes)
    return out


def test_adapter_library_exports_reference_symbols():
    """CPU check: the adapter library was built and exports ORB_SLAM2::ORBextractor's constructor and operator()."""
    lib = os.path.join(ROOT, "slam-dynamic_b200", "libsdyn_host.so")
    assert os.path.exists(lib), "run __graft_entry__.build()"
    syms = subprocess.check_output(["nm", "-DC", lib]).decode()
    assert "ORB_SLAM2::ORBextractor::ORBextractor(int, float, int, int, int)" in syms
    assert "ORB_SLAM2::ORBextractor::operator()(cv::_InputArray const&, cv::_InputArray const&, std::vector<cv::KeyPoint" in syms


def test_adapter_templates_typecheck():
    """CPU check: every adapter template of host/sdyn_adapters.hpp — including the ones that need OpenCV's matrix algebra
    and are therefore not instantiated by tests/cpp/test_adapter.cc — instantiates against declaration-only mocks with the
    reference's member names (tests/cpp/adapter_syntax.cc)."""
    src = os.path.join(ROOT, "tests", "cpp", "adapter_syntax.cc")
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-Wextra", src], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_adapter_classes_match_oracle(tmp_path):
    out = str(tmp_path / "adapter.bin")
    r = subprocess.run([BIN, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    d = read_dump(out)
    W, H = 640, 480
    img1 = np.frombuffer(d["img1"], np.uint8).reshape(H, W)
    E = orc.Extractor(1000, 1.2, 8, 20, 7)
    k, desc = E(img1)
    ck = np.frombuffer(d["cur_keys"], pysdyn.KP_DTYPE); cd = np.frombuffer(d["cur_desc"], np.uint8).reshape(-1, 32)
    assert len(ck) == len(k)
    for name in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(ck[name], k[name]), name
    assert np.max(np.abs(ck["angle"] - k["angle"])) <= 1e-3 and (cd == desc).all(1).mean() >= 0.999
    lw, lh = struct.unpack("<ii", d["level3_dims"])
    assert np.array_equal(np.frombuffer(d["level3"], np.uint8).reshape(lh, lw), E.level(3)[19:-19, 19:-19])   # mvImagePyramid[3]
    scale = np.frombuffer(d["scale"], np.float32)
    assert np.array_equal(scale, E.scale)

    lk = np.frombuffer(d["last_keys"], pysdyn.KP_DTYPE); ld = np.frombuffer(d["last_desc"], np.uint8).reshape(-1, 32)
    lp = np.frombuffer(d["last_points"], pysdyn.LASTPOINT_DTYPE)
    cam = scenario.KITTI_CAM
    camt = (cam["fx"], cam["fy"], cam["cx"], cam["cy"], np.float32(379.8145), np.float32(379.8145) / np.float32(cam["fx"]))
    cur = pysdyn.FrameView(ck, cd, scale, (0.0, 0.0, float(W), float(H)), u_right=np.full(len(ck), -1, np.float32), cam=camt)
    last = pysdyn.FrameView(lk, ld, scale, (0.0, 0.0, float(W), float(H)), u_right=np.full(len(lk), -1, np.float32), cam=camt)
    n1, a1, l1, pairs = orc.match_projection_frame(cur, last, lp, 15.0, True, True, want_pairs=True)
    assert struct.unpack("<i", d["frame_n"])[0] == n1 and n1 > 100
    assert np.array_equal(np.frombuffer(d["frame_assign"], np.int32), a1)
    assert np.array_equal(np.frombuffer(d["frame_pairs"], np.float32).reshape(-1, 4), pairs)

    mp = np.frombuffer(d["map_points"], pysdyn.MAPPOINT_DTYPE)
    a0 = np.where(a1 >= 0, -2, -1).astype(np.int32)
    n2, a2, l2 = orc.match_projection_map(cur, mp, 3.0, 0.8, a0, l1)
    expect = np.where(a2 >= 0, 100000 + a2, np.where(a2 == -2, a1, -1)).astype(np.int32)
    assert struct.unpack("<i", d["map_n"])[0] == n2 and n2 > 50
    assert np.array_equal(np.frombuffer(d["map_assign"], np.int32), expect)

    # context re-creation keeps the camera model: undistorted keypoints of the wide image after wide -> tall -> wide
    rk = np.frombuffer(d["recreate_keys"], pysdyn.KP_DTYPE); rku = np.frombuffer(d["recreate_keys_un"], pysdyn.KP_DTYPE)
    rimg = np.frombuffer(d["recreate_img"], np.uint8).reshape(240, 640)
    ek, _ = orc.Extractor(500, 1.2, 8, 20, 7)(rimg)
    assert len(ek) == len(rk) and np.array_equal(ek["x"], rk["x"]) and np.array_equal(ek["y"], rk["y"])
    exy = orc.undistort_points(np.stack([ek["x"], ek["y"]], 1), np.float32(517.306408), np.float32(516.469215), np.float32(318.643040),
                               np.float32(255.313989), np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32))
    assert np.array_equal(rku["x"].view(np.uint32), exy[:, 0].view(np.uint32)) and np.array_equal(rku["y"].view(np.uint32), exy[:, 1].view(np.uint32))

    # stereo constructor path: Frame::ComputeStereoMatches through the adapter
    imgR = np.frombuffer(d["imgR"], np.uint8).reshape(H, W)
    EL, ER = orc.Extractor(1000, 1.2, 8, 20, 7), orc.Extractor(1000, 1.2, 8, 20, 7)
    kl, dl = EL(img1); kr, dr = ER(imgR)
    mbf = np.float32(379.8145); mb = mbf / np.float32(cam["fx"])
    ur, dp, kept = orc.stereo_matches(EL, ER, kl, dl, kr, dr, float(mb), float(mbf))
    assert kept > 200
    assert np.array_equal(np.frombuffer(d["stereo_uright"], np.float32).view(np.uint32), ur.view(np.uint32))
    assert np.array_equal(np.frombuffer(d["stereo_depth"], np.float32).view(np.uint32), dp.view(np.uint32))
