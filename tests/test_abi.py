"""CPU-side checks of the C-ABI boundary: the library loads, exports every entry point include/sdyn.h declares,
its host-only pieces agree with the oracle, and without a device it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import numpy as np

import orc
import pysdyn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sdyn.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sdyn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = pysdyn.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libsdyn.so does not export " + n


def test_struct_layouts_match_the_header():
    assert pysdyn.KP_DTYPE.itemsize == 28                      # cv::KeyPoint
    assert pysdyn.MAPPOINT_DTYPE.itemsize == 56 and pysdyn.LASTPOINT_DTYPE.itemsize == 48
    assert C.sizeof(pysdyn.OrbParams) == 20
    assert C.sizeof(pysdyn.FrameViewC) == 8 + 5 * 8 + 10 * 4 + 48


def test_host_tables_match_oracle():
    """ORBextractor constructor tables (scale chain in float*double, cvRound quotas, disc half-widths)."""
    lib = pysdyn.lib()
    lib.sdyn_orb_tables.argtypes = [C.POINTER(pysdyn.OrbParams), C.POINTER(pysdyn.ScaleInfo), C.POINTER(C.c_int32 * 16)]
    for nf, sf, nl in [(2000, 1.2, 8), (1000, 1.2, 8), (1500, 1.2, 8), (8000, 1.2, 8), (500, 1.5, 5), (4000, 2.0, 4), (1200, 1.1, 12)]:
        p = pysdyn.OrbParams(nf, sf, nl, 20, 7)
        si = pysdyn.ScaleInfo(); um = (C.c_int32 * 16)()
        assert lib.sdyn_orb_tables(C.byref(p), C.byref(si), C.byref(um)) == 0
        o = orc.Extractor(nf, sf, nl, 20, 7)
        assert np.array_equal(np.array(si.scale[:nl], np.float32), o.scale)
        assert np.array_equal(np.array(si.inv_scale[:nl], np.float32), o.inv_scale)
        assert np.array_equal(np.array(si.sigma2[:nl], np.float32), o.sigma2)
        assert np.array_equal(np.array(si.inv_sigma2[:nl], np.float32), o.inv_sigma2)
        assert list(si.features_per_level[:nl]) == o.quota.tolist()
        assert list(um) == o.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_quota_tables_of_the_survey():
    """mnFeaturesPerLevel for the shipped YAMLs."""
    assert orc.Extractor(2000, 1.2, 8, 12, 7).quota.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert orc.Extractor(1000, 1.2, 8, 20, 7).quota.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert orc.Extractor(8000, 1.2, 8, 20, 7).quota.tolist() == [1737, 1448, 1207, 1005, 838, 698, 582, 485]


def test_hamming_host_function():
    r = np.random.default_rng(0)
    a = r.integers(0, 256, (100, 32), dtype=np.uint8); b = r.integers(0, 256, (100, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert pysdyn.Matcher.DescriptorDistance(x, y) == orc.hamming(x, y)


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        return                                                # covered by the -m gpu suite on the B200 box
    try:
        pysdyn.Extractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    except pysdyn.SdynError as e:
        assert e.code == -2                                   # SDYN_ERR_CUDA
    else:
        raise AssertionError("sdyn_create succeeded without a CUDA device")


def test_bad_arguments_are_rejected_before_touching_the_device():
    lib = pysdyn.lib()
    h = C.c_void_p()
    for bad in [pysdyn.OrbParams(0, 1.2, 8, 20, 7), pysdyn.OrbParams(1000, 1.0, 8, 20, 7), pysdyn.OrbParams(1000, 1.2, 17, 20, 7),
                pysdyn.OrbParams(1000, 1.2, 8, 0, 7)]:
        assert lib.sdyn_create(C.byref(bad), 640, 480, 1, 0, C.byref(h)) == -1 and not h.value
    assert lib.sdyn_destroy(None) == 0 and lib.sdyn_last_error(None) is not None


def test_track_input_layout_offsets():
    """sdyn_track_input_layout: arrays in upload order, each on a 256-byte boundary; an array takes room only in its form."""
    import pysdyn
    for n, strides in [(1, (100, 50, 16)), (64, (2200, 3000, 256)), (3, (0, 0, 0))]:
        for forms in (0, 1, 2, 4, 6, 14):
            sep, rl, rm, fr = bool(forms & 1) and not forms & 2, bool(forms & 2), bool(forms & 4), bool(forms & 8)
            lay, total = pysdyn.track_input_layout(n, strides, forms)
            per_frame = {"last_points": 0 if rl else strides[0] * 48, "last_keys": 0 if rl else strides[0] * 28,
                         "last_keys_un": strides[0] * 28 if sep else 0, "n_last": 0 if rl else 4,
                         "map_points": 0 if rm else strides[1] * 56, "n_map": 4, "boxes": 64 * 32, "n_boxes": 4, "ref_box": 256,
                         "ref_desc": strides[2] * 32, "ref_xy": strides[2] * 8, "ref_off": 260, "fmat": 36, "poses": 96,
                         "last_ids": strides[0] * 4 if rl else 0, "last_flags": strides[0] if rl else 0,
                         "map_ids": strides[1] * 4 if rm else 0, "map_proj": strides[1] * 24 if rm and not fr else 0,
                         "map_flags": strides[1] if rm and fr else 0}
            off = 0
            for name in pysdyn.TRACK_ARRAYS:
                assert lay[name] == off and off % 256 == 0, (name, lay[name], off)
                off += (per_frame[name] * n + 255) // 256 * 256
            assert total == off
    import pytest
    with pytest.raises(pysdyn.SdynError):
        pysdyn.track_input_layout(0, (1, 1, 1))
