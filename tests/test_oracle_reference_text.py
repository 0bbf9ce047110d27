"""The C oracle against golden vectors generated from the reference's SOURCE TEXT
(tests/golden/gen_from_reference.py: the D function bodies of jpegload.d, scanline.d, qoiplane10.d, qoi2avg.d and
qoi10b.d transliterated mechanically and executed with exact int32 / IEEE-single semantics). Exact equality
everywhere: this is what pins the JPEG, PixelType-converter and QOIX-predictor oracles to the reference."""
import ctypes as C
import os

import numpy as np
import pytest

from gamut_b200.types import PixelType as PT

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_jpeg_idct_sparse_paths(oracle):
    """idct (jpegload.d:308-376) through the reference's own Row!N / Col!N instantiations for every max_zag 1..64:
    the oracle's dense evaluation must give the same 64 bytes."""
    L = oracle.lib()
    d = _load("ref_jpeg_idct.npz")
    blocks, mz, exp = d["blocks"], d["max_zag"], d["out"]
    assert blocks.shape[0] == 64 * 24 and set(mz.tolist()) == set(range(1, 65))
    got = np.zeros(64, np.uint8)
    for i in range(blocks.shape[0]):
        b = np.ascontiguousarray(blocks[i])
        L.or_test_jpeg_idct(b.ctypes.data_as(C.c_void_p), int(mz[i]), got.ctypes.data_as(C.c_void_p))
        assert np.array_equal(got, exp[i]), (i, int(mz[i]))


def test_jpeg_dct_upsample(oracle):
    """DCT_Upsample.P_Q!(R,C) / R_S!(R,C) (jpegload.d:914-1069, the hand-expanded lines), the s_max_rc dispatch and
    idct_4x4, for every max_zag: the oracle's looped dense form must give the same four 8x8 tiles."""
    L = oracle.lib()
    d = _load("ref_jpeg_upsample.npz")
    blocks, exp = d["blocks"], d["out"]
    assert blocks.shape[0] == 64 * 16
    got = np.zeros(256, np.uint8)
    for i in range(blocks.shape[0]):
        b = np.ascontiguousarray(blocks[i])
        L.or_test_jpeg_upsample(b.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p))
        assert np.array_equal(got.reshape(4, 64), exp[i]), i


def test_jpeg_colour(oracle):
    """create_look_ups (jpegload.d:2079-2094) and the YCbCr -> RGB pixel expression of H1V1Convert (:2543-2548)."""
    L = oracle.lib()
    d = _load("ref_jpeg_colour.npz")
    tables = np.zeros(1024, np.int32)
    rgb = np.zeros(3, np.uint8)
    for (y, cb, cr), e in zip(d["ycc"], d["rgb"]):
        L.or_test_jpeg_ycc(int(y), int(cb), int(cr), rgb.ctypes.data_as(C.c_void_p), tables.ctypes.data_as(C.c_void_p))
        assert np.array_equal(rgb, e), (y, cb, cr)
    for k, name in enumerate(("crr", "cbb", "crg", "cbg")):
        assert np.array_equal(tables[k * 256:(k + 1) * 256], d[name]), name


def _scan_cases():
    d = _load("ref_scanline.npz")
    return d, [str(n) for n in d["names"]]


@pytest.mark.parametrize("name", _scan_cases()[1])
def test_scanline_function(oracle, name):
    """Every scanline_convert_* body of scanline.d through or_scanlinesConvert (single-stage pairs: the source or
    the destination is the intermediate type, so exactly that reference function runs)."""
    d, _ = _scan_cases()
    src_t, dst_t = name.split("_to_")
    if src_t in ("bgra8", "bgr8") or dst_t in ("bgra8", "bgr8") or (src_t, dst_t) == ("l8", "rgb8"):
        fn = getattr(oracle.lib(), "or_scanline_" + name)           # the three helpers outside scanlinesConvert
        x = np.ascontiguousarray(d["in_" + name])
        out = np.zeros_like(d["out_" + name])
        fn(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), int(d["width"]))
        assert np.array_equal(out, d["out_" + name])
        return
    w = int(d["width"])
    x = np.ascontiguousarray(d["in_" + name]).view(np.uint8)
    exp = np.ascontiguousarray(d["out_" + name]).view(np.uint8)
    got = np.zeros_like(exp)
    assert oracle.scanlines_convert(PT[src_t], x, x.size, PT[dst_t], got, got.size, w, 1)
    assert np.array_equal(got, exp), name


def test_predictors(oracle):
    L = oracle.lib()
    d = _load("ref_predictors.npz")
    for (a, b, c), e in zip(d["loco10_in"].tolist(), d["loco10_out"].tolist()):
        assert L.or_test_loco_predict10(a, b, c) == e, (a, b, c)
    for px, e in zip(d["loco8_in"].tolist(), d["loco8_out"].tolist()):
        for ch in range(4):
            assert L.or_test_loco8(px[0][ch], px[1][ch], px[2][ch]) == e[ch], (px, ch)
    for px, e in zip(d["loco10b_in"].tolist(), d["loco10b_out"].tolist()):
        for ch in range(4):
            assert L.or_test_loco10(px[0][ch], px[1][ch], px[2][ch]) == e[ch], (px, ch)
