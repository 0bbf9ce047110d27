"""JPEG oracle (oracle/jpeg_oracle.c): the reference's own assertions (issue35.jpg loads, issue46.jpg
fails; examples/test-suite/source/main.d:28-49) plus sanity cross-checks against libjpeg-turbo. Pixel
values are UNPINNED by the reference; libjpeg's ISLOW IDCT runs columns-then-rows (jpgd: rows-then-
columns), so 4:4:4 agrees within a few LSB, not exactly; 4:2:0 uses a different upsampler altogether."""
import io
import math
import os

import numpy as np
import pytest
from PIL import Image as PILImage

from jpegutil import encode, photo

G = os.path.join(os.path.dirname(__file__), "golden")


def pil(data, mode):
    return np.asarray(PILImage.open(io.BytesIO(data)).convert(mode)).astype(int)


def test_issue35_loads_issue46_fails(oracle):
    data = open(os.path.join(G, "issue35.jpg"), "rb").read()
    r = oracle.jpeg_load(data, 4)        # LOAD_RGB|LOAD_ALPHA -> 4 components (main.d:28-35)
    assert r is not None
    px, ac, par, dpi = r
    assert px.shape == (235, 232, 4) and ac == 3 and (px[:, :, 3] == 255).all()
    assert par == 1.0 and dpi == 72.0
    assert np.abs(px[:, :, :3].astype(int) - pil(data, "RGB")).max() <= 4
    assert oracle.jpeg_load(b"", -1) is None                    # issue46.jpg is an empty file
    assert oracle.jpeg_load(open(os.path.join(G, "issue76.png"), "rb").read(), -1) is None


@pytest.mark.parametrize("ss", [0, 1, 2])
def test_colour_close_to_libjpeg(oracle, ss):
    img = photo(150, 203, 3, 5 + ss)
    data = encode(img, 92, ss)
    px, ac, par, dpi = oracle.jpeg_load(data, -1)
    assert ac == 3 and px.shape == img.shape
    d = np.abs(px.astype(int) - pil(data, "RGB"))
    assert d.max() <= (4 if ss == 0 else 60) and d.mean() < (0.1 if ss == 0 else 3.0)   # jpgd replicates / freq-upsamples chroma, libjpeg interpolates


def test_grey_and_req_comps(oracle):
    g = photo(97, 131, 1, 3)
    data = encode(g, 85)
    px, ac, par, dpi = oracle.jpeg_load(data, -1)
    assert ac == 1 and px.shape == (97, 131, 1)
    assert np.abs(px[:, :, 0].astype(int) - pil(data, "L")).max() <= 1
    p3 = oracle.jpeg_load(data, 3)[0]
    p4 = oracle.jpeg_load(data, 4)[0]
    assert (p3 == px).all() and (p4[:, :, :3] == px).all() and (p4[:, :, 3] == 255).all()
    img = photo(64, 80, 3, 4)
    data = encode(img, 90, 2)
    rgb = oracle.jpeg_load(data, 3)[0].astype(int)
    y = oracle.jpeg_load(data, 1)[0][:, :, 0]
    assert np.array_equal(y, ((rgb[:, :, 0] * 19595 + rgb[:, :, 1] * 38470 + rgb[:, :, 2] * 7471 + 32768) >> 16))
    assert oracle.jpeg_load(data, 2) is None                    # req_comps 2 is rejected (jpegload.d:3727)


def test_restart_markers_do_not_change_pixels(oracle):
    img = photo(120, 176, 3, 9)
    for ss in (0, 1, 2):
        a = oracle.jpeg_load(encode(img, 88, ss), -1)[0]
        b = oracle.jpeg_load(encode(img, 88, ss, restart_rows=1), -1)[0]
        c = oracle.jpeg_load(encode(img, 88, ss, restart_blocks=3), -1)[0]
        assert np.array_equal(a, b) and np.array_equal(a, c)


def test_density_and_nan_without_jfif(oracle):
    img = photo(16, 16, 3, 1)
    px, ac, par, dpi = oracle.jpeg_load(encode(img, 80, 0, dpi=(300, 150)), -1)
    assert dpi == 150.0 and par == 2.0
    data = encode(img, 80, 0)
    # strip the APP0 segment: no density at all -> D float.init (NaN) in the reference
    i = data.index(b"\xff\xe0")
    n = int.from_bytes(data[i + 2:i + 4], "big")
    px, ac, par, dpi = oracle.jpeg_load(data[:i] + data[i + 2 + n:], -1)
    assert math.isnan(par) and math.isnan(dpi)


def test_progressive_equals_baseline(oracle):
    """Progressive (SOF2, jpegload.d:3299-3683) and sequential files written by libjpeg from the same pixels with the
    same quality hold the same quantised coefficients (progressive mode only changes the entropy coding), so the
    restated progressive path (10 scans: DC first/refine, AC first/refine with EOB runs) must reproduce the pixels of
    the sequential path -- which is pinned to the reference's source text -- exactly. Every subsampling, grey,
    optimised tables, restart intervals inside the scans."""
    for (h, w, c, ss, q, kw) in ((48, 64, 3, 2, 90, {}), (77, 130, 3, 0, 75, {}), (150, 200, 3, 1, 85, {}), (61, 97, 1, 0, 90, {}),
                                  (250, 333, 3, 2, 95, {"optimize": True}), (256, 256, 3, 2, 50, {"restart_blocks": 7}),
                                  (173, 211, 3, 1, 80, {"restart_rows": 2}), (9, 7, 3, 2, 60, {}), (8, 8, 1, 0, 100, {})):
        img = photo(h, w, c, 3 + h)
        base = oracle.jpeg_load(encode(img, q, ss, **kw), -1)
        prog_file = encode(img, q, ss, progressive=True, **kw)
        assert prog_file.count(b"\xff\xda") > 1
        prog = oracle.jpeg_load(prog_file, -1)
        assert base is not None and prog is not None
        assert prog[1:] == base[1:]
        assert np.array_equal(np.asarray(prog[0]), np.asarray(base[0])), (h, w, c, ss)
    # channel adaptation on the progressive path
    f = encode(photo(40, 56, 3, 9), 85, 2, progressive=True)
    for rc in (1, 3, 4):
        assert np.array_equal(np.asarray(oracle.jpeg_load(f, rc)[0]), np.asarray(oracle.jpeg_load(encode(photo(40, 56, 3, 9), 85, 2), rc)[0]))


def test_progressive_against_libjpeg(oracle):
    """Independent sanity check of the progressive entropy layer: libjpeg-turbo's own decode of the same file (it runs
    the same LL&M IDCT in the other pass order; 4:2:0 chroma is upsampled differently)."""
    for (c, ss, tol) in ((1, 0, 1), (3, 0, 4), (3, 2, 60)):
        data = encode(photo(72, 88, c, 5), 90, ss, progressive=True)
        px = np.asarray(oracle.jpeg_load(data, -1)[0])
        ref = np.asarray(PILImage.open(io.BytesIO(data)))
        assert np.abs(px.reshape(ref.shape).astype(int) - ref.astype(int)).max() <= tol


def test_progressive_corrupt_and_truncated(oracle):
    img = photo(40, 40, 3, 2)
    prog = encode(img, 90, 2, progressive=True)
    # cut inside the first scan: the missing scans simply never refine the coefficients; the data that is there runs
    # into the FF D9 padding of get_char (jpegload.d:640-655) and decodes as 1-bits -- an image comes back or the
    # decode fails, but nothing crashes
    for cut in (200, len(prog) // 2, len(prog) - 3):
        oracle.jpeg_load(prog[:cut], -1)
    # a refinement scan whose successive-approximation pair is inconsistent is rejected (:3637-3641)
    i = prog.rindex(b"\xff\xda")
    n = int.from_bytes(prog[i + 2:i + 4], "big")
    bad = bytearray(prog)
    bad[i + 2 + n - 1] = 0x31                 # Ah = 3, Al = 1
    assert oracle.jpeg_load(bytes(bad), -1) is None
    data = encode(img, 90, 2)
    assert oracle.jpeg_load(data[:200], -1) is None
