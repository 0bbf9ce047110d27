"""`python -m gamut_b200.convert -i in.{png,jpg,qoi,qoix} -o out.{pam,raw,npy}` -- the equivalent of the
reference's examples/convert (examples/convert/source/main.d:121-148) for the decode side of the path:

    image.loadFromFile(input, LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS)   (main.d:123)
    [optional] image.convertTo(type)                                   (main.d:130-139, `-b 8|10|16|f32`)
    save

Decoding and conversion run on the GPU through the C ABI; there is no CPU path. The reference's encoders
(saveToFile for PNG/JPEG/QOI/QOIX, main.d:141) are SURVEY 8(f1) "next" rows and are not built, so the decoded
pixels are written in container formats that need no codec: PAM (8/16-bit), raw bytes, or .npy.
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np


def save(im, path: str) -> None:
    from .types import pixelTypeNumChannels
    px = im.pixels()
    h, w = px.shape[:2]
    if path.endswith(".npy"):
        np.save(path, px)
    elif path.endswith(".pam"):
        if px.dtype == np.float32:
            raise SystemExit("convert: PAM holds 8/16-bit samples; use .raw or .npy for fp32 (or -b 8/16)")
        ch = pixelTypeNumChannels(im.type())
        tupl = {1: "GRAYSCALE", 2: "GRAYSCALE_ALPHA", 3: "RGB", 4: "RGB_ALPHA"}[ch]
        maxval = 255 if px.dtype == np.uint8 else 65535
        with open(path, "wb") as f:
            f.write(b"P7\nWIDTH %d\nHEIGHT %d\nDEPTH %d\nMAXVAL %d\nTUPLTYPE %s\nENDHDR\n" % (w, h, ch, maxval, tupl.encode()))
            f.write(px.astype(">u2").tobytes() if px.dtype == np.uint16 else px.tobytes())
    elif path.endswith(".raw"):
        with open(path, "wb") as f:
            f.write(px.tobytes())
    else:
        raise SystemExit("convert: output must be .pam, .raw or .npy (the PNG/JPEG/QOI/QOIX encoders are not "
                         "part of the decode hot path: SURVEY 8(f1))")


def main(argv=None) -> int:
    from .image import Image, convertPixelTypeTo8Bit, convertPixelTypeTo16Bit, convertPixelTypeToFP32
    from .types import LAYOUT_GAPLESS, LAYOUT_VERT_STRAIGHT
    ap = argparse.ArgumentParser(prog="python -m gamut_b200.convert", description=__doc__.split("\n\n")[0])
    ap.add_argument("-i", "--input", required=True)
    ap.add_argument("-o", "--output", required=True)
    ap.add_argument("-b", "--bitdepth", choices=["8", "16", "f32"], default=None,
                    help="convert to this component depth before saving (examples/convert main.d:130-139)")
    ap.add_argument("-f", "--flags", type=lambda s: int(s, 0), default=0, help="extra LoadFlags (numeric)")
    a = ap.parse_args(argv)
    data = open(a.input, "rb").read()
    im = Image()
    t0 = time.perf_counter()
    im.loadFromMemory(data, a.flags | LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS)
    t1 = time.perf_counter()
    if im.isError():
        sys.stderr.write("convert: %s\n" % im.errorMessage())
        return 1
    if a.bitdepth:
        fn = {"8": convertPixelTypeTo8Bit, "16": convertPixelTypeTo16Bit, "f32": convertPixelTypeToFP32}[a.bitdepth]
        if not im.convertTo(fn(im.type()), LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS):
            sys.stderr.write("convert: %s\n" % im.errorMessage())
            return 1
    save(im, a.output)
    px = im.width() * im.height()
    print("Opened %s: %dx%d %s, decoded in %.3f ms (%.1f Mpixels/s) => %s" %
          (a.input, im.width(), im.height(), im.type().name, (t1 - t0) * 1e3, px / max(t1 - t0, 1e-9) / 1e6, a.output))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
