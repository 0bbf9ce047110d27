"""`python -m gamut_b200.convert input output [options]` -- the equivalent of the reference's examples/convert
(examples/convert/source/main.d:28-148), the plumbing case of BASELINE configs[0]:

    image.loadFromFile(input, LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS)                    (main.d:121)
    -b 8|16|auto, --rgb | --grey, --alpha | --drop-alpha, -p/--premul | --unpremul      (main.d:129-148)
    image.saveToFile(output)                                                            (main.d:150)

Decoding, conversion and encoding run on the GPU through the C ABI; there is no CPU path. saveToFile picks the format
from the output's extension like the reference (plugin.d:55-97); the encoders built so far are QOI (.qoi), QOIX for
8-bit images and 10-bit greyscale (.qoix) and TGA (.tga). For everything else -- the reference would write PNG / JPEG /
BMP / ... -- the decoded pixels can be written in containers that need no codec: PAM (8/16-bit), raw bytes or .npy.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np


def save_plain(im, path: str) -> bool:
    """PAM / raw / npy: the pixels as they are, no codec."""
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
        return False
    return True


def main(argv=None) -> int:
    from .image import Image, identifyImageFormatFromFilename
    from .types import ImageFormat, LAYOUT_GAPLESS, LAYOUT_VERT_STRAIGHT
    ap = argparse.ArgumentParser(prog="python -m gamut_b200.convert", description=__doc__.split("\n\n")[0])
    ap.add_argument("input")
    ap.add_argument("output")
    ap.add_argument("-b", "--bitness", choices=["8", "16", "auto"], default="auto")
    ap.add_argument("-p", "--premul", action="store_true")
    ap.add_argument("--unpremul", action="store_true")
    ap.add_argument("--grey", action="store_true")
    ap.add_argument("--rgb", action="store_true")
    ap.add_argument("--alpha", action="store_true")
    ap.add_argument("--drop-alpha", action="store_true")
    a = ap.parse_args(argv)
    if a.rgb and a.grey:
        raise SystemExit("Can't use --grey and --rgb at the same time")
    if a.alpha and a.drop_alpha:
        raise SystemExit("Can't use --alpha and --drop-alpha at the same time")
    if a.premul and a.unpremul:
        raise SystemExit("Cannot have both --premul and --unpremul")
    im = Image()
    t0 = time.perf_counter()
    im.loadFromFile(a.input, LAYOUT_VERT_STRAIGHT | LAYOUT_GAPLESS)
    t1 = time.perf_counter()
    if im.isError():
        sys.stderr.write("Couldn't open file %s: %s\n" % (a.input, im.errorMessage()))
        return 1
    steps = []
    if a.bitness == "8":
        steps.append(im.convertTo8Bit)
    elif a.bitness == "16":
        steps.append(im.convertTo16Bit)
    if a.rgb:
        steps.append(im.convertToRGB)
    elif a.grey:
        steps.append(im.convertToGreyscale)
    if a.alpha:
        steps.append(im.addAlphaChannel)
    elif a.drop_alpha:
        steps.append(im.dropAlphaChannel)
    if a.premul:
        steps.append(im.premultiply)
    if a.unpremul:
        steps.append(im.unpremultiply)
    for step in steps:
        if not step():
            sys.stderr.write("convert: %s\n" % im.errorMessage())
            return 1
    t2 = time.perf_counter()
    fif = identifyImageFormatFromFilename(a.output)
    if fif != ImageFormat.unknown:
        ok = im.saveToFile(a.output)
        if not ok:
            sys.stderr.write("Couldn't save file %s (the %s encoder is not built for %s images; .pam / .raw / .npy take any image)\n"
                             % (a.output, fif.name, im.type().name))
            return 1
    elif not save_plain(im, a.output):
        sys.stderr.write("Couldn't save file %s: unknown extension\n" % a.output)
        return 1
    t3 = time.perf_counter()
    px = im.width() * im.height()
    print("Opened %s: %dx%d, decoded in %.3f ms (%.1f Mpixels/s)" % (a.input, im.width(), im.height(), (t1 - t0) * 1e3, px / max(t1 - t0, 1e-9) / 1e6))
    print("Converted to %s in %.3f ms; encoded %s (%d bytes) in %.3f ms (%.1f Mpixels/s)" %
          (im.type().name, (t2 - t1) * 1e3, a.output, os.path.getsize(a.output), (t3 - t2) * 1e3, px / max(t3 - t2, 1e-9) / 1e6))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
