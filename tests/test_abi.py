"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gamut_b200.h declares; without a GPU compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gamut_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gb200_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(gb):
    from gamut_b200 import _lib
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 10
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_enum_values_match_reference_abi(gb):
    from gamut_b200.types import PixelType as PT
    from gamut_b200 import _lib
    L = _lib.lib()
    assert [int(t) for t in PT][1:] == list(range(18)) and int(PT.unknown) == -1
    sizes = [1, 2, 4, 2, 4, 8, 2, 4, 8, 3, 6, 12, 4, 8, 16, 4, 8, 16]   # types.d:62-86
    assert [L.gb200_pixel_type_size(t) for t in range(18)] == sizes
    assert L.gb200_scanlines_inter_type(PT.l8, PT.rgb8) == PT.rgba8
    assert L.gb200_scanlines_inter_type(PT.l8, PT.lap8) == PT.rgbaf32


def test_no_cpu_fallback_without_gpu(gb):
    import torch
    if torch.cuda.is_available():
        return
    from gamut_b200 import scanlinesConvert, last_error
    from gamut_b200.types import PixelType as PT
    src = np.zeros(16, np.uint8)
    dst = np.zeros(64, np.uint8)
    assert scanlinesConvert(PT.rgba8, src, 16, PT.rgbaf32, dst, 64, 4, 1) is False
    assert "no CPU fallback" in last_error() or "CUDA" in last_error()


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "gamut_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_header_is_plain_c99(tmp_path):
    """include/gamut_b200.h is the drop-in boundary: it must compile as C (no C++, no torch types) for cgo / D / ctypes
    style bindings; the structs a binding fills by position keep their field order."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "gamut_b200.h"\n'
                   'int main(void) { gb200_tga_desc t = {1, 1, 3, 9}; gb200_bmp_desc b = {1, 1, 3, 9, -1.f, -1.f};\n'
                   '  gb200_qoi_desc q = {1, 1, 3, 0}; gb200_qoix_desc x = {1, 1, 4, 4, 8, 0, 0, -1.f, -1.f};\n'
                   '  return (int)(sizeof(t) + sizeof(b) + sizeof(q) + sizeof(x) + sizeof(gb200_image)) == 0; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
