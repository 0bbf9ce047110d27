"""Builds an emulation harness (tests/emu_*.cpp: a .cuh of the product compiled for the host under tests/cuda_emu.h)
into tests/_build/. With GB200_EMU_SANITIZE=1 the harness is built with AddressSanitizer + UBSan -- every global and
shared-memory access of the kernels is then bounds-checked on the host; run such a session as

    GB200_EMU_SANITIZE=1 ASAN_OPTIONS=detect_leaks=0 LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
        python -m pytest tests -m "not gpu" -k emulated
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build(name, srcs):
    """srcs[0] = the .cpp to compile; the others are its dependencies (rebuilt when any of them is newer)."""
    san = os.environ.get("GB200_EMU_SANITIZE") == "1"
    os.makedirs(BUILD, exist_ok=True)
    lib = os.path.join(BUILD, "lib%s%s.so" % (name, "_asan" if san else ""))
    if not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in srcs):
        flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if san else ["-O1"]
        subprocess.check_call(["g++", *flags, "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-o", lib, srcs[0]])
    return ctypes.CDLL(lib)
