"""tests/cuda_emu.h checked against the documented semantics of the CUDA primitives it maps onto host threads: warp
shuffles and ballots (per 32-thread warp of a 96-thread CTA), __syncthreads with shared memory, shared atomics across
CTAs run one after the other, __clz / __ffs / __byte_perm. The emulated kernel tests (test_*_emulated.py) lean on these."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_primitives():
    import emu_build
    L = emu_build.build("emu_selfcheck", [os.path.join(HERE, "emu_selfcheck.cpp"), os.path.join(HERE, "cuda_emu.h")])
    rng = np.random.default_rng(0)
    bits_in = np.concatenate([[0, 1, 0x80000000, 0xffffffff, 0x00010000], rng.integers(0, 2 ** 32, 200)]).astype(np.uint32)
    n = len(bits_in)
    warp, cta, bits = np.zeros(96 * 4, np.uint32), np.zeros(9, np.uint32), np.zeros(4 * n, np.uint32)
    L.emu_selfcheck(warp.ctypes.data_as(C.c_void_p), cta.ctypes.data_as(C.c_void_p), bits_in.ctypes.data_as(C.c_void_p),
                    bits.ctypes.data_as(C.c_void_p), n)
    warp = warp.reshape(96, 4)
    for t in range(96):
        lane, w0 = t & 31, t & ~31
        assert warp[t, 0] == (t - 1 if lane >= 1 else t) and warp[t, 1] == (t - 5 if lane >= 5 else t)       # shfl_up keeps the own value below delta
        assert warp[t, 2] == sum(1 << l for l in range(32) if (w0 + l) % 3 == 0)                              # ballot is per warp
        assert warp[t, 3] == sum(range(w0, t + 1))                                                            # the scan idiom of the kernels
    for b in range(3):
        assert cta[3 * b] == sum(t + b for t in range(256)) and cta[3 * b + 1] == 0xffffffff and cta[3 * b + 2] == max(t * 7 % 251 for t in range(256))
    bits = bits.reshape(n, 4)
    for i, v in enumerate(bits_in.tolist()):
        assert bits[i, 0] == (32 if v == 0 else 32 - v.bit_length())
        assert bits[i, 1] == (0 if v == 0 else (v & -v).bit_length())
        assert bits[i, 2] == int.from_bytes(v.to_bytes(4, "little"), "big")
        src = v.to_bytes(4, "little") + ((~v) & 0xffffffff).to_bytes(4, "little")
        assert bits[i, 3] == int.from_bytes(bytes([src[1], src[3], src[5], src[7]]), "little")
