"""Seeded synthetic BAR workloads shared by bench.py's two arms and the tests.

Deliberately NOT part of the engine: `workload/libbarsynth.so` is a plain host library (g++), so that the reference arm
of bench.py, which must not load any product code, can generate the very same inputs as the GPU arm.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libbarsynth.so")
_LIB = None


def build(force=False):
    src = os.path.join(_HERE, "synth.cpp")
    if force or not os.path.exists(_LIBPATH) or os.path.getmtime(src) > os.path.getmtime(_LIBPATH):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", _LIBPATH, src])
    return _LIBPATH


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIBPATH):
            build()
        lib = C.CDLL(_LIBPATH)
        vp, i64, ci = C.c_void_p, C.c_int64, C.c_int
        lib.barsynth_end.argtypes = [C.c_uint64, C.c_uint64, ci, ci, C.c_double, C.c_double, C.c_double, vp, vp]
        lib.barsynth_end.restype = i64
        lib.barsynth_pair.argtypes = [C.c_uint64, C.c_uint64, ci, C.c_double, C.c_double, C.c_double, ci, vp, vp, vp, vp, vp]
        lib.barsynth_pair.restype = i64
        _LIB = lib
    return _LIB


def synth_ends(first_end, n_ends, K, L, seed=0xBA5E0000, sub=0.02, ins=0.005, dele=0.005):
    """Seeded synthetic ends -> (n_seq[int32 n], lens[int32 n*K], flat uint8 codes 0..3). Deterministic in (seed, end index)."""
    lib = _lib()
    n_seq = np.full(n_ends, K, np.int32)
    lens = np.zeros(n_ends * K, np.int32)
    flat = np.zeros(n_ends * K * (2 * L + 16), np.uint8)
    o = 0
    for e in range(n_ends):
        n = lib.barsynth_end(seed, first_end + e, K, L, sub, ins, dele, flat.ctypes.data + o, lens.ctypes.data + 4 * e * K)
        if n < 0:
            raise ValueError("barsynth_end failed")
        o += n
    return n_seq, lens, flat[:o].copy()


def synth_pairs(first_pair, n_pairs, L, k_anchor=50, seed=0xBA5E0000, sub=0.02, ins=0.005, dele=0.005):
    """Seeded synthetic sequence pairs with MUM-like anchors -> list of (sX, sY, anchors[n, 2], False, False)."""
    lib = _lib()
    bx, by = C.create_string_buffer(2 * L + 16), C.create_string_buffer(2 * L + 16)
    an = np.zeros((2 * L, 2), np.int64)
    lx, ly = C.c_int64(), C.c_int64()
    out = []
    for i in range(n_pairs):
        na = lib.barsynth_pair(seed, first_pair + i, L, sub, ins, dele, k_anchor, bx, C.byref(lx), by, C.byref(ly), an.ctypes.data)
        if na < 0:
            raise ValueError("barsynth_pair failed")
        out.append((bx.raw[:lx.value], by.raw[:ly.value], an[:na].copy(), False, False))
    return out
