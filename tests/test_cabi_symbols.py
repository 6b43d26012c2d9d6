"""CPU: the C-ABI library loads and exports every symbol include/barb200.h declares (no compute without a GPU),
host-only entry points work, and creating a context without a CUDA device fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import workload  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cactus_b200 import build as b
    b.build()
    import cactus_b200 as cb
    return cb.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "barb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(barb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 18
    raw = C.CDLL(os.path.join(ROOT, "cactus_b200", "libbarb200.so"))
    for n in sorted(names):
        assert hasattr(raw, n), "missing export: " + n


def test_defaults_are_cactus_config(lib):
    import cactus_b200 as cb
    from cactus_b200.api import _CParams
    p = _CParams()
    lib.barb200_params_default(C.byref(p))
    assert list(p.mat) == cb.api.CACTUS_SUBMAT
    assert (p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2) == (400, 30, 1200, 1)
    assert p.wb == 1000 and abs(p.wf - 0.1) < 1e-7 and (p.k, p.w, p.min_w) == (15, 5, 500)
    assert p.progressive_poa == 1 and p.disable_seeding == 1


def test_synth_is_deterministic_and_sorted(lib):
    import cactus_b200 as cb
    a = workload.synth_ends(5, 3, 8, 300)
    b = workload.synth_ends(5, 3, 8, 300)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    lens = a[1].reshape(3, 8)
    assert np.all(lens[:, :-1] >= lens[:, 1:]) and a[2].max() <= 3
    c = workload.synth_ends(6, 1, 8, 300)
    assert np.array_equal(c[1], a[1][8:16])        # end index, not call order, defines the data


def test_bulk_release_and_row_gather(lib):
    """barb200_pack_rows / barb200_free_many: host-only helpers (no device needed)"""
    libc = C.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = C.c_void_p, [C.c_size_t]
    rng = np.random.default_rng(7)
    sizes = rng.integers(0, 5000, 300).astype(np.int64)
    sizes[[3, 77]] = 0
    rows = (C.c_void_p * len(sizes))()
    want = []
    for i, n in enumerate(sizes):
        data = rng.integers(0, 256, int(n)).astype(np.uint8)
        rows[i] = libc.malloc(max(1, int(n)))
        C.memmove(rows[i], data.ctypes.data, int(n))
        want.append(data)
    dst = np.full(int(sizes.sum()) + 8, 0xEE, np.uint8)
    lib.barb200_pack_rows(rows, sizes.ctypes.data, len(sizes), dst.ctypes.data)
    assert np.array_equal(dst[:-8], np.concatenate(want)) and np.all(dst[-8:] == 0xEE)
    lib.barb200_free_many(rows, len(sizes))
    lib.barb200_free_many(None, 0)
    lib.barb200_pack_rows(None, None, 0, None)


def test_no_cpu_fallback():
    import torch
    import cactus_b200 as cb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cb.BarB200Error):
        cb.Engine()


def test_product_does_not_touch_oracle():
    """the product tree must not reference the checkers"""
    for root, _, files in os.walk(os.path.join(ROOT, "cactus_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("the oracle", "") or f == "api.py" and False, f
                assert "_reflib" not in txt and "libpoa_oracle" not in txt and "libabpoa_ref" not in txt, f
