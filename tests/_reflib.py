"""ctypes readers for the CHECKER libraries (test infrastructure only).

* ``oracle/_ref/libabpoa_ref.so``  -- the unmodified reference abPOA + oracle/ref_harness.c
* ``oracle/_ref/libbar_ref.so``    -- the reference bar/impl/poaBarAligner.c + oracle/bar_ref_harness.c
* ``oracle/_build/libpoa_oracle.so`` -- the plain-C restatement (oracle/poa_oracle.c, bar_oracle.c)

Nothing in ``cactus_b200`` may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libabpoa_ref.so")
BAR_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbar_ref.so")
BAR_SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libbar_shim.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libpoa_oracle.so")

# Cactus defaults, src/cactus/cactus_progressive_config.xml:307-325
CACTUS_MAT = [91, -114, -61, -123, -100,
              -114, 100, -125, -61, -100,
              -61, -125, 100, -114, -100,
              -123, -61, -114, 91, -100,
              -100, -100, -100, -100, 100]


class RefParams(C.Structure):
    _fields_ = [("wb", C.c_int), ("wf", C.c_float),
                ("gap_open1", C.c_int), ("gap_ext1", C.c_int), ("gap_open2", C.c_int), ("gap_ext2", C.c_int),
                ("mat", C.c_int * 25),
                ("k", C.c_int), ("w", C.c_int), ("min_w", C.c_int),
                ("progressive_poa", C.c_int), ("disable_seeding", C.c_int)]


def cactus_params(wb=1000, wf=0.1, o1=400, e1=30, o2=1200, e2=1, mat=None, k=15, w=5, min_w=500,
                  progressive=1, disable_seeding=1):
    p = RefParams()
    p.wb, p.wf = wb, wf
    p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2 = o1, e1, o2, e2
    for i, v in enumerate(mat or CACTUS_MAT):
        p.mat[i] = v
    p.k, p.w, p.min_w = k, w, min_w
    p.progressive_poa, p.disable_seeding = progressive, disable_seeding
    return p


def params_dict(p):
    return dict(wb=p.wb, wf=float(p.wf), o1=p.gap_open1, e1=p.gap_ext1, o2=p.gap_open2, e2=p.gap_ext2,
                mat=list(p.mat), k=p.k, w=p.w, min_w=p.min_w, progressive=p.progressive_poa,
                disable_seeding=p.disable_seeding)


def have_ref():
    return os.path.exists(REF_SO)


def have_bar_ref():
    return os.path.exists(BAR_REF_SO)


def have_bar_shim():
    return os.path.exists(BAR_SHIM_SO)


def build_oracle():
    """(Re)build the plain-C oracle; cheap, and keeps tests independent of build() ordering."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORACLE_SO


_libs = {}


def _load(path):
    if path not in _libs:
        _libs[path] = C.CDLL(path)
    return _libs[path]


def _flat(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.int32)
    flat = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    return lens, np.ascontiguousarray(flat)


def _parse_trace(words, n_seq):
    w = words
    assert w[0] == n_seq
    msa_len, total_cells = int(w[1]), int(w[2])
    pos = 3
    read_id_map = [int(x) for x in w[pos:pos + n_seq]]
    pos += n_seq
    alns = []
    for _ in range(n_seq):
        read_id, qlen, node_n, n_cigar, best, n_rows = (int(x) for x in w[pos:pos + 6])
        pos += 6
        cigar = w[pos:pos + n_cigar].astype(np.uint64, copy=True) if n_cigar else np.zeros(0, np.uint64)
        pos += n_cigar
        dp_beg = w[pos:pos + n_rows].astype(np.int32)
        pos += n_rows
        dp_end = w[pos:pos + n_rows].astype(np.int32)
        pos += n_rows
        alns.append(dict(read_id=read_id, qlen=qlen, node_n=node_n, best_score=best,
                         cigar=cigar.view(np.uint64), dp_beg=dp_beg, dp_end=dp_end))
    msa = w[pos:].view(np.uint8)[: n_seq * msa_len].reshape(n_seq, msa_len).copy()
    return dict(msa=msa, msa_len=msa_len, cells=total_cells, read_id_map=read_id_map, alns=alns)


def _msa_fn(lib, name):
    f = getattr(lib, name)
    f.restype = C.c_int
    f.argtypes = [C.POINTER(RefParams), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    return f


def _trace_fn(lib, name):
    f = getattr(lib, name)
    f.restype = C.c_void_p
    f.argtypes = [C.POINTER(RefParams), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    return f


def _call_msa(lib, fname, freename, seqs, p):
    lens, flat = _flat(seqs)
    out = C.c_void_p()
    n = _msa_fn(lib, fname)(C.byref(p), len(seqs), lens.ctypes.data, flat.ctypes.data, C.byref(out))
    msa = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(len(seqs) * max(n, 1),))[: len(seqs) * n]
    msa = msa.reshape(len(seqs), n).copy()
    free = getattr(lib, freename)
    free.argtypes = [C.c_void_p]
    free(out)
    return msa


def _call_trace(lib, fname, freename, seqs, p):
    lens, flat = _flat(seqs)
    nw = C.c_int64()
    ptr = _trace_fn(lib, fname)(C.byref(p), len(seqs), lens.ctypes.data, flat.ctypes.data, C.byref(nw))
    words = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(nw.value,)).copy()
    free = getattr(lib, freename)
    free.argtypes = [C.c_void_p]
    free(ptr)
    return _parse_trace(words, len(seqs))


def ref_poa_msa(seqs, p=None):
    """abpoa_msa of the UNMODIFIED reference. seqs: list of uint8 arrays (0..4). -> uint8 [n_seq, msa_len]."""
    return _call_msa(_load(REF_SO), "ref_poa_msa", "ref_free", seqs, p or cactus_params())


def ref_poa_msa_trace(seqs, p=None):
    return _call_trace(_load(REF_SO), "ref_poa_msa_trace", "ref_free", seqs, p or cactus_params())


def oracle_poa_msa(seqs, p=None):
    return _call_msa(_load(build_oracle()), "oracle_poa_msa", "oracle_free", seqs, p or cactus_params())


def oracle_poa_msa_trace(seqs, p=None):
    return _call_trace(_load(build_oracle()), "oracle_poa_msa_trace", "oracle_free", seqs, p or cactus_params())


# ---------------------------------------------------------------------------------------------------
# BAR level: msa_make_partial_order_alignment / make_consistent_partial_order_alignments
# ---------------------------------------------------------------------------------------------------
class OracleMsa(C.Structure):
    _fields_ = [("seq_no", C.c_int64), ("column_no", C.c_int64), ("seq_lens", C.POINTER(C.c_int)),
                ("msa", C.POINTER(C.c_uint8))]


def _cstrings(strs):
    arr = (C.c_char_p * len(strs))(*[s if isinstance(s, bytes) else s.encode() for s in strs])
    lens = (C.c_int * len(strs))(*[len(s) for s in strs])
    return arr, lens


def _msa_make(libname, strs, window_size, max_prog_rows, max_prog_length_diff, p):
    p = p or cactus_params()
    arr, lens = _cstrings(strs)
    n = len(strs)
    if libname in ("ref", "shim"):
        lib = _load(BAR_REF_SO if libname == "ref" else BAR_SHIM_SO)
        f = lib.bar_ref_msa_make_partial_order_alignment
        f.restype = C.c_int64
        f.argtypes = [C.POINTER(RefParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                      C.POINTER(C.c_void_p)]
        out = C.c_void_p()
        cols = f(C.byref(p), arr, lens, n, window_size, max_prog_rows, max_prog_length_diff, C.byref(out))
        msa = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_uint8)), shape=(n * max(cols, 1),))[: n * cols]
        msa = msa.reshape(n, cols).copy()
        lib.bar_ref_free.argtypes = [C.c_void_p]
        lib.bar_ref_free(out)
        return msa
    lib = _load(build_oracle())
    f = lib.oracle_msa_make_partial_order_alignment
    f.restype = C.POINTER(OracleMsa)
    f.argtypes = [C.POINTER(RefParams), C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double]
    m = f(C.byref(p), arr, lens, n, window_size, max_prog_rows, max_prog_length_diff)
    cols = m.contents.column_no
    msa = np.ctypeslib.as_array(m.contents.msa, shape=(n * max(cols, 1),))[: n * cols].reshape(n, cols).copy()
    lib.oracle_msa_destruct.argtypes = [C.c_void_p]
    lib.oracle_msa_destruct(m)
    return msa


def ref_msa_make_partial_order_alignment(strs, window_size=10000, max_prog_rows=5000, max_prog_length_diff=1.0, p=None):
    return _msa_make("ref", strs, window_size, max_prog_rows, max_prog_length_diff, p)


def shim_msa_make_partial_order_alignment(strs, window_size=10000, max_prog_rows=5000, max_prog_length_diff=1.0, p=None):
    """the reference's msa_make_partial_order_alignment SYMBOL as re-exported by shim/cactus_bar_shim.c (GPU needed)"""
    return _msa_make("shim", strs, window_size, max_prog_rows, max_prog_length_diff, p)


def oracle_msa_make_partial_order_alignment(strs, window_size=10000, max_prog_rows=5000, max_prog_length_diff=1.0, p=None):
    return _msa_make("oracle", strs, window_size, max_prog_rows, max_prog_length_diff, p)


def _consistent(libname, ends, right_end_indexes, right_end_row_indexes, overlaps, window_size, max_prog_rows,
                max_prog_length_diff, p):
    """ends: list (per end) of list of ASCII strings. right_*/overlaps: list of int lists, same shape."""
    p = p or cactus_params()
    end_no = len(ends)
    keep = []
    end_lengths = (C.c_int64 * end_no)(*[len(e) for e in ends])
    es = (C.c_void_p * end_no)()
    el = (C.c_void_p * end_no)()
    ri = (C.c_void_p * end_no)()
    rr = (C.c_void_p * end_no)()
    ov = (C.c_void_p * end_no)()
    for i, e in enumerate(ends):
        arr, lens = _cstrings(e)
        a = (C.c_int64 * len(e))(*right_end_indexes[i])
        b = (C.c_int64 * len(e))(*right_end_row_indexes[i])
        c = (C.c_int64 * len(e))(*overlaps[i])
        keep += [arr, lens, a, b, c]
        es[i], el[i] = C.cast(arr, C.c_void_p), C.cast(lens, C.c_void_p)
        ri[i], rr[i], ov[i] = C.cast(a, C.c_void_p), C.cast(b, C.c_void_p), C.cast(c, C.c_void_p)
    out = []
    if libname in ("ref", "shim"):
        lib = _load(BAR_REF_SO if libname == "ref" else BAR_SHIM_SO)
        f = lib.bar_ref_make_consistent_partial_order_alignments
        f.restype = None
        f.argtypes = [C.POINTER(RefParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                      C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
        cols = (C.c_int64 * end_no)()
        outs = (C.c_void_p * end_no)()
        f(C.byref(p), end_no, end_lengths, es, el, ri, rr, ov, window_size, max_prog_rows, max_prog_length_diff, cols, outs)
        lib.bar_ref_free.argtypes = [C.c_void_p]
        for i in range(end_no):
            n, c = len(ends[i]), cols[i]
            m = np.ctypeslib.as_array(C.cast(outs[i], C.POINTER(C.c_uint8)), shape=(n * max(c, 1),))[: n * c]
            out.append(m.reshape(n, c).copy())
            lib.bar_ref_free(outs[i])
        return out
    lib = _load(build_oracle())
    f = lib.oracle_make_consistent_partial_order_alignments
    f.restype = C.POINTER(C.POINTER(OracleMsa))
    f.argtypes = [C.POINTER(RefParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_int64, C.c_int64, C.c_double]
    ms = f(C.byref(p), end_no, end_lengths, es, el, ri, rr, ov, window_size, max_prog_rows, max_prog_length_diff)
    lib.oracle_msa_destruct.argtypes = [C.c_void_p]
    for i in range(end_no):
        n, c = len(ends[i]), ms[i].contents.column_no
        m = np.ctypeslib.as_array(ms[i].contents.msa, shape=(n * max(c, 1),))[: n * c]
        out.append(m.reshape(n, c).copy())
        lib.oracle_msa_destruct(ms[i])
    lib.oracle_free.argtypes = [C.c_void_p]
    lib.oracle_free(ms)
    return out


def ref_make_consistent_partial_order_alignments(ends, ri, rr, ov, window_size=10000, max_prog_rows=5000,
                                                 max_prog_length_diff=1.0, p=None):
    return _consistent("ref", ends, ri, rr, ov, window_size, max_prog_rows, max_prog_length_diff, p)


def shim_make_consistent_partial_order_alignments(ends, ri, rr, ov, window_size=10000, max_prog_rows=5000,
                                                  max_prog_length_diff=1.0, p=None):
    """the reference's make_consistent_partial_order_alignments SYMBOL as re-exported by the shim (GPU needed)"""
    return _consistent("shim", ends, ri, rr, ov, window_size, max_prog_rows, max_prog_length_diff, p)


def oracle_make_consistent_partial_order_alignments(ends, ri, rr, ov, window_size=10000, max_prog_rows=5000,
                                                    max_prog_length_diff=1.0, p=None):
    return _consistent("oracle", ends, ri, rr, ov, window_size, max_prog_rows, max_prog_length_diff, p)


# ---------------------------------------------------------------------------------------------------
# host build of the product's __host__ __device__ graph code (tests/hosttest), CPU only
# ---------------------------------------------------------------------------------------------------
HOSTTEST_SO = os.environ.get("HOSTTEST_SO") or os.path.join(ROOT, "tests", "hosttest", "_build", "libhosttest.so")


def build_hosttest():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "hosttest")])
    return HOSTTEST_SO


def hosttest_poa_msa_trace(seqs, p=None):
    p = p or cactus_params()
    lib = _load(build_hosttest())
    lens, flat = _flat(seqs)
    nw = C.c_int64()
    status = C.c_int()
    f = lib.hosttest_poa_msa_trace
    f.restype = C.c_void_p
    f.argtypes = [C.POINTER(RefParams), C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    ptr = f(C.byref(p), len(seqs), lens.ctypes.data, flat.ctypes.data, C.byref(nw), C.byref(status))
    words = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(nw.value,)).copy()
    lib.hosttest_free.argtypes = [C.c_void_p]
    lib.hosttest_free(ptr)
    assert status.value == 0, "hosttest job status %d" % status.value
    return _parse_trace(words, len(seqs))


def hosttest_flowers(flowers, p=None, n_lanes=2, n_threads=3, max_jobs=6144, fail_every=0, window_size=10000, max_prog_rows=5000,
                     max_prog_length_diff=1.0, consistent=True):
    """The product's end queue + window logic (end_queue.h, bar_windows.h) on the CPU with a stand-in device (hosttest.cpp).
    flowers: list of (ends, right_end_indexes, right_end_row_indexes, overlaps) with ends = list of lists of ASCII strings.
    Returns (per flower: list of uint8 [K, cols] or None if the ticket failed, per-flower rc, batches)."""
    lib = _load(build_hosttest())
    p = p or cactus_params()
    hp = p            # RefParams has HtParams' layout (tests/hosttest/hosttest.cpp)
    end_no = np.array([len(f[0]) for f in flowers], np.int64)
    end_lengths = np.array([len(e) for f in flowers for e in f[0]], np.int64)
    strs = [s if isinstance(s, (bytes, bytearray)) else s.encode() for f in flowers for e in f[0] for s in e]
    string_lens = np.array([len(s) for s in strs] or [0], np.int64)
    blob = b"".join(strs) + b"\0"
    flat = lambda k: np.array([int(v) for f in flowers for r in f[k] for v in r] or [0], np.int64)   # noqa: E731
    rei, reri, ov = (flat(1), flat(2), flat(3)) if consistent else (None, None, None)
    n_ends = int(end_no.sum())
    cols = np.zeros(max(n_ends, 1), np.int64)
    outs = (C.c_void_p * max(n_ends, 1))()
    batches = C.c_int64()
    rcs = np.zeros(max(len(flowers), 1), np.int32)
    lib.hosttest_flowers.restype = C.c_int
    lib.hosttest_free.argtypes = [C.c_void_p]
    lib.hosttest_free.restype = None
    lib.hosttest_flowers.argtypes = [C.POINTER(RefParams), C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
    rc = lib.hosttest_flowers(C.byref(hp), n_lanes, n_threads, max_jobs, fail_every, len(flowers), end_no.ctypes.data, end_lengths.ctypes.data, blob,
                              string_lens.ctypes.data, rei.ctypes.data if consistent else None, reri.ctypes.data if consistent else None,
                              ov.ctypes.data if consistent else None, window_size, max_prog_rows, max_prog_length_diff, cols.ctypes.data, outs,
                              C.byref(batches), rcs.ctypes.data)
    if rc != 0:
        raise RuntimeError("hosttest_flowers failed: %d" % rc)
    res, e = [], 0
    for fi, f in enumerate(flowers):
        ms = []
        for end in f[0]:
            if cols[e] < 0:
                ms = None
            elif ms is not None:
                k, c = len(end), int(cols[e])
                a = np.ctypeslib.as_array(C.cast(outs[e], C.POINTER(C.c_uint8)), shape=(max(k * c, 1),))[: k * c].reshape(k, c).copy()
                ms.append(a)
            if outs[e]:
                lib.hosttest_free(outs[e])
            e += 1
        res.append(ms)
    return res, rcs[: len(flowers)].copy(), batches.value


def msa_hash(msa):
    """FNV-1a over (msa_len, bytes) of one K x msa_len matrix -- the per-end hash of bench.py's parity gate
    (same function as oracle/ref_harness.c:msa_hash)."""
    m = np.ascontiguousarray(msa, np.uint8)
    h = ((1469598103934665603 ^ (m.shape[1] & 0xffffffff)) * 1099511628211) & 0xffffffffffffffff
    for b in m.reshape(-1).tolist():
        h = ((h ^ b) * 1099511628211) & 0xffffffffffffffff
    return h


def cpu_poa_msa_many(n_seq, lens, flat, threads=0, p=None, prefer_ref=True, malloc_mode=None, want_hashes=False):
    """Time n independent abpoa_msa calls on the host cores (OpenMP over jobs). Returns (seconds, kind, checksum) or, with
    want_hashes, (seconds, kind, checksum, hashes[uint64 n]). malloc_mode (reference library only): 0 = glibc defaults,
    1 = large blocks retained and reused (the jemalloc stand-in, see oracle/ref_harness.c:ref_malloc_mode)."""
    p = p or cactus_params()
    n_seq = np.ascontiguousarray(n_seq, np.int32)
    lens = np.ascontiguousarray(lens, np.int32)
    flat = np.ascontiguousarray(flat, np.uint8)
    if prefer_ref and have_ref():
        lib, name, kind = _load(REF_SO), "ref_poa_msa_many", "reference"
        if malloc_mode is not None:
            lib.ref_malloc_mode.argtypes = [C.c_int]
            lib.ref_malloc_mode.restype = None
            lib.ref_malloc_mode(int(malloc_mode))
    else:
        lib, name, kind = _load(build_oracle()), "oracle_poa_msa_many", "port"
    f = getattr(lib, name)
    f.restype = C.c_double
    f.argtypes = [C.POINTER(RefParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
    ck = C.c_uint64()
    hashes = np.zeros(len(n_seq), np.uint64)
    secs = f(C.byref(p), len(n_seq), n_seq.ctypes.data, lens.ctypes.data, flat.ctypes.data, threads, None, C.byref(ck),
             hashes.ctypes.data if want_hashes else None)
    return (secs, kind, ck.value, hashes) if want_hashes else (secs, kind, ck.value)


# ---------------------------------------------------------------------------------------------------
# cPecan mode checkers: the compiled reference (oracle/_ref/libpecan_ref.so, oracle/pecan_ref_harness.c), the plain-C
# oracle (oracle/pecan_oracle.c) and the host emulation of the product's block program (tests/hosttest)
# ---------------------------------------------------------------------------------------------------
PECAN_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpecan_ref.so")


class PecanParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("minDiagsBetweenTraceBack", C.c_int64), ("traceBackDiagonals", C.c_int64),
                ("diagonalExpansion", C.c_int64)]


def pecan_params(threshold=0.01, min_diags=1000, tb_diags=40, expansion=20):
    return PecanParams(threshold, min_diags, tb_diags, expansion)


def have_pecan_ref():
    return os.path.exists(PECAN_REF_SO)


def _anch(anchors):
    a = np.ascontiguousarray(np.asarray(anchors, dtype=np.int64).reshape(-1, 2))
    return a, len(a)


def _take(lib_free, ptr, n, ctype, dtype, width=1):
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(max(n * width, 1),))[: n * width].astype(dtype).copy()
    lib_free(ptr)
    return arr.reshape(n, width) if width > 1 else arr


def _posteriors(lib, fname, freename, sx, sy, anchors, rl, rr, p):
    f = getattr(lib, fname)
    f.restype = C.c_int64
    f.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(PecanParams),
                  C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    fr = getattr(lib, freename)
    fr.argtypes = [C.c_void_p]
    fr.restype = None
    a, na = _anch(anchors)
    xs, ys, ps = C.c_void_p(), C.c_void_p(), C.c_void_p()
    n = f(sx, len(sx), sy, len(sy), a.ctypes.data, na, int(rl), int(rr), C.byref(p), C.byref(xs), C.byref(ys), C.byref(ps))
    if n == 0:
        for q in (xs, ys, ps):
            if q.value:
                fr(q)
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float64)
    return (_take(fr, xs, n, C.c_int64, np.int64), _take(fr, ys, n, C.c_int64, np.int64), _take(fr, ps, n, C.c_double, np.float64))


def ref_pecan_posteriors(sx, sy, anchors=(), ragged_left=False, ragged_right=False, p=None):
    """pre-floor match posteriors of ONE sub-matrix from the unmodified reference getPosteriorProbsWithBanding"""
    return _posteriors(_load(PECAN_REF_SO), "pecan_ref_posteriors", "pecan_ref_free", sx, sy, anchors, ragged_left, ragged_right, p or pecan_params())


def oracle_pecan_posteriors(sx, sy, anchors=(), ragged_left=False, ragged_right=False, p=None):
    return _posteriors(_load(build_oracle()), "oracle_pecan_posteriors", "oracle_pecan_free", sx, sy, anchors, ragged_left, ragged_right, p or pecan_params())


def ref_pecan_aligned_pairs(sx, sy, anchors=(), ragged_left=False, ragged_right=False, p=None, split_bigger=3000 * 3000):
    """integer triples (score, x, y) from the reference's public getAlignedPairsUsingAnchors"""
    lib = _load(PECAN_REF_SO)
    f = lib.pecan_ref_aligned_pairs2
    f.restype = C.c_int64
    f.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(PecanParams), C.c_int64,
                  C.POINTER(C.c_void_p)]
    lib.pecan_ref_free.argtypes = [C.c_void_p]
    lib.pecan_ref_free.restype = None
    a, na = _anch(anchors)
    t = C.c_void_p()
    p = p or pecan_params()
    n = f(sx, len(sx), sy, len(sy), a.ctypes.data, na, int(ragged_left), int(ragged_right), C.byref(p), split_bigger, C.byref(t))
    return _take(lib.pecan_ref_free, t, n, C.c_int64, np.int64, 3) if n else (lib.pecan_ref_free(t), np.zeros((0, 3), np.int64))[1]


def _aligned_pairs(lib, fname, freename, sx, sy, anchors, rl, rr, p, split_bigger, with_cells):
    f = getattr(lib, fname)
    f.restype = C.c_int64
    args = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(PecanParams), C.c_int64,
            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    if with_cells:
        args.append(C.POINTER(C.c_int64))
    f.argtypes = args
    fr = getattr(lib, freename)
    fr.argtypes = [C.c_void_p]
    fr.restype = None
    a, na = _anch(anchors)
    t, po, cells = C.c_void_p(), C.c_void_p(), C.c_int64()
    extra = [C.byref(cells)] if with_cells else []
    n = f(sx, len(sx), sy, len(sy), a.ctypes.data, na, int(rl), int(rr), C.byref(p), split_bigger, C.byref(t), C.byref(po), *extra)
    assert n >= 0, "%s failed: %d" % (fname, n)
    if n == 0:
        fr(t), fr(po)
        return np.zeros((0, 3), np.int64), np.zeros(0, np.float64), cells.value
    return _take(fr, t, n, C.c_int64, np.int64, 3), _take(fr, po, n, C.c_double, np.float64), cells.value


def oracle_pecan_aligned_pairs(sx, sy, anchors=(), ragged_left=False, ragged_right=False, p=None, split_bigger=3000 * 3000):
    """(triples, pre-floor posteriors) of the plain-C restatement of getAlignedPairsUsingAnchors"""
    t, po, _ = _aligned_pairs(_load(build_oracle()), "oracle_pecan_aligned_pairs", "oracle_pecan_free", sx, sy, anchors, ragged_left,
                              ragged_right, p or pecan_params(), split_bigger, False)
    return t, po


def hosttest_pecan_aligned_pairs(sx, sy, anchors=(), ragged_left=False, ragged_right=False, p=None, split_bigger=3000 * 3000, threads=32,
                                 ring_width=0, ring_extra=0):
    """the product's block program emulated on the host with `threads` threads per block; the diagonal ring has
    (widest diagonal + ring_extra) positions of which the first ring_width (0 = all) are the "shared memory" part and the
    rest the overflow block (tests/hosttest): (triples, posteriors, banded cells)"""
    _load(build_hosttest()).hosttest_pecan_set_threads(int(threads))
    _load(build_hosttest()).hosttest_pecan_set_ring_width(int(ring_width))
    _load(build_hosttest()).hosttest_pecan_set_ring_extra(int(ring_extra))
    return _aligned_pairs(_load(build_hosttest()), "hosttest_pecan_aligned_pairs", "hosttest_free", sx, sy, anchors, ragged_left,
                          ragged_right, p or pecan_params(), split_bigger, True)


def oracle_pecan_band(lx, ly, anchors, expansion):
    lib = _load(build_oracle())
    f = lib.oracle_pecan_band
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    a, na = _anch(anchors)
    L, R = np.zeros(lx + ly + 1, np.int64), np.zeros(lx + ly + 1, np.int64)
    f(a.ctypes.data, na, lx, ly, expansion, L.ctypes.data, R.ctypes.data)
    return L, R


def oracle_pecan_split_points(lx, ly, anchors, split_bigger, ragged_left, ragged_right):
    lib = _load(build_oracle())
    f = lib.oracle_pecan_split_points
    f.restype = C.c_int64
    f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.oracle_pecan_free.argtypes = [C.c_void_p]
    a, na = _anch(anchors)
    o = C.c_void_p()
    n = f(a.ctypes.data, na, lx, ly, split_bigger, int(ragged_left), int(ragged_right), C.byref(o))
    return _take(lib.oracle_pecan_free, o, n, C.c_int64, np.int64, 4) if n else np.zeros((0, 4), np.int64)


def cpu_pecan_many(pairs, threads=0, p=None, prefer_ref=True):
    """Time getAlignedPairsUsingAnchors over many pairs on the host cores. pairs: list of (sx, sy, anchors, rl, rr).
    Returns (seconds, kind). Python-level loop over a thread pool of ctypes calls (the libraries release the GIL)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    p = p or pecan_params()
    use_ref = prefer_ref and have_pecan_ref()
    fn = (lambda q: ref_pecan_aligned_pairs(q[0], q[1], q[2], q[3], q[4], p)) if use_ref else \
        (lambda q: oracle_pecan_aligned_pairs(q[0], q[1], q[2], q[3], q[4], p))
    fn(pairs[0])
    t0 = time.perf_counter()
    if threads and threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(fn, pairs))
    else:
        for q in pairs:
            fn(q)
    return time.perf_counter() - t0, ("reference" if use_ref else "port")


PECAN_SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libpecan_shim.so")


def have_pecan_shim():
    return os.path.exists(PECAN_SHIM_SO)


def make_all_pairwise(seqs, left_end, right_end, libname="ref"):
    """makeAllPairwiseAlignments (multipleAligner.c:667-680) of the unmodified reference ("ref") or of the shim library, where
    the reference symbol is served by libbarb200 (GPU needed). Returns (tuples [n, 5], scores [npairs, 3])."""
    lib = _load(PECAN_REF_SO if libname == "ref" else PECAN_SHIM_SO)
    f = lib.pecan_ref_make_all_pairwise
    f.restype = C.c_int64
    f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.pecan_ref_free.argtypes = [C.c_void_p]
    lib.pecan_ref_free.restype = None
    n = len(seqs)
    arr = (C.c_char_p * n)(*seqs)
    le, re_ = np.asarray(left_end, np.int64), np.asarray(right_end, np.int64)
    t, s, ns = C.c_void_p(), C.c_void_p(), C.c_int64()
    k = f(n, arr, le.ctypes.data, re_.ctypes.data, C.byref(t), C.byref(s), C.byref(ns))
    tup = np.ctypeslib.as_array(C.cast(t, C.POINTER(C.c_int64)), shape=(max(5 * k, 1),))[: 5 * k].reshape(k, 5).copy()
    sc = np.ctypeslib.as_array(C.cast(s, C.POINTER(C.c_int64)), shape=(max(3 * ns.value, 1),))[: 3 * ns.value].reshape(ns.value, 3).copy()
    lib.pecan_ref_free(t)
    lib.pecan_ref_free(s)
    return tup, sc
