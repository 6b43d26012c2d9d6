"""Flower-level checkers (test infrastructure): seeded random flowers and a ctypes driver for oracle/flower_harness.c, which
runs the reference's make_flower_alignment_poa / stPinchIterator_constructFromAlignedBlocks / bar() either from the
UNMODIFIED reference objects (oracle/_ref/libflower_ref.so) or from the drop-in build with the shims linked in
(oracle/_ref/libflower_shim.so). Both libraries are built by oracle/Makefile from /root/reference."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLOWER_REF_SO = os.path.join(ROOT, "oracle", "_ref", "libflower_ref.so")
FLOWER_SHIM_SO = os.path.join(ROOT, "oracle", "_ref", "libflower_shim.so")
FLOWER_STANDIN_SO = os.path.join(ROOT, "oracle", "_ref", "libflower_standin.so")   # shims + real host code over a CPU stand-in device
_PATHS = {"ref": FLOWER_REF_SO, "shim": FLOWER_SHIM_SO, "standin": FLOWER_STANDIN_SO,
          "harvest": os.path.join(ROOT, "oracle", "_ref", "libflower_harvest.so")}    # reference + input recorder (shim/cactus_bar_harvest.c)

# the <bar> element of src/cactus/cactus_progressive_config.xml:246-325 (values only; the keys are the XML path)
CACTUS_BAR_CONFIG = {
    "bar/runBar": "1", "bar/bandingLimit": "1000000", "bar/partialOrderAlignment": "1", "bar/minimumBlockDegree": "2",
    "bar/minimumIngroupDegree": "1", "bar/minimumOutgroupDegree": "0", "bar/minimumNumberOfSpecies": "1",
    "bar/pecan/spanningTrees": "5", "bar/pecan/gapGamma": "0.0", "bar/pecan/matchGamma": "0.2", "bar/pecan/useBanding": "1",
    "bar/pecan/splitMatrixBiggerThanThis": "3000", "bar/pecan/anchorMatrixBiggerThanThis": "500",
    "bar/pecan/repeatMaskMatrixBiggerThanThis": "500", "bar/pecan/diagonalExpansion": "20", "bar/pecan/constraintDiagonalTrim": "14",
    "bar/pecan/alignAmbiguityCharacters": "1", "bar/pecan/useProgressiveMerging": "1", "bar/pecan/pruneOutStubAlignments": "1",
    "bar/pecan/useMumAnchors": "1", "bar/pecan/recursiveMums": "1",
    "bar/poa/partialOrderAlignmentWindow": "10000", "bar/poa/partialOrderAlignmentMaskFilter": "-1",
    "bar/poa/partialOrderAlignmentBandConstant": "1000", "bar/poa/partialOrderAlignmentBandFraction": "0.1",
    "bar/poa/partialOrderAlignmentSubMatrix": "91 -114 -61 -123 -100 -114 100 -125 -61 -100 -61 -125 100 -114 -100 -123 -61 -114 91 -100 -100 -100 -100 -100 100",
    "bar/poa/partialOrderAlignmentGapOpenPenalty1": "400", "bar/poa/partialOrderAlignmentGapExtensionPenalty1": "30",
    "bar/poa/partialOrderAlignmentGapOpenPenalty2": "1200", "bar/poa/partialOrderAlignmentGapExtensionPenalty2": "1",
    "bar/poa/partialOrderAlignmentDisableSeeding": "1", "bar/poa/partialOrderAlignmentMinimizerK": "15",
    "bar/poa/partialOrderAlignmentMinimizerW": "5", "bar/poa/partialOrderAlignmentMinimizerMinW": "500",
    "bar/poa/partialOrderAlignmentProgressiveMode": "1", "bar/poa/partialOrderAlignmentProgressiveMaxRows": "5000",
    "bar/poa/partialOrderAlignmentProgressiveMaxLengthDiff": "1.0",
}

_LIBS = {}


def have(which):
    return os.path.exists(_PATHS[which])


def _lib(which):
    if which not in _LIBS:
        lib = C.CDLL(_PATHS[which])
        lib.flower_harness_set_param.argtypes = [C.c_char_p, C.c_char_p]
        lib.flower_harness_set_param.restype = None
        lib.flower_harness_clear_params.restype = None
        lib.flower_harness_begin.argtypes = [C.c_int64]
        lib.flower_harness_begin.restype = C.c_void_p
        lib.flower_harness_add_flower.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
        lib.flower_harness_add_flower.restype = C.c_int64
        lib.flower_harness_blocks.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        lib.flower_harness_blocks.restype = C.c_void_p
        lib.flower_harness_bar.argtypes = [C.c_void_p, C.c_int]
        lib.flower_harness_bar.restype = C.c_double
        lib.flower_harness_dump.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        lib.flower_harness_dump.restype = C.c_void_p
        lib.flower_harness_end.argtypes = [C.c_void_p]
        lib.flower_harness_end.restype = None
        lib.flower_harness_free.argtypes = [C.c_void_p]
        lib.flower_harness_free.restype = None
        _LIBS[which] = lib
    return _LIBS[which]


_RC = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")


def revcomp(s):
    return s.translate(_RC)[::-1]


def flowers_shared():
    """The flower of the reference's own tests (bar/tests/flowersShared.h:70-134): four sequences, three ends, six adjacencies
    incl. a negative-strand thread, a self loop and a zero-length adjacency."""
    seqs = [b"ACTGACTGAC", b"AACCGGAA", b"CGGG", b"C"]
    ends = [0, 1, 1]
    # (seq, lo, hi, positive representation, end A (side-0 view), end B (side-1 view))
    adj = [(0, 0, 5, 1, 0, 1), (0, 5, 11, 1, 0, 1), (1, 7, 9, 0, 0, 2), (1, 0, 7, 0, 2, 0), (2, 0, 5, 1, 0, 0), (3, 1, 2, 1, 1, 2)]
    return {"n_events": 1, "seqs": seqs, "seq_event": [0, 0, 0, 0], "end_side": ends, "adj": adj}


def random_flower(seed, n_threads=6, n_blocks=4, seg_len=60, n_events=3, sub=0.05, indel=0.02, p_neg=0.3, p_skip=0.1, p_loop=0.04,
                  p_empty=0.05, lower=0.0, alphabet=b"ACGT"):
    """A seeded random flower in flower_harness.c's flat form. An ancestor of n_blocks + 1 segments separated by single
    "block" bases; block b has a left end (side 1) and a right end (side 0); thread t is a mutated copy of a contiguous
    range of segments (possibly stored as its reverse complement, i.e. in the negative-strand representation), walking
    right end of block i -> left end of block i+1 (sometimes skipping a block, sometimes looping back to the end it left).
    Ends therefore see several homologous adjacency strings, which is what BAR aligns."""
    rng = np.random.default_rng(seed)

    def rand(n):
        return bytes(rng.choice(list(alphabet), size=n).astype(np.uint8)) if n > 0 else b""

    def mutate(s):
        out = bytearray()
        for ch in s:
            u = rng.random()
            if u < indel:
                continue
            out.append(int(rng.choice(list(b"ACGT"))) if u < indel + sub else ch)
            if rng.random() < indel:
                out.append(int(rng.choice(list(b"ACGT"))))
        if lower > 0:
            for i in range(len(out)):
                if rng.random() < lower:
                    out[i] = ord(chr(out[i]).lower())
        return bytes(out)

    anc = [rand(0 if rng.random() < p_empty else int(rng.integers(max(1, seg_len // 2), seg_len * 3 // 2 + 1))) for _ in range(n_blocks + 1)]
    # ends: 0 = left stub (side 0); block b (1..n_blocks): left end 2b-1 (side 1), right end 2b (side 0); last = right stub (side 1)
    end_side = [0] + [1, 0] * n_blocks + [1]
    right_end_of = lambda b: 0 if b == 0 else 2 * b             # noqa: E731  end a thread leaves block b through
    left_end_of = lambda b: 2 * b - 1 if b <= n_blocks else 2 * n_blocks + 1   # noqa: E731  end a thread enters block b through
    seqs, seq_event, adj = [], [], []
    for t in range(n_threads):
        a = int(rng.integers(0, n_blocks)) if rng.random() < 0.3 else 0
        pieces, walk, b = [], [], a                             # walk: (from block, to block or -1 for a self loop, piece index)
        while b <= n_blocks:
            u = rng.random()
            if u < p_loop and b > 0:
                pieces.append(mutate(anc[b]) + revcomp(mutate(anc[b])))
                walk.append((b, -1))
                break
            nb = b + 2 if (u < p_loop + p_skip and b + 2 <= n_blocks + 1) else b + 1
            pieces.append(b"".join(mutate(anc[x]) for x in range(b, nb)))
            walk.append((b, nb))
            if nb > n_blocks:
                break
            b = nb
        s = bytearray()
        bounds = []
        for k, piece in enumerate(pieces):
            lo = len(s)                                          # coordinate of the base before the piece (sequence starts at 1)
            s += piece
            hi = len(s) + 1
            bounds.append((lo, hi))
            if k + 1 < len(pieces):
                s += b"ACGT"[int(rng.integers(0, 4)):][:1]       # the one-base block between two adjacencies
        s = bytes(s)
        if len(s) == 0:
            continue                                             # (a thread of zero bases cannot be a Sequence)
        neg = rng.random() < p_neg
        idx = len(seqs)
        n = len(s)
        seqs.append(revcomp(s) if neg else s)
        seq_event.append(t % n_events)
        for (lo, hi), (fb, tb) in zip(bounds, walk):
            ea = right_end_of(fb)
            eb = ea if tb < 0 else left_end_of(tb)
            if neg:
                adj.append((idx, n + 1 - hi, n + 1 - lo, 0, ea, eb))
            else:
                adj.append((idx, lo, hi, 1, ea, eb))
    # ends no thread passes through are dropped (an End without caps is not a valid input: the reference asserts seq_no > 0,
    # poaBarAligner.c:466)
    used = sorted({a[4] for a in adj} | {a[5] for a in adj})
    remap = {e: i for i, e in enumerate(used)}
    adj = [(a[0], a[1], a[2], a[3], remap[a[4]], remap[a[5]]) for a in adj]
    end_side = [end_side[e] for e in used]
    return {"n_events": n_events, "seqs": seqs, "seq_event": seq_event, "end_side": end_side, "adj": adj}


def _set_params(lib, params):
    lib.flower_harness_clear_params()
    cfg = dict(CACTUS_BAR_CONFIG)
    cfg.update(params or {})
    for k, v in cfg.items():
        lib.flower_harness_set_param(k.encode(), str(v).encode())


def _add(lib, h, flower):
    seqs = flower["seqs"]
    arr = (C.c_char_p * len(seqs))(*seqs)
    ev = np.ascontiguousarray(flower["seq_event"], np.int32)
    es = np.ascontiguousarray(flower["end_side"], np.int32)
    adj = np.ascontiguousarray(np.asarray(flower["adj"], np.int64).reshape(-1, 6))
    return lib.flower_harness_add_flower(h, len(seqs), arr, ev.ctypes.data, len(es), es.ctypes.data, len(adj), adj.ctypes.data)


def _words(lib, p, n):
    w = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int64)), shape=(max(n.value, 1),))[: n.value].copy()
    lib.flower_harness_free(p)
    return w


def blocks(which, flower, params=None):
    """make_flower_alignment_poa + stPinchIterator_constructFromAlignedBlocks on one flower ->
    {"blocks": [[(cap, position, strand, length), ...], ...], "pinches": int64 [n, 6], "raw": the whole stream}"""
    lib = _lib(which)
    _set_params(lib, params)
    h = lib.flower_harness_begin(max(flower["seq_event"]) + 1)
    i = _add(lib, h, flower)
    n = C.c_int64()
    w = _words(lib, lib.flower_harness_blocks(h, i, C.byref(n)), n)
    lib.flower_harness_end(h)
    o, out = 1, []
    for _ in range(int(w[0])):
        c = int(w[o]); o += 1
        out.append([tuple(int(x) for x in w[o + 4 * k: o + 4 * k + 4]) for k in range(c)])
        o += 4 * c
    npinch = int(w[o]); o += 1
    return {"blocks": out, "pinches": w[o: o + 6 * npinch].reshape(npinch, 6), "raw": w}


def bar(which, flowers, params=None, threads=0, want_seconds=False):
    """bar() over a list of flowers in ONE cactus disk -> list of int64 streams, the canonical block list of every flower after
    BAR (with want_seconds: (streams, wall seconds of the bar() call))"""
    lib = _lib(which)
    _set_params(lib, params)
    h = lib.flower_harness_begin(max(max(f["seq_event"]) for f in flowers) + 1)
    for f in flowers:
        _add(lib, h, f)
    secs = lib.flower_harness_bar(h, threads)
    out = []
    for i in range(len(flowers)):
        n = C.c_int64()
        out.append(_words(lib, lib.flower_harness_dump(h, i, C.byref(n)), n))
    lib.flower_harness_end(h)
    return (out, secs) if want_seconds else out
