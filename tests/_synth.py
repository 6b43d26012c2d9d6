"""Seeded synthetic end generators for the tests (numpy; small sizes). Test infrastructure."""
import numpy as np

ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)
COMP = {ord("A"): "T", ord("C"): "G", ord("G"): "C", ord("T"): "A", ord("N"): "N"}


def evolve(parent, rng, sub=0.02, ins=0.005, dele=0.005, nfrac=0.0):
    """per-base substitution / insertion / deletion events, like evolveSequence in
    submodules/cPecan/impl/randomSequences.c (used by bar/tests/poaBarTest.c:36-88)"""
    out = []
    for b in parent:
        r = rng.random()
        if r < dele:
            continue
        if r < dele + sub:
            out.append((int(b) + int(rng.integers(1, 4))) % 4 if b < 4 else int(rng.integers(0, 4)))
        else:
            out.append(int(b))
        if rng.random() < ins:
            out.append(int(rng.integers(0, 4)))
    out = np.array(out, dtype=np.uint8)
    if nfrac > 0 and len(out):
        out[rng.random(len(out)) < nfrac] = 4
    if len(out) == 0:
        out = np.array([0], dtype=np.uint8)
    return out


def family(rng, K, L, sort=True, **kw):
    parent = rng.integers(0, 4, L).astype(np.uint8)
    seqs = [evolve(parent, rng, **kw) for _ in range(K)]
    if sort:
        seqs.sort(key=lambda s: -len(s))
    return seqs


def to_ascii(codes):
    return ASCII[np.asarray(codes, dtype=np.uint8)].tobytes()


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s)).encode()


def two_end_problem(rng, K, L, **kw):
    """Two ends whose strings are reverse complements of each other with full-length overlap, as in
    bar/tests/poaBarTest.c:93-179 (test_make_consistent_partial_order_alignments_two_ends)."""
    seqs = family(rng, K, L, sort=False, **kw)
    s1 = [to_ascii(s) for s in seqs]
    perm = [int(x) for x in rng.permutation(K)]
    s2 = [None] * K
    for i, pi in enumerate(perm):
        s2[pi] = revcomp(s1[i])
    ends = [s1, s2]
    right_end = [[1] * K, [0] * K]
    inv = [0] * K
    for i, pi in enumerate(perm):
        inv[pi] = i
    right_row = [perm, inv]
    overlaps = [[len(s) for s in s1], [len(s) for s in s2]]
    return ends, right_end, right_row, overlaps


def evolve_with_map(parent, rng, sub=0.02, ins=0.005, dele=0.005, nfrac=0.0):
    """like evolve(), also returning for every output base the parent position it descends from (-1 for insertions)
    and whether it was copied unchanged"""
    out, src, same = [], [], []
    for i, b in enumerate(parent):
        r = rng.random()
        if r >= dele:
            if r < dele + sub:
                out.append((int(b) + int(rng.integers(1, 4))) % 4 if b < 4 else int(rng.integers(0, 4)))
                same.append(False)
            else:
                out.append(int(b))
                same.append(True)
            src.append(i)
        if rng.random() < ins:
            out.append(int(rng.integers(0, 4)))
            src.append(-1)
            same.append(False)
    out = np.array(out, dtype=np.uint8)
    if nfrac > 0 and len(out):
        nm = rng.random(len(out)) < nfrac
        out[nm] = 4
        same = list(np.asarray(same) & ~nm)
    return out, np.array(src, dtype=np.int64), np.array(same, dtype=bool)


def pecan_pair(rng, L, k_anchor=12, keep=1.0, **kw):
    """Two descendants of a random parent as ASCII plus anchor pairs the way cPecan's MUM anchoring would place them:
    every position of every run of >= k_anchor identical, co-linear bases of the true alignment (a fraction `keep` of
    the runs is used). Returns (sx, sy, anchors[n, 2])."""
    parent = rng.integers(0, 4, L).astype(np.uint8)
    x, sx_src, sx_same = evolve_with_map(parent, rng, **kw)
    y, sy_src, sy_same = evolve_with_map(parent, rng, **kw)
    ypos = {int(s): j for j, s in enumerate(sy_src) if s >= 0 and sy_same[j]}
    cols = [(i, ypos[int(s)]) for i, s in enumerate(sx_src) if s >= 0 and sx_same[i] and int(s) in ypos and x[i] < 4]
    anchors, run = [], []
    for c in cols + [(-10, -10)]:
        if run and c[0] == run[-1][0] + 1 and c[1] == run[-1][1] + 1:
            run.append(c)
            continue
        if len(run) >= k_anchor and rng.random() < keep:
            anchors.extend(run)
        run = [c]
    a = np.array(anchors, dtype=np.int64).reshape(-1, 2)
    return to_ascii(x), to_ascii(y), a


def gapped_family(rng, K, L, gap_lens):
    """K copies of one parent, each with a point-mutation background and a few block insertions / deletions of the given lengths:
    the traceback's insertion scan (whole insertions, F1 within its crossover distance, F2 beyond) and deletion chains"""
    parent = rng.integers(0, 4, L).astype(np.uint8)
    seqs = []
    for _ in range(K):
        s = parent.copy()
        flip = rng.random(len(s)) < 0.02
        s[flip] = (s[flip] + rng.integers(1, 4, int(flip.sum()))) % 4
        for g in rng.permutation(gap_lens)[: int(rng.integers(0, 4))]:
            at = int(rng.integers(0, max(1, len(s) - g)))
            if rng.random() < 0.5:
                s = np.concatenate([s[:at], rng.integers(0, 4, int(g)).astype(np.uint8), s[at:]])
            elif len(s) > g + 8:
                s = np.concatenate([s[:at], s[at + int(g):]])
        seqs.append(s.astype(np.uint8))
    seqs.sort(key=lambda x: -len(x))
    return seqs
