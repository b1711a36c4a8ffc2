#!/usr/bin/env python
"""Small invocations of every kernel family, for compute-sanitizer (tools/sanitize.sh).  Host pointers through the C
ABI only (no torch), each result checked against the oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import oracle_bindings as ob  # noqa: E402
import trialign_b200 as ta  # noqa: E402

which = sys.argv[1:] or ["k1", "k2s", "k2l", "k2l_checked", "literal", "traceback", "all_triples"]
rng = np.random.default_rng(1)


def batch(n, L):
    a = rng.integers(0, 4, (n, L), dtype=np.uint8)
    b, c = a.copy(), a.copy()
    m = rng.random((n, L)) < 0.2
    b[m] = rng.integers(0, 4, int(m.sum()))
    m = rng.random((n, L)) < 0.2
    c[m] = rng.integers(0, 4, int(m.sum()))
    return a, b, c


for w in which:
    if w == "k1":
        a, b, c = batch(5, 64)
        got = ta.TriAlign(12).align_batch(a, b, c)
        assert ta.last_path() == "fast_batch" and got.tolist() == ob.score_batch(a, b, c, 12).tolist()
    elif w == "k2s":
        a, b, c = batch(3, 130)
        got = ta.TriAlign(12).align_batch(a, b, c)
        assert ta.last_path() == "fast_stream" and got.tolist() == ob.score_batch(a, b, c, 12).tolist()
    elif w in ("k2l", "k2l_checked"):
        a, b, c = batch(1, 200)
        a, b, c = a[0][:150], b[0][:200], c[0][:140]
        bits = 16 if w == "k2l" else 8   # 8 bits: not provably wrap-free -> literal; 10: range-checked kernel
        if w == "k2l_checked":
            bits = 10
        s, fin = ta.TriAlign(bits).align(a, b, c, return_states=True)
        es, efin = ob.score(a, b, c, bits)
        assert (s, fin.tolist()) == (es, efin.tolist()), (w, ta.last_path())
    elif w == "literal":
        a, b, c = batch(3, 40)
        got = ta.TriAlign(7, backend=ta.BACKEND_GPU_LITERAL).align_batch(a, b, c)
        assert got.tolist() == ob.score_batch(a, b, c, 7).tolist()
    elif w == "traceback":
        a, b, c = batch(1, 48)
        s, start, moves = ta.TriAlign(12).traceback(a[0], b[0][:40], c[0][:33])
        es, estart, emoves = ob.align(a[0], b[0][:40], c[0][:33], 12)
        assert s == es and tuple(start) == tuple(estart) and moves.tolist() == emoves.tolist()
    elif w == "all_triples":
        seqs = rng.integers(0, 4, (6, 32), dtype=np.uint8)
        got = ta.TriAlign(12).align_all_triples(seqs)
        idx = ta.triple_indices(6)
        want = ob.score_batch(seqs[idx[:, 0]], seqs[idx[:, 1]], seqs[idx[:, 2]], 12)
        assert got.tolist() == want.tolist()
    print("ok", w, flush=True)
