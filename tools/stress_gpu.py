"""Randomised cross-check of the fast kernels against the literal kernel (two independent GPU implementations behind the
same C ABI) AND, wherever the case is small enough for it, against the CPU oracle, over many shapes: K2S batches (len
65..520, odd batch sizes), K2L single triples (ragged shapes, with and without the run-time range check), K1 batches
(len 1..64); score_bits from 6 to 16, so that the range-checked batch kernels and the flagged-triple re-run take part.
usage (GPU box): python tools/stress_gpu.py [seconds]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import oracle_bindings as ob
import trialign_b200 as ta

ORACLE_CELLS = 4e8  # cases up to this many cells are also checked against the oracle

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(2026)
t_end = time.time() + budget
counts = {"fast_batch": 0, "fast_stream": 0, "fast_long": 0, "fast_then_literal": 0, "oracle_checked": 0, "flagged": 0}


def mutated(n, ln, rate):
    a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    b, c = a.copy(), a.copy()
    for x in (b, c):
        m = rng.random((n, ln)) < rate
        x[m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
    return a, b, c


rounds = 0
while time.time() < t_end:
    kind = rounds % 3
    w = int(rng.choice([12, 16, 12, 16, 6, 7, 8, 9, 10, 11]))
    if kind == 0:  # K2S
        ln, n = int(rng.integers(65, 521)), int(rng.integers(1, 700))
        if n * ln ** 3 > 6e9:
            n = max(1, int(6e9 / ln ** 3))
        a, b, c = mutated(n, ln, float(rng.choice([0.02, 0.15, 0.75])))
        fast = ta.TriAlign(w).align_batch(a, b, c)
        path = ta.last_path()
        counts["flagged"] += ta.last_flagged()
        lit = ta.TriAlign(w, backend=ta.BACKEND_GPU_LITERAL).align_batch(a, b, c)
        assert fast.tolist() == lit.tolist(), ("batch", ln, n, w, path)
        if n * ln ** 3 <= ORACLE_CELLS:
            assert fast.tolist() == ob.score_batch(a, b, c, w).tolist(), ("batch vs oracle", ln, n, w, path)
            counts["oracle_checked"] += 1
    elif kind == 1:  # K2L
        la, lb, lc = (int(x) for x in rng.integers(1, 700, 3))
        a, b, c = mutated(1, max(la, lb, lc), float(rng.choice([0.01, 0.2, 0.75])))
        a, b, c = a[0, :la], b[0, :lb], c[0, :lc]
        fast = ta.TriAlign(w).align(a, b, c, return_states=True)
        path = ta.last_path()
        lit = ta.TriAlign(w, backend=ta.BACKEND_GPU_LITERAL).align(a, b, c, return_states=True)
        assert fast[0] == lit[0] and fast[1].tolist() == lit[1].tolist(), ("single", la, lb, lc, w, path)
        if la * lb * lc <= ORACLE_CELLS:
            es, efin = ob.score_mt(a, b, c, w)
            assert fast[0] == es and fast[1].tolist() == efin.tolist(), ("single vs oracle", la, lb, lc, w, path)
            counts["oracle_checked"] += 1
    else:  # K1
        ln, n = int(rng.integers(1, 65)), int(rng.integers(1, 3000))
        a, b, c = mutated(n, ln, float(rng.choice([0.02, 0.15, 0.75])))
        fast = ta.TriAlign(w).align_batch(a, b, c)
        path = ta.last_path()
        counts["flagged"] += ta.last_flagged()
        lit = ta.TriAlign(w, backend=ta.BACKEND_GPU_LITERAL).align_batch(a, b, c)
        assert fast.tolist() == lit.tolist(), ("short batch", ln, n, w, path)
        if n * ln ** 3 <= ORACLE_CELLS:
            assert fast.tolist() == ob.score_batch(a, b, c, w).tolist(), ("short batch vs oracle", ln, n, w, path)
            counts["oracle_checked"] += 1
    counts[path] = counts.get(path, 0) + 1
    rounds += 1
print("stress ok:", rounds, "rounds", counts, flush=True)
