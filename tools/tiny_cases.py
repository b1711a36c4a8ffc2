"""Exercises every kernel family once on tiny inputs and checks the oracle (seconds on a GPU box; used before timing
runs in tools/perf_quick.sh).  usage: python tools/tiny_cases.py [all|k1|k2b|k2l|lit]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import oracle_bindings as ob
import trialign_b200 as ta

rng = np.random.default_rng(3)


def batch(n, ln, w=16):
    a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    for t in range(0, n, 2):
        a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
    got = ta.TriAlign(w).align_batch(a, b, c)
    assert got.tolist() == ob.score_batch(a, b, c, w).tolist(), (n, ln, ta.last_path())
    print("batch", n, ln, ta.last_path(), "ok", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "k1"):
    batch(9, 64)       # K1
if which in ("all", "k2b"):
    batch(3, 130)      # K2 batch: 3x3 blocks per pair
if which in ("all", "k2l"):
    x, y, z = ob.mutated_triple(rng, 200)
    s, fin = ta.TriAlign(16).align(x[:70], y[:130], z[:200], return_states=True)   # K2 long: 3 x 2 tasks
    es, efin = ob.score(x[:70], y[:130], z[:200], 16)
    assert s == es and fin.tolist() == efin.tolist()
    print("long", ta.last_path(), "ok", flush=True)
if which in ("all", "lit"):
    x, y, z = ob.mutated_triple(rng, 40)
    assert ta.TriAlign(9, backend=ta.BACKEND_GPU_LITERAL).align(x, y, z) == ob.score(x, y, z, 9)[0]
    print("literal ok", flush=True)
