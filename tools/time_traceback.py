"""Times trialign_align (KT, the traceback path) on mutated-copy triples of a few lengths and re-scores every path.
usage (GPU box): python tools/time_traceback.py [lengths...]  -> one JSON line per length"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import oracle_bindings as ob
import trialign_b200 as ta

for L in [int(x) for x in sys.argv[1:]] or [64, 256, 512, 1024]:
    rng = np.random.default_rng(L)
    a, b, c = ob.mutated_triple(rng, L, rate=0.2)
    eng = ta.TriAlign(score_bits=16)
    eng.traceback(a, b, c)  # warm-up: allocations, module load
    t0 = time.perf_counter()
    s, start, moves = eng.traceback(a, b, c)
    wall = time.perf_counter() - t0
    dev, cells = ta.last_timing()
    ok = ob.rescore_path(a, b, c, start, moves) == s and s == eng.align(a, b, c)
    print(json.dumps({"length": L, "score": s, "columns": int(moves.size), "start": start, "device_ms": dev * 1e3,
                      "wall_ms": wall * 1e3, "gcups": cells / dev / 1e9, "direction_cube_mb": 4.0 * L ** 3 / 1e6,
                      "segments": ta.last_segments(), "path_rescored_ok": bool(ok)}), flush=True)
    if L >= 256:
        # the checkpointed form on the same triple: a budget of 40 % of the cube forces segments
        ta.set_option("kt_budget_mb", max(1, int(0.4 * 4.0 * L ** 3 / 2 ** 20)))
        try:
            eng.traceback(a, b, c)
            t0 = time.perf_counter()
            s2, start2, moves2 = eng.traceback(a, b, c)
            wall = time.perf_counter() - t0
            print(json.dumps({"length": L, "form": "checkpointed", "segments": ta.last_segments(), "wall_ms": wall * 1e3,
                              "gcups_wall": cells / wall / 1e9,
                              "same_path": bool(s2 == s and start2 == start and moves2.tolist() == moves.tolist())}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"length": L, "form": "checkpointed", "error": str(e)}), flush=True)
        finally:
            ta.set_option("kt_budget_mb", None)
