"""Kernel-only time of single long triples of various block-grid shapes (how the step time and the block-to-block
pipeline fill of K2L add up).  usage (GPU box): python tools/time_long_shapes.py [name=value ...]   (library options,
e.g. k2_gate=8; SHAPES=short in the environment selects the short shape list)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import trialign_b200 as ta

ta.apply_option_args(sys.argv[1:])
TAG = " ".join(sys.argv[1:])
SHAPES = [(1024, 64, 128), (1024, 128, 128), (1024, 256, 128), (1024, 1024, 128), (1024, 64, 256), (1024, 64, 1024),
          (1024, 1024, 1024), (4096, 64, 128), (4096, 1024, 1024), (2048, 2048, 2048)]
if os.environ.get("SHAPES") == "short":
    SHAPES = [(1024, 64, 256), (1024, 64, 1024), (1024, 1024, 128), (1024, 1024, 1024), (4096, 1024, 1024)]
for w in ((16,) if os.environ.get("SHAPES") == "short" else (16, 12)):
    for la, lb, lc in SHAPES:
        a, b, c = (ta.synth(0x5EED0003, 0, s, n) for s, n in ((0, la), (1, lb), (2, lc)))
        eng = ta.TriAlign(score_bits=w)
        eng.align(a, b, c)
        best = 1e9
        for _ in range(3):
            eng.align(a, b, c)
            best = min(best, ta.last_timing()[0])
        tasks_j, tasks_k = (lb + 63) // 64, (lc + 127) // 128
        print(json.dumps({"tag": TAG, "w": w, "shape": [la, lb, lc], "blocks": [tasks_j, tasks_k], "path": ta.last_path(),
                          "ms": round(best * 1e3, 4), "us_per_plane": round(best * 1e6 / la, 4),
                          "gcups": round(la * lb * lc / best / 1e9, 1)}), flush=True)
