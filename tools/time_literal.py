#!/usr/bin/env python
"""GCUPS of the exactness fallbacks: the literal plane-sweep kernel KL (batches and single triples), the range-checked
batch kernels (K1 / K2S CHECK variants on data that never trips the check), and a batch in which a fraction of the
triples is flagged and re-scored literally.  Device-resident inputs, CUDA events.  python tools/time_literal.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import trialign_b200 as ta  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def batch(n, ln, w, backend, label, hot=0):
    a, b, c = ta.synth_batch_cuda(0x5EED0002, 0, n, ln)
    if hot:  # every hot-th triple becomes identical sequences: real wraps at small W
        b[::hot] = a[::hot]
        c[::hot] = a[::hot]
    eng = ta.TriAlign(score_bits=w, backend=backend)
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    ms = timed(lambda: eng.align_batch_cuda(a, b, c, out=out))
    print(json.dumps({"case": label, "triples": n, "len": ln, "score_bits": w, "path": ta.last_path(), "flagged": ta.last_flagged(),
                      "ms": round(ms, 3), "gcups": round(n * ln ** 3 / ms / 1e6, 1)}), flush=True)


def single(ln, w, backend, label):
    a, b, c = (torch.from_numpy(ta.synth(0x5EED0003, 0, s, ln)).cuda() for s in range(3))
    eng = ta.TriAlign(score_bits=w, backend=backend)
    out = torch.empty(8, dtype=torch.int32, device="cuda")
    ms = timed(lambda: eng.align_cuda(a, b, c, out=out), reps=2)
    print(json.dumps({"case": label, "len": ln, "score_bits": w, "path": ta.last_path(), "ms": round(ms, 3),
                      "gcups": round(ln ** 3 / ms / 1e6, 1)}), flush=True)


def batch_ex(n, ln, w, sc, label):
    """Explicit scoring parameters (host-pointer entry point; device time of the kernels from trialign_last_timing)."""
    a, b, c = ta.synth_batch(0x5EED0002, 0, n, ln)
    eng = ta.TriAlign(score_bits=w, scoring=ta.Scoring(*sc))
    eng.align_batch(a, b, c)
    eng.align_batch(a, b, c)
    dev, cells = ta.last_timing()
    print(json.dumps({"case": label, "scoring": list(sc), "triples": n, "len": ln, "score_bits": w, "path": ta.last_path(),
                      "flagged": ta.last_flagged(), "ms": round(dev * 1e3, 3), "gcups": round(cells / dev / 1e9, 1)}), flush=True)


batch(1 << 16, 64, 12, ta.BACKEND_GPU_LITERAL, "KL literal batch")
batch(1 << 12, 256, 12, ta.BACKEND_GPU_LITERAL, "KL literal batch")
single(512, 12, ta.BACKEND_GPU_LITERAL, "KL literal single")
single(1024, 12, ta.BACKEND_GPU_LITERAL, "KL literal single")
batch(1 << 20, 64, 12, ta.BACKEND_AUTO, "K1 unchecked (W=12)")
batch(1 << 20, 64, 8, ta.BACKEND_AUTO, "K1 range-checked (W=8), nothing flagged")
batch(1 << 20, 64, 8, ta.BACKEND_AUTO, "K1 range-checked (W=8), 1/1024 flagged -> literal", hot=1024)
batch(1 << 14, 256, 10, ta.BACKEND_AUTO, "K2S range-checked (W=10), nothing flagged")
batch(1 << 14, 256, 12, ta.BACKEND_AUTO, "K2S unchecked (W=12)")
batch_ex(1 << 18, 64, 12, (2, -1, 2, 1, 0), "K1 run-time scoring, the testbench's MATCH=2 / MISMATCH=-1 / GO=2 / GE=1")
batch_ex(1 << 18, 64, 12, (1, -1, 2, 1, 1), "K1 run-time scoring, sum of pairs")
batch_ex(1 << 13, 256, 12, (2, -1, 2, 1, 0), "K2S run-time scoring, MATCH=2 / MISMATCH=-1 / GO=2 / GE=1")
batch_ex(1 << 14, 64, 12, (5, -4, 10, 1, 1), "outside the fast domain (GO = 10): literal kernel")
