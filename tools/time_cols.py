#!/usr/bin/env python
"""Intrinsic cost of the column pipeline's ranges: runs the block-column ranges of one long triple ONE AFTER THE OTHER
on a single GPU (the boundary faces / progress words of range g-1 are complete when range g starts) and times each
launch with CUDA events.  8 x (range time) against the one-launch sweep separates what a narrow strip costs by itself
(claim order, fences, boundary stores) from pipeline effects across GPUs (lag, waiting).
    python tools/time_cols.py [--len 8192] [--ranges 1 2 4 8] [--opt name=value ...]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import trialign_b200 as ta  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--len", type=int, default=8192)
ap.add_argument("--ranges", type=int, nargs="*", default=[1, 2, 4, 8])
ap.add_argument("--opt", action="append")
args = ap.parse_args()
ta.apply_option_args(args.opt)
n = args.len
a, b, c = (torch.from_numpy(ta.synth(0x5EED0005, 0, s, n)).cuda() for s in range(3))
L = ta.library()
fb, pb = C.c_int64(0), C.c_int64(0)
L.trialign_cols_buffer_bytes(n, n, C.byref(fb), C.byref(pb))
bufs = []
for _ in range(2):
    pf, pp = C.c_void_p(), C.c_void_p()
    assert L.trialign_device_alloc(0, fb.value, C.byref(pf)) == 0
    assert L.trialign_device_alloc(0, max(pb.value, 256), C.byref(pp)) == 0
    bufs.append((pf, pp))
out = torch.zeros(8, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
epoch = 0
for S in args.ranges:
    times = []
    for rep in range(2):
        epoch += 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(S + 1)]
        ev[0].record()
        for g in range(S):
            c0, c1 = ta.col_range(n, g, S)
            src = bufs[(g - 1) & 1] if g > 0 else (None, None)
            dst = bufs[g & 1] if g + 1 < S else (None, None)
            rc = L.trialign_cols_score_dev(0, a.data_ptr(), n, b.data_ptr(), n, c.data_ptr(), n, c0, c1, src[0], src[1],
                                           dst[0], dst[1], epoch, 16, out.data_ptr(), st)
            assert rc == 0, rc
            ev[g + 1].record()
        torch.cuda.synchronize()
        times = [ev[g].elapsed_time(ev[g + 1]) for g in range(S)]
    print(json.dumps({"len": n, "ranges": S, "score": int(out[0].item()), "ms_per_range": [round(t, 2) for t in times],
                      "ms_total": round(sum(times), 2), "options": args.opt or []}), flush=True)
