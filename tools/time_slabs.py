#!/usr/bin/env python
"""Where does the i-slab pipeline lose time?  Runs the slabs of one long triple ONE AFTER THE OTHER on a single GPU
(no waiting on a neighbour: the previous slab has finished) and times every launch with CUDA events.
    python tools/time_slabs.py [--len 8192] [--slabs 8] [--opt name=value ...]
8 x (slab time) against the one-launch sweep separates the per-task cost of short slabs (set-up, drain, state
hand-over) from pipeline effects (fill, waiting)."""
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
ap.add_argument("--slabs", type=int, nargs="*", default=[1, 2, 4, 8])
ap.add_argument("--opt", action="append")
args = ap.parse_args()
ta.apply_option_args(args.opt)
n = args.len
a, b, c = (torch.from_numpy(ta.synth(0x5EED0005, 0, s, n)).cuda() for s in range(3))
L = ta.library()
sb, fb = C.c_int64(0), C.c_int64(0)
L.trialign_slab_buffer_bytes(n, n, C.byref(sb), C.byref(fb))
bufs = []
for _ in range(2):
    ps, pf = C.c_void_p(), C.c_void_p()
    assert L.trialign_device_alloc(0, sb.value, C.byref(ps)) == 0
    assert L.trialign_device_alloc(0, fb.value, C.byref(pf)) == 0
    bufs.append((ps, pf))
out = torch.zeros(8, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
epoch = 0
for S in args.slabs:
    times = []
    for rep in range(2):
        epoch += 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(S + 1)]
        ev[0].record()
        for g in range(S):
            i0, i1 = ta.slab_range(n, g, S)
            src = bufs[(g - 1) & 1] if g > 0 else (None, None)
            dst = bufs[g & 1] if g + 1 < S else (None, None)
            if S == 1:
                ta.TriAlign(16).align_cuda(a, b, c, out=out)
            else:
                rc = L.trialign_slab_score_dev(0, a.data_ptr(), n, b.data_ptr(), n, c.data_ptr(), n, i0, i1, src[0], src[1],
                                               dst[0], dst[1], epoch, 16, out.data_ptr(), st)
                assert rc == 0, rc
            ev[g + 1].record()
        torch.cuda.synchronize()
        times = [ev[g].elapsed_time(ev[g + 1]) for g in range(S)]
    print(json.dumps({"len": n, "slabs": S, "score": int(out[0].item()), "ms_per_slab": [round(t, 2) for t in times],
                      "ms_total": round(sum(times), 2), "options": args.opt or []}), flush=True)
