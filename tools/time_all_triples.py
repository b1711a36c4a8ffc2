"""Times trialign_score_all_triples (all C(N,3) triples of a sequence set) end to end from host
memory and compares it with the same triples sent as a materialised batch through trialign_score_batch.
usage (GPU box): python tools/time_all_triples.py [nseq] [len]  -> one JSON line"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import trialign_b200 as ta

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 186
ln = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(nseq)
seqs = rng.integers(0, 4, (nseq, ln), dtype=np.uint8)
eng = ta.TriAlign(score_bits=12)
eng.align_all_triples(seqs)  # warm-up
reps = 3
t0 = time.perf_counter()
for _ in range(reps):
    got = eng.align_all_triples(seqs)
e2e = (time.perf_counter() - t0) / reps
dev, cells = ta.last_timing()
idx = ta.triple_indices(nseq)
a, b, c = seqs[idx[:, 0]], seqs[idx[:, 1]], seqs[idx[:, 2]]
eng.align_batch(a, b, c)
t0 = time.perf_counter()
for _ in range(reps):
    want = eng.align_batch(a, b, c)
e2e_batch = (time.perf_counter() - t0) / reps
print(json.dumps({"nseq": nseq, "length": ln, "triples": int(got.size), "identical_to_materialised_batch": bool((got == want).all()),
                  "all_triples": {"e2e_ms": e2e * 1e3, "e2e_gcups": cells / e2e / 1e9, "device_ms": dev * 1e3,
                                  "h2d_bytes": int(seqs.nbytes), "d2h_bytes": int(got.nbytes)},
                  "materialised_batch": {"e2e_ms": e2e_batch * 1e3, "e2e_gcups": cells / e2e_batch / 1e9,
                                         "h2d_bytes": int(3 * a.nbytes), "d2h_bytes": int(want.nbytes),
                                         "note": "pageable numpy buffers; bench.py's e2e uses pinned memory"}}), flush=True)
