#!/usr/bin/env python
"""Regenerates profiles/ncu_kernel_stats.json from the summaries tools/ncu_export.sh wrote (their `json:` lines), stamping
every entry with the hash of the kernel sources it was captured from (bench.csrc_hash): bench.py marks an entry `stale`
when the sources have changed since.  usage: python tools/ncu_stats.py [tag]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
MAP = {"batch64": ("k1", 1048576), "batch256": ("k2s", 4096), "long1024": ("k2l_1024", 1), "long8192": ("k2l_8192", 1)}
out = {}
for workload, (k, triples) in MAP.items():
    path = os.path.join(ROOT, "profiles", f"{k}_{tag}_ncu_summary.txt")
    if not os.path.exists(path):
        continue
    text = open(path).read()
    frag = json.loads(re.search(r"^json: (.*)$", text, re.M).group(1))
    kern = re.search(r"^# (.*\(.*)$", text, re.M)
    frag.update({"kernel": kern.group(1).strip() if kern else k, "triples_in_capture": triples,
                 "capture": f"profiles/{k}_{tag}_ncu_summary.txt", "csrc_hash": bench.csrc_hash(workload)})
    out[workload] = frag
with open(os.path.join(ROOT, "profiles", "ncu_kernel_stats.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps({k: (v["alu_pipe_pct"], v["alu_instr_per_cell_pair"], v["csrc_hash"]) for k, v in out.items()}))
