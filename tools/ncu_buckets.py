"""Summarise an `ncu --page source --csv` dump: stall samples and executed warp-instructions per bucket of SASS
instructions, with the top stall reasons, plus shared-memory wavefront excess per instruction.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_buckets.py src.csv [bucket=40]"""
import csv
import sys

path = sys.argv[1]
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print(f"# {rows[0][1] if rows[0] else ''}")
print(f"# total samples {tot}; {len(data)} SASS instructions; columns: first index, samples, share, executed warp-instr (M), first instruction, top stall reasons")
for b in range(0, len(data), bucket):
    chunk = data[b:b + bucket]
    smp = sum(int(r[col["# Samples"]] or 0) for r in chunk)
    ex = sum(int(r[col["Instructions Executed"]] or 0) for r in chunk)
    agg = {s: sum(int(r[col[s]] or 0) for r in chunk) for s in stalls}
    top = sorted(agg.items(), key=lambda kv: -kv[1])[:4]
    print(f"{b:5d} {smp:8d} {100.0 * smp / max(tot, 1):5.1f}%  exec={ex / 1e6:8.1f}M  {chunk[0][col['Source']][:32]:32s} " +
          " ".join(f"{k[6:]}={v}" for k, v in top))
exc = [(int(r[col["L1 Wavefronts Shared Excessive"]] or 0), int(r[col["L1 Wavefronts Shared"]] or 0), i, r[col["Source"]]) for i, r in enumerate(data)]
exc = [e for e in exc if e[1] > 0]
print("# shared-memory instructions with excessive wavefronts (excess, total, index, SASS)")
for e in sorted(exc, reverse=True)[:12]:
    print(f"  {e[0] / 1e6:8.1f}M of {e[1] / 1e6:8.1f}M  #{e[2]:4d} {e[3][:70]}")
