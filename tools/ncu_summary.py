"""Summary of one `ncu --set full` capture for profiles/: the key metrics (tools/ncu_keys.py), the executed opcode mix
per cell pair, and a JSON fragment for profiles/ncu_kernel_stats.json.
usage: python tools/ncu_summary.py raw.csv src.csv CELLS  (CELLS = useful cell updates of the captured launch)
A "cell pair" is one 32-bit word of work (two 16-bit problems); warp-instructions per cell pair =
executed warp-level instructions x 32 lanes / (CELLS / 2)."""
import csv
import json
import subprocess
import sys
from collections import Counter

raw, src, cells = sys.argv[1], sys.argv[2], float(sys.argv[3])
print(subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_keys.py"), raw], capture_output=True, text=True).stdout, end="")
rows = list(csv.reader(open(raw)))
d = dict(zip(rows[0], rows[2]))
num = lambda k: float(d[k].replace(",", ""))
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ops = Counter()
for r in data:
    toks = [t for t in r[col["Source"]].split() if not t.startswith("@")]
    ops[toks[0].rstrip(";")] += int(r[col["Instructions Executed"]] or 0)
pairs = cells / 2.0
total = sum(ops.values())
ALU = ("VIMNMX", "VIADDMNMX", "VIADD", "IADD3", "LOP3", "PRMT", "ISETP", "SEL", "LEA", "SHF", "IABS", "PLOP3", "FLO", "POPC", "BREV", "VABSDIFF", "MOV")
alu = sum(v for k, v in ops.items() if k.startswith(ALU) and not k.startswith("IMAD"))
print(f"\nderived: executed warp-instructions per cell pair = {total * 32 / pairs:.2f}")
print(f"derived: of which on the ALU pipe (max/min/logic/add/shift/compare opcodes) = {alu * 32 / pairs:.2f}")
for k, v in ops.most_common(16):
    print(f"  {k:30s} {v * 32 / pairs:6.2f} per cell pair")
rr = list(csv.reader(open(raw)))
unit = dict(zip(rr[0], rr[1]))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
frag = {"dram_bytes_read": num("dram__bytes_read.sum") * scale[unit["dram__bytes_read.sum"]],
        "dram_bytes_write": num("dram__bytes_write.sum") * scale[unit["dram__bytes_write.sum"]],
        "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "issue_slot_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instr_per_cell_pair": round(total * 32 / pairs, 2), "alu_instr_per_cell_pair": round(alu * 32 / pairs, 2)}
print("json:", json.dumps(frag))
