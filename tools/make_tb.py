"""Replay any triple on the reference RTL: regenerate TriAlign_tb.sv's stimulus for it and (optionally) simulate.

The reference testbench (src/TriAlign_tb.sv) hard-codes its three sequences in an `initial` block of 1,536
`seq?_ram[w][hi:lo] <= SYM;` assignments (:423-1959), its lengths in localparams (:48-50) and dumps an FSDB through
Verdi system tasks (:266-267) that open-source simulators do not have.  This tool takes a checkout of the reference and
three dat files (or a synthetic triple) and writes a testbench variant that
  * sets A/B/C_LENGTH (and SCORE_BITS when asked),
  * replaces the initial block by the lines libtrialign's `trialign_write_tb_init` produces,
  * comments the $fsdbDump* calls out,
then, with --run, compiles it together with the 1cyc or 2cyc RTL under Icarus Verilog or Verilator when one is on PATH,
parses the `TriAlign Score:` line the testbench prints (:341) and compares it with libtrialign's score (GPU) or, with
--expect, a given number.  No simulator exists in the build image, so the --run leg is exercised by nobody here: it
is the hook that closes the "hand-transcribed RTL model" gap on a machine that has one.

    python tools/make_tb.py --ref /path/to/reference --a dat/A_seq.dat --b dat/B_seq.dat --c dat/C_seq.dat \
        --out /tmp/TriAlign_tb_gen.sv [--variant 1cyc|2cyc] [--score-bits 12] [--run] [--expect 1]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rewrite_testbench(tb_text: str, init_lines: str, la: int, lb: int, lc: int, score_bits=None) -> str:
    """Pure text transformation (unit-tested against the reference's own testbench)."""
    lines = tb_text.split("\n")
    first = next(i for i, l in enumerate(lines) if l.startswith("seqA_ram[0][3:0]"))
    last = max(i for i, l in enumerate(lines) if re.match(r"seq[ABC]_ram\[\d+\]\[\d+:\d+\] <= ", l))
    if not all(re.match(r"seq[ABC]_ram\[\d+\]\[\d+:\d+\] <= \w+;\s*$", l) for l in lines[first:last + 1]):
        raise ValueError("unexpected line inside the testbench's initial block")
    body = init_lines.rstrip("\n").split("\n")
    out = lines[:first] + body + lines[last + 1:]
    text = "\n".join(out)
    for name, val in (("A_LENGTH", la), ("B_LENGTH", lb), ("C_LENGTH", lc)):
        text, k = re.subn(r"(localparam\s+%s\s*=\s*)\d+(\s*;)" % name, r"\g<1>%d\g<2>" % val, text)
        if k != 1:
            raise ValueError(f"localparam {name} not found exactly once")
    if score_bits is not None:
        text, k = re.subn(r"(localparam\s+SCORE_BITS\s*=\s*)\d+(\s*;)", r"\g<1>%d\g<2>" % score_bits, text)
        if k != 1:
            raise ValueError("localparam SCORE_BITS not found exactly once")
    text = re.sub(r"^(\s*)(\$fsdbDump\w+\(.*\);)", r"\1// \2", text, flags=re.M)
    return text


def find_simulator():
    for name in ("iverilog", "verilator"):
        path = shutil.which(name)
        if path:
            return name, path
    return None, None


def simulate(sim, tb_path, rtl_files, workdir):
    """Returns the integer the testbench prints after 'TriAlign Score:'."""
    if sim == "iverilog":
        exe = os.path.join(workdir, "tb.vvp")
        subprocess.run(["iverilog", "-g2012", "-o", exe, tb_path] + rtl_files, check=True)
        out = subprocess.run(["vvp", exe], check=True, capture_output=True, text=True).stdout
    else:
        subprocess.run(["verilator", "--binary", "--timing", "-Wno-fatal", "--top-module", "TriAlign_tb", "-Mdir", workdir,
                        tb_path] + rtl_files, check=True)
        out = subprocess.run([os.path.join(workdir, "VTriAlign_tb")], check=True, capture_output=True, text=True).stdout
    m = re.search(r"TriAlign Score:\s*(-?\d+)", out)
    if not m:
        raise RuntimeError("the simulation printed no score:\n" + out[-2000:])
    return int(m.group(1))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref", required=True, help="checkout of the reference repository")
    ap.add_argument("--a"), ap.add_argument("--b"), ap.add_argument("--c")
    ap.add_argument("--synth", type=int, default=0, help="use the synthetic triple of this length instead of dat files")
    ap.add_argument("--seed", type=lambda x: int(x, 0), default=0x5EED0001)
    ap.add_argument("--out", required=True)
    ap.add_argument("--variant", default="1cyc", choices=["1cyc", "2cyc"])
    ap.add_argument("--score-bits", type=int, default=None)
    ap.add_argument("--total-len", type=int, default=512)
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--expect", type=int, default=None)
    args = ap.parse_args()

    import trialign_b200 as ta

    if args.synth:
        a, b, c = (ta.synth(args.seed, 0, w, args.synth) for w in range(3))
    else:
        a, b, c = (ta.read_seq_dat(p) for p in (args.a, args.b, args.c))
    with tempfile.TemporaryDirectory() as tmp:
        init = os.path.join(tmp, "init.sv")
        ta.write_tb_init(init, a, b, c, args.total_len)
        with open(os.path.join(args.ref, "src", "TriAlign_tb.sv")) as f:
            tb = f.read()
        with open(init) as f:
            text = rewrite_testbench(tb, f.read(), len(a), len(b), len(c), args.score_bits)
        with open(args.out, "w") as f:
            f.write(text)
        print(f"wrote {args.out}: lengths {len(a)},{len(b)},{len(c)}")
        if not args.run:
            return 0
        sim, _ = find_simulator()
        if not sim:
            print("no Verilog simulator (iverilog / verilator) on PATH: nothing simulated", file=sys.stderr)
            return 2
        rtl = [os.path.join(args.ref, "src", f"TriAlign_{args.variant}.v"), os.path.join(args.ref, "src", f"PE_{args.variant}.v")]
        got = simulate(sim, args.out, rtl, tmp)
        want = args.expect
        if want is None:
            want = ta.TriAlign(score_bits=args.score_bits or 12).align(a, b, c)
        print(f"RTL ({sim}, {args.variant}) score {got}; expected {want}")
        return 0 if got == want else 1


if __name__ == "__main__":
    sys.exit(main())
