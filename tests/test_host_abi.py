"""CPU-side tests of libtrialign's C ABI: the library loads, exports every symbol include/trialign.h declares,
file formats round-trip, argument errors come back as codes, and no compute happens without a GPU."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    with open(os.path.join(ROOT, "include", "trialign.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trialign_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ta):
    names = _declared_functions()
    assert len(names) >= 15
    lib = C.CDLL(ta.library_path())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/trialign.h but not exported"


def test_version_and_strerror(ta):
    assert ta.version() == 0x00010000
    L = ta.library()
    assert L.trialign_strerror(0) == b"ok"
    for rc in range(-9, 0):
        assert len(L.trialign_strerror(rc)) > 3
    assert L.trialign_strerror(-99) == b"unknown error"


def test_seq_dat_roundtrip_crlf_and_lf(ta, golden, tmp_path):
    A = np.array(golden["dat"]["A"], np.uint8)
    p = str(tmp_path / "A_seq.dat")
    ta.write_seq_dat(p, A, crlf=True)
    raw = open(p, "rb").read()
    assert len(raw) == 192 and raw[:3] == b"3\r\n"
    # byte-identical to the reference's dat/A_seq.dat
    assert hashlib.sha256(raw).hexdigest() == golden["dat"]["A_seq_sha256"]
    assert ta.read_seq_dat(p).tolist() == A.tolist()
    ta.write_seq_dat(p, A, crlf=False)
    assert ta.read_seq_dat(p).tolist() == A.tolist()
    # no trailing newline is fine too
    open(p, "wb").write(b"0\n1\n2\n3")
    assert ta.read_seq_dat(p).tolist() == [0, 1, 2, 3]


@pytest.mark.skipif(not os.path.exists("/root/reference/dat/A_seq.dat"), reason="reference only in the build container")
def test_reads_reference_dat_files_directly(ta, golden):
    for nm in "ABC":
        got = ta.read_seq_dat(f"/root/reference/dat/{nm}_seq.dat")
        assert got.tolist() == golden["dat"][nm]
    assert ta.read_scores("/root/reference/dat/Pattern1.dat").tolist() == golden["pattern1"]


def test_seq_dat_errors(ta, tmp_path):
    L = ta.library()
    n = C.c_int(0)
    buf = (C.c_uint8 * 8)()
    assert L.trialign_read_seq_dat(b"/nonexistent/file.dat", buf, 8, C.byref(n)) == -4  # EIO
    p = str(tmp_path / "bad.dat")
    open(p, "wb").write(b"0\n4\n")
    assert L.trialign_read_seq_dat(p.encode(), buf, 8, C.byref(n)) == -5  # EPARSE: digit outside 0-3
    open(p, "wb").write(b"01\n")
    assert L.trialign_read_seq_dat(p.encode(), buf, 8, C.byref(n)) == -5  # two symbols on a line
    open(p, "wb").write(b"0\n1\n2\n3\n0\n1\n2\n3\n0\n")
    assert L.trialign_read_seq_dat(p.encode(), buf, 8, C.byref(n)) == -8 and n.value == 9  # ECAPACITY
    assert L.trialign_read_seq_dat(None, buf, 8, C.byref(n)) == -1


def test_scores_writer_reproduces_pattern1_layout(ta, golden, tmp_path):
    vals = np.array(golden["pattern1"], np.int32)
    assert len(vals) == 572
    p = str(tmp_path / "Pattern1.dat")
    ta.write_scores(p, vals)
    raw = open(p, "rb").read()
    assert len(raw) == 1147 and not raw.endswith(b"\n") and b"\r" not in raw
    assert hashlib.sha256(raw).hexdigest() == golden["dat"]["Pattern1_sha256"]
    assert ta.read_scores(p).tolist() == vals.tolist()
    ta.write_scores(p, np.array([-5, 0, 2047, -2048], np.int32))
    assert open(p, "rb").read() == b"-5\n0\n2047\n-2048"
    ta.write_scores(p, np.zeros(0, np.int32))
    assert open(p, "rb").read() == b""


def test_synth_matches_oracle_generator(ta, ob, golden):
    for seed, t, s, n in ((0x5EED0002, 0, 0, 64), (0x5EED0004, 123456, 2, 256), (0x5EED0005, 0, 1, 1000)):
        assert ta.synth(seed, t, s, n).tolist() == ob.synth(seed, t, s, n).tolist()
    ent = golden["synth"][0]
    assert ta.synth(ent["seed"], 0, 0, 16).tolist() == ent["first_symbols"]["A"]


def test_argument_errors_are_codes_not_crashes(ta):
    L = ta.library()
    s = C.c_int32(7)
    a = (C.c_uint8 * 4)(0, 1, 2, 3)
    assert L.trialign_score(a, 4, a, 4, a, 4, 12, 0, 0, None, None) == -1
    assert L.trialign_score(a, -1, a, 4, a, 4, 12, 0, 0, C.byref(s), None) == -1
    assert L.trialign_score(a, 4, a, 4, a, 4, 1, 0, 0, C.byref(s), None) == -1    # score_bits < 2
    assert L.trialign_score(a, 4, a, 4, a, 4, 33, 0, 0, C.byref(s), None) == -1   # score_bits > 32
    assert L.trialign_score(None, 4, a, 4, a, 4, 12, 0, 0, C.byref(s), None) == -1
    assert L.trialign_score(a, 4, a, 4, a, 4, 12, 77, 0, C.byref(s), None) == -1  # unknown backend
    # the CPU checkers are deliberately not reachable through the product library
    for backend in (1, 2, 3):
        assert L.trialign_score(a, 4, a, 4, a, 4, 12, backend, 0, C.byref(s), None) == -6
    assert L.trialign_score_batch(a, a, a, -1, 4, 12, 0, 0, None) == -1


def test_no_cpu_fallback_without_a_gpu(ta):
    if ta.device_count() > 0:
        pytest.skip("a GPU is present; covered by the gpu tests")
    with pytest.raises(ta.TriAlignError) as ei:
        ta.TriAlign().align([0, 1, 2], [0, 1, 2], [0, 1, 2])
    assert ei.value.rc == -7  # ENODEVICE, not a silently computed score
    with pytest.raises(ta.TriAlignError) as ei:
        ta.TriAlign(12).traceback([0, 1, 2], [0, 1, 2], [0, 1, 2])
    assert ei.value.rc == -7
    with pytest.raises(ta.TriAlignError) as ei:
        ta.TriAlign(12).align_all_triples(np.zeros((4, 8), np.uint8))
    assert ei.value.rc == -7
    assert ta.library().trialign_triple_count(186) == 1055240 and ta.library().trialign_triple_count(2) == 0
    assert ta.triple_indices(5).tolist()[:4] == [[0, 1, 2], [0, 1, 3], [0, 1, 4], [0, 2, 3]]
    with pytest.raises(ta.TriAlignError):
        ta.TriAlign().align_batch(np.zeros((2, 4), np.uint8), np.zeros((2, 4), np.uint8), np.zeros((2, 4), np.uint8))
    assert ta.launch_count() == 0


def test_product_does_not_link_the_oracle(ta):
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", ta.library_path()], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "rtl_model" not in out
    srcs = os.path.join(ROOT, "hw-accelerator-three-sequence-alignment_b200")
    for dirpath, _, files in os.walk(srcs):
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".cpp", ".h", ".py")):
                text = open(os.path.join(dirpath, fn)).read()
                for needle in ("liboracle", "trialign_oracle.h", "rtl_model.h", "oracle_bindings", "oracle_score"):
                    assert needle not in text, f"{fn} references the test oracle ({needle})"


# ---- the testbench's packed sequence RAMs and stimulus block ----
def _tb_ram_symbol(ram_rows, n):
    """src/TriAlign_tb.sv:150-169 restated: word n>>5 of the 128-bit RAM, nibble n&31 counted from the LSB."""
    word = int.from_bytes(bytes(ram_rows[n >> 5].tolist()), "little")
    return (word >> (4 * (n & 31))) & 15


def test_seq_ram128_layout_and_roundtrip(ta):
    rng = np.random.default_rng(5)
    for n in (0, 1, 31, 32, 33, 64, 100, 512):
        s = rng.integers(0, 5, n, dtype=np.uint8)           # codes A,T,C,G,N
        ram = ta.pack_seq_ram128(s)
        assert ram.shape == ((n + 31) // 32, 16)
        assert [_tb_ram_symbol(ram, p) for p in range(n)] == s.tolist()
        assert all(_tb_ram_symbol(ram, p) == 0 for p in range(n, 32 * ram.shape[0]))  # padding = 'A' like the tb
        assert ta.unpack_seq_ram128(ram, n).tolist() == s.tolist()
    L = ta.library()
    buf = np.zeros(16, np.uint8)
    s = np.zeros(33, np.uint8)
    u8p = C.POINTER(C.c_uint8)
    assert L.trialign_pack_seq_ram128(s.ctypes.data_as(u8p), 33, buf.ctypes.data_as(u8p), 1) == -8   # ECAPACITY
    assert L.trialign_pack_seq_ram128(None, 3, buf.ctypes.data_as(u8p), 1) == -1                      # EINVAL
    assert L.trialign_unpack_seq_ram128(buf.ctypes.data_as(u8p), 5, s.ctypes.data_as(u8p), 4) == -8


def test_tb_init_block_lines(ta, golden, tmp_path):
    A, B, Cs = (np.array(golden["dat"][k], np.uint8) for k in "ABC")
    p = str(tmp_path / "init.sv")
    ta.write_tb_init(p, A, B, Cs, 512)
    lines = open(p).read().split("\n")
    assert lines[-1] == "" and len(lines) == 3 * 512 + 1
    names = "ATCG"
    assert lines[0] == f"seqA_ram[0][3:0] <= {names[A[0]]};"
    assert lines[63] == f"seqA_ram[1][127:124] <= {names[A[63]]};"
    assert lines[64] == "seqA_ram[2][3:0] <= A;"                      # beyond the length: 'A', as the tb ships
    assert lines[512 + 33] == f"seqB_ram[1][7:4] <= {names[B[33]]};"
    assert lines[1024 + 5] == f"seqC_ram[0][23:20] <= {names[Cs[5]]};"
    with pytest.raises(ta.TriAlignError):
        ta.write_tb_init(p, np.zeros(600, np.uint8), B, Cs, 512)       # longer than the RAM


@pytest.mark.skipif(not os.path.exists("/root/reference/src/TriAlign_tb.sv"), reason="reference only in the build container")
def test_tb_generator_reproduces_the_shipped_testbench(ta, tmp_path):
    """tools/make_tb.py's rewrite, fed the triple the reference testbench ships (all-'A', lengths 64), must give the
    reference's own file back except for the commented-out FSDB calls -- i.e. the generated stimulus is in exactly
    the form the RTL's environment expects; with the dat/ triple only lengths and symbol names change."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_tb", os.path.join(ROOT, "tools", "make_tb.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    tb = open("/root/reference/src/TriAlign_tb.sv").read()
    z = np.zeros(64, np.uint8)
    p = str(tmp_path / "init.sv")
    ta.write_tb_init(p, z, z, z, 512)
    out = mk.rewrite_testbench(tb, open(p).read(), 64, 64, 64, 12)
    diff = [(a, b) for a, b in zip(tb.split("\n"), out.split("\n")) if a != b]
    assert len(tb.split("\n")) == len(out.split("\n"))
    assert len(diff) == 2 and all("$fsdbDump" in a and b.strip().startswith("// $fsdbDump") for a, b in diff)
    A = ta.read_seq_dat("/root/reference/dat/A_seq.dat")
    ta.write_tb_init(p, A[:40], A[:24], A[:16], 512)
    out2 = mk.rewrite_testbench(tb, open(p).read(), 40, 24, 16, 10)
    assert "localparam A_LENGTH = 40;" in out2 and "localparam B_LENGTH = 24;" in out2 and "localparam C_LENGTH = 16;" in out2
    assert "localparam SCORE_BITS = 10;" in out2 and "seqA_ram[1][31:28] <= " in out2


def test_scoring_struct_defaults_and_argument_errors(ta):
    """trialign_scoring_rtl fills in the PE localparams (src/PE_1cyc.v:53-56); out-of-range parameters are refused
    before any device work (codes, not crashes)."""
    import ctypes as C

    L = ta.library()
    sc = ta.Scoring(0, 0, 0, 0, 7)
    L.trialign_scoring_rtl(C.byref(sc))
    assert sc.as_list() == [1, -1, 2, 1, 0] and ta.Scoring().as_list() == [1, -1, 2, 1, 0]
    a = np.zeros(4, np.uint8)
    s = C.c_int32(0)
    u8p = C.POINTER(C.c_uint8)
    for bad in (ta.Scoring(gap_open=-1), ta.Scoring(match=5000), ta.Scoring(triple_rule=2), ta.Scoring(gap_extend=2000)):
        rc = L.trialign_score_ex(a.ctypes.data_as(u8p), 4, a.ctypes.data_as(u8p), 4, a.ctypes.data_as(u8p), 4, 12,
                                 C.byref(bad), 0, C.byref(s), None)
        assert rc == -1
    rc = L.trialign_score_ex(a.ctypes.data_as(u8p), 4, a.ctypes.data_as(u8p), 4, a.ctypes.data_as(u8p), 4, 12, None, 0,
                             C.byref(s), None)
    assert rc == -1


def test_options_are_explicit_and_checked(ta):
    """Experiment switches are library options (trialign_set_option), never environment variables."""
    assert ta.get_option("k2_gate") == ""
    ta.set_option("k2_gate", 8)
    assert ta.get_option("k2_gate") == "8"
    ta.set_option("k2_gate", None)
    assert ta.get_option("k2_gate") == ""
    with pytest.raises(ta.TriAlignError) as ei:
        ta.set_option("no_such_option", "1")
    assert ei.value.rc == -1
    L = ta.library()
    assert L.trialign_set_option(None, b"1") == -1
    buf = C.create_string_buffer(2)
    ta.set_option("multi", "cols")
    assert L.trialign_get_option(b"multi", buf, 2) == -8  # ECAPACITY
    ta.set_option("multi", None)
    # the product sources do not read the environment
    src = os.path.join(ROOT, "hw-accelerator-three-sequence-alignment_b200", "csrc")
    for name in os.listdir(src):
        with open(os.path.join(src, name)) as f:
            assert "getenv" not in f.read(), name


def test_last_devices_before_any_call(ta):
    assert ta.last_devices() >= 0
    assert ta.library().trialign_slab_score_dev(0, None, 1, None, 1, None, 1, 0, 1, None, None, None, None, 0, 16, None,
                                                None) == -1


def test_c_host_example_links_against_the_c_abi(ta, tmp_path):
    """The C host of INTEGRATION.md section 2(a) (plus the device-mask, options and explicit-scoring calls) compiles with
    plain gcc against include/trialign.h, links to lib/libtrialign.so and -- on a box without a GPU -- fails loudly with
    the library's own error text instead of computing anything on the CPU."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = os.path.join(ROOT, "tests", "examples", "host_example.c")
    exe = str(tmp_path / "host_example")
    libdir = os.path.dirname(ta.library_path())
    subprocess.run([gcc, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", libdir, "-ltrialign", "-o", exe],
                   check=True, capture_output=True)
    env = dict(os.environ, LD_LIBRARY_PATH=libdir + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    proc = subprocess.run([exe], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    if ta.device_count() == 0:
        assert proc.returncode == 1 and "no usable CUDA device" in proc.stderr
