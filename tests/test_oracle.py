"""CPU tests of the checkers themselves: O1 (functional oracle) against the frozen golden vectors and against
O2 (cycle-accurate RTL model), plus the survey-derived known answers of SURVEY.md 8(c)."""
import os

import numpy as np
import pytest


def _abc(case):
    return (np.array(case[k], np.uint8) for k in ("a", "b", "c"))


def test_known_answers_of_survey(golden, ob):
    by = {c["name"]: c for c in golden["cases"]}
    assert by["dat_triple"]["score"] == 1
    assert by["dat_triple"]["final_states"] == [-1, 1, -4, -5, 1, -7, -2]
    assert by["dat_triple"]["range"] == [-19, 17]
    assert by["dat_triple"]["rtl_1cyc_cycles"] == 9282
    assert by["tb_as_shipped_allA_64"]["score"] == 192
    assert by["tb_as_shipped_allA_64"]["final_states"] == [192, 185, 185, 185, 188, 188, 188]
    assert [by[f"dat_prefix_{L}"]["score"] for L in range(8, 65, 8)] == [-3, -5, -2, 1, -1, 5, 2, 1]
    assert by["dat_prefix_16"]["final_states"] == [-5, -6, -6, -7, -5, -6, -5]
    assert by["dat_prefix_16"]["rtl_1cyc_cycles"] == 390
    assert [by[f"dat_perm_{p}"]["score"] for p in ("ACB", "BAC", "BCA", "CAB", "CBA")] == [0, 1, -4, 0, -4]
    assert by["wrap_W7_allA_24"]["score"] == 62 and by["wrap_W7_allA_24"]["rtl_1cyc_cycles"] == 947
    assert by["wrap_W8_allA_48"]["score"] == 125 and by["wrap_W8_allA_48"]["rtl_1cyc_cycles"] == 4646


def test_oracle_reproduces_golden(golden, ob):
    for case in golden["cases"]:
        a, b, c = _abc(case)
        s, fin, rng = ob.score(a, b, c, case["score_bits"], want_range=True)
        assert s == case["score"], case["name"]
        assert fin.tolist() == case["final_states"], case["name"]
        assert list(rng) == case["range"], case["name"]


def test_wavefront_oracle_equals_serial(golden, ob):
    for case in golden["cases"][:25]:
        a, b, c = _abc(case)
        s, fin = ob.score_mt(a, b, c, case["score_bits"], threads=3)
        assert s == case["score"] and fin.tolist() == case["final_states"], case["name"]


def test_identical_sequences_score_3L(ob):
    rng = np.random.default_rng(3)
    for L in (1, 7, 40, 100):
        a = rng.integers(0, 4, L, dtype=np.uint8)
        assert ob.score(a, a, a, 16)[0] == 3 * L


def test_zero_length_is_zero(ob):
    a = np.array([1, 2, 3], np.uint8)
    e = np.zeros(0, np.uint8)
    assert ob.score(e, a, a)[0] == 0
    assert ob.score(a, e, a)[0] == 0
    assert ob.score(a, a, e)[0] == 0


def test_temp_abc_precedence_quirk(ob):
    # A==B!=C scores 0, A!=B scores -3 even when A==C or B==C (PE_1cyc.v:162); single-cell problems expose it
    cube = ob.cube([0], [0], [1])
    assert cube[1, 1, 1, 0] == 0
    assert ob.cube([0], [1], [0])[1, 1, 1, 0] == -3
    assert ob.cube([1], [0], [0])[1, 1, 1, 0] == -3
    assert ob.cube([2], [2], [2])[1, 1, 1, 0] == 3


def test_rtl_model_equals_oracle_random(ob):
    rng = np.random.default_rng(11)
    for (la, lb, lc) in [(8, 8, 8), (16, 16, 16), (40, 16, 16), (24, 16, 8), (48, 40, 32), (64, 64, 64), (72, 64, 56)]:
        for mutated in (False, True):
            if mutated:
                a, b, c = ob.mutated_triple(rng, max(la, lb, lc))
                a, b, c = a[:la], b[:lb], c[:lc]
            else:
                a, b, c = (rng.integers(0, 4, n, dtype=np.uint8) for n in (la, lb, lc))
            s, fin = ob.score(a, b, c, 12)
            rc, rs, rfin, cyc = ob.rtl_run(a, b, c, 1)
            assert rc == ob.RTL_OK
            assert rs == s and rfin.tolist() == fin.tolist()
            assert cyc == ob.rtl_cycle_formula(1, 8, la, lb, lc)


def test_rtl_model_wrap_order_is_observable(ob):
    # all-A L=24 at W=7: true score 72 > 63; the RTL gives 62, not wrap(72) -- candidates wrap before the max
    z = np.zeros(24, np.uint8)
    rc, rs, rfin, _ = ob.rtl_run(z, z, z, 1, score_bits=7)
    assert rc == ob.RTL_OK and rs == 62 and rfin.tolist() == [62, 59, 59, 59, 62, 62, 62]
    assert ob.score(z, z, z, 7)[0] == 62


def test_rtl_model_domain_guard(ob):
    rng = np.random.default_rng(5)
    a, b, c = (rng.integers(0, 4, n, dtype=np.uint8) for n in (16, 24, 24))
    assert ob.rtl_run(a, b, c, 1)[0] == ob.RTL_EDOMAIN      # A_idx < B_idx with C_idx > P
    assert ob.rtl_run(a[:10], b[:12], c[:8], 1)[0] == ob.RTL_EDOMAIN  # B not a multiple of PE_LEN
    # outside the domain the raw model reproduces the RTL's own mis-computation instead of the recurrence
    rc, rs, _, cyc = ob.rtl_run(a, b, c, 1, strict=False)
    assert rc in (ob.RTL_OK, ob.RTL_EX) and cyc == ob.rtl_cycle_formula(1, 8, 16, 24, 24)


def test_rtl_model_counter_width_hang(ob):
    # input_counter is SCORE_BITS wide: at W<7 it can never reach P*P=64 and the FSM stays in INITIAL
    z = np.zeros(8, np.uint8)
    rc = ob.rtl_run(z, z, z, 1, score_bits=6, strict=False, max_cycles=5000)[0]
    assert rc == ob.RTL_ETIMEOUT


def test_synth_generator_properties(golden, ob):
    for ent in golden["synth"]:
        if "first_symbols" in ent:
            for s, nm in enumerate("ABC"):
                assert ob.synth(ent["seed"], 0, s, 16).tolist() == ent["first_symbols"][nm]
    x = ob.synth(0x5EED0002, 12345, 1, 4096)
    assert set(np.unique(x)) == {0, 1, 2, 3}
    assert abs(np.bincount(x, minlength=4) / 4096 - 0.25).max() < 0.05


def test_batch_oracle_matches_golden(golden, ob):
    ent = next(e for e in golden["synth"] if e["config"] == 2)
    n = len(ent["scores"])
    a, b, c = ob.synth_batch(ent["seed"], 0, n, ent["len"])
    assert ob.score_batch(a, b, c, ent["score_bits"]).tolist() == ent["scores"]


def test_rtl_2cyc_model_equals_oracle_and_published_cycle_counts(ob, golden):
    """PE_2cyc + TriAlign_2cyc compute the same function with 2 clocks per wavefront step; the cycle counts are the
    ones that reproduce the paper's FPGA/ASIC runtimes (14,402 @ L=64, 90,370 @ L=128)."""
    d = golden["dat"]
    rc, s, fin, cyc = ob.rtl_run(d["A"], d["B"], d["C"], 2)
    assert (rc, s, fin.tolist(), cyc) == (0, 1, [-1, 1, -4, -5, 1, -7, -2], 14402)
    z = np.zeros(64, np.uint8)
    rc, s, fin, cyc = ob.rtl_run(z, z, z, 2)
    assert (rc, s, fin.tolist(), cyc) == (0, 192, [192, 185, 185, 185, 188, 188, 188], 14402)
    rng = np.random.default_rng(21)
    for L in (64, 128):
        for mutated in (False, True):
            a, b, c = ob.mutated_triple(rng, L) if mutated else tuple(rng.integers(0, 4, L, dtype=np.uint8) for _ in range(3))
            es, efin = ob.score(a, b, c, 12)
            rc, s, fin, cyc = ob.rtl_run(a, b, c, 2)
            assert rc == ob.RTL_OK and s == es and fin.tolist() == efin.tolist()
            assert cyc == ob.rtl_cycle_formula(2, 8, L, L, L) == {64: 14402, 128: 90370}[L]


def test_rtl_2cyc_bank_select_quirk(ob):
    """Outside its design points the 2cyc control reads and writes one single-port SRAM at once: the model's X
    poisoning reaches the score at L=24 and L=32, while 8, 16, 40, 48 still come out right."""
    rng = np.random.default_rng(0)
    for L, ok in ((8, True), (16, True), (24, False), (32, False), (40, True), (48, True)):
        a, b, c = (rng.integers(0, 4, L, dtype=np.uint8) for _ in range(3))
        assert ob.rtl_run(a, b, c, 2)[0] == ob.RTL_EDOMAIN           # strict mode refuses all of them
        rc, s, fin, cyc = ob.rtl_run(a, b, c, 2, strict=False)
        assert cyc == ob.rtl_cycle_formula(2, 8, L, L, L)
        if ok:
            assert rc == ob.RTL_OK and s == ob.score(a, b, c, 12)[0]
        else:
            assert rc == ob.RTL_EX


# ---- traceback ----
def test_traceback_of_identical_sequences_is_the_main_diagonal(ob):
    rng = np.random.default_rng(2)
    a = rng.integers(0, 4, 40, dtype=np.uint8)
    s, start, moves = ob.align(a, a, a, 12)
    assert s == 3 * 40 and start == (0, 0, 0) and moves.tolist() == [7] * 40


@pytest.mark.parametrize("shape", [(1, 1, 1), (5, 7, 3), (24, 24, 24), (40, 17, 33), (48, 48, 48)])
def test_traceback_path_rescored_independently_gives_the_score(ob, shape):
    """Without wrap the path the traceback returns must be worth exactly the score, it must end at (la,lb,lc) and
    begin on a zero border plane."""
    rng = np.random.default_rng(sum(shape))
    for mutated in (True, False):
        if mutated:
            a, b, c = ob.mutated_triple(rng, max(shape))
            a, b, c = a[: shape[0]], b[: shape[1]], c[: shape[2]]
        else:
            a, b, c = (rng.integers(0, 4, n, dtype=np.uint8) for n in shape)
        s, start, moves = ob.align(a, b, c, 16)
        assert s == ob.score(a, b, c, 16)[0]
        assert min(start) == 0 and all(m in (1, 2, 3, 4, 5, 6, 7) for m in moves)
        assert ob.rescore_path(a, b, c, start, moves) == s


def test_traceback_follows_wrapped_values_too(ob):
    """At a score_bits where candidates wrap the traceback still follows what the RTL's MAX7 instances would select
    (internal check of oracle_align: every step's selected candidate equals the cell's stored state)."""
    z = np.zeros(24, np.uint8)
    s, start, moves = ob.align(z, z, z, 7)   # all-'A' L=24 at W=7: score 62, not wrap(72)
    assert s == 62 and len(moves) >= 24
    rng = np.random.default_rng(9)
    for w in (5, 6, 7, 8):
        a, b, c = ob.mutated_triple(rng, 30, rate=0.05)
        s, start, moves = ob.align(a, b, c, w)
        assert s == ob.score(a, b, c, w)[0]


def test_traceback_golden_paths(ob, golden):
    """tests/golden/golden.json freezes oracle_align's path for every golden case (make_golden.py also re-scores each
    wrap-free one); the oracle must keep reproducing them."""
    for case in golden["cases"]:
        a, b, c = (np.array(case[k], np.uint8) for k in "abc")
        s, start, moves = ob.align(a, b, c, case["score_bits"])
        assert s == case["score"], case["name"]
        assert list(start) == case["traceback"]["start"], case["name"]
        assert "".join(str(int(m)) for m in moves) == case["traceback"]["moves"], case["name"]


# ---- explicit scoring parameters ----
def test_score_ex_with_rtl_constants_is_the_literal_oracle(ob, golden):
    """oracle_score_ex is written independently of pe_cell() (table-driven); with the PE localparams
    {MATCH=1, MISMATCH=-1, GO=2, GE=1, temp_ABC rule} it must reproduce oracle_score on score and all 7 states,
    including where candidates wrap."""
    import numpy as np

    rng = np.random.default_rng(21)
    d = golden["dat"]
    s, fin = ob.score_ex(d["A"], d["B"], d["C"], 12, [1, -1, 2, 1, 0])
    assert s == 1 and fin.tolist() == [-1, 1, -4, -5, 1, -7, -2]
    for shape, w in [((9, 9, 9), 12), ((24, 24, 24), 7), ((17, 5, 30), 12), ((40, 33, 21), 8), ((30, 30, 30), 5)]:
        a, b, c = ob.mutated_triple(rng, max(shape), rate=0.1)
        a, b, c = a[:shape[0]], b[:shape[1]], c[:shape[2]]
        s, fin = ob.score_ex(a, b, c, w, [1, -1, 2, 1, 0])
        es, efin = ob.score(a, b, c, w)
        assert s == es and fin.tolist() == efin.tolist(), (shape, w)


def test_score_ex_sum_of_pairs_is_symmetric_and_rtl_rule_is_not(ob, golden):
    """The paper's S(a,b,c) does not care which sequence plays which role; the RTL's temp_ABC does (SURVEY.md 8(a) a3:
    the dat/ triple scores 1, 0, 1, -4, 0, -4 over the six role permutations)."""
    import itertools

    d = golden["dat"]
    seqs = [d["A"], d["B"], d["C"]]
    rtl = [ob.score_ex(seqs[p[0]], seqs[p[1]], seqs[p[2]], 12, [1, -1, 2, 1, 0])[0] for p in itertools.permutations(range(3))]
    assert sorted(rtl) == sorted([1, 0, 1, -4, 0, -4])
    sop = {ob.score_ex(seqs[p[0]], seqs[p[1]], seqs[p[2]], 12, [1, -1, 2, 1, 1])[0] for p in itertools.permutations(range(3))}
    assert len(sop) == 1
    # identical sequences: every column is a three-way match worth 3*MATCH under both rules
    x = seqs[0][:20]
    assert ob.score_ex(x, x, x, 12, [2, -1, 3, 1, 0])[0] == 3 * 2 * 20
    assert ob.score_ex(x, x, x, 12, [2, -1, 3, 1, 1])[0] == 3 * 2 * 20


def test_score_ex_reproduces_the_committed_scoring_fixture(ob):
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_scoring.json")) as f:
        g = json.load(f)
    assert len(g["cases"]) == 40
    for case in g["cases"]:
        s, fin = ob.score_ex(case["a"], case["b"], case["c"], case["score_bits"], case["scoring"])
        assert s == case["score"] and fin.tolist() == case["final_states"], (case["name"], case["scoring"])


def test_blocked_oracle_equals_serial(ob):
    """oracle_score_blocked (the RTL's slicing as a CPU schedule, used for the long goldens) against the serial form:
    ragged shapes, block sizes that do and do not divide the lengths, wrap cases, several threads."""
    rng = np.random.default_rng(20261019)
    for (la, lb, lc) in [(1, 1, 1), (5, 9, 3), (17, 33, 29), (64, 64, 64), (65, 63, 70), (100, 37, 129)]:
        for w in (7, 9, 12, 16):
            for blk in (1, 3, 8, 64):
                a, b, c = ob.mutated_triple(rng, max(la, lb, lc))
                a, b, c = a[:la], b[:lb], c[:lc]
                s, fin = ob.score(a, b, c, w)
                s2, fin2 = ob.score_blocked(a, b, c, w, 3, blk)
                assert s == s2 and fin.tolist() == fin2.tolist(), (la, lb, lc, w, blk)


def test_long_golden_file_is_reproducible(ob, ta):
    """tests/golden/golden_long.json: inputs rebuild identically from the oracle's and the product's generator, and the
    smallest case re-computes to the frozen score."""
    import sys as _sys
    _sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import long_cases

    cases = long_cases.load()
    assert "pipe_1500_w16" in cases
    for name, e in cases.items():
        a, b, c = long_cases.build_case(ob.synth, e["seed"], e["kind"], [min(n, 2000) for n in e["lens"]])
        a2, b2, c2 = long_cases.build_case(ta.synth, e["seed"], e["kind"], [min(n, 2000) for n in e["lens"]])
        assert a.tolist() == a2.tolist() and b.tolist() == b2.tolist() and c.tolist() == c2.tolist()
        assert a[:16].tolist() == e["first_symbols"]["A"] and c[:16].tolist() == e["first_symbols"]["C"], name
    e = cases["long_1024_mut_w12"]
    a, b, c = long_cases.build_case(ob.synth, e["seed"], e["kind"], e["lens"])
    s, fin = ob.score_blocked(a[:300], b[:300], c[:300], 12, 0, 64)
    assert (s, fin.tolist()) == (ob.score_mt(a[:300], b[:300], c[:300], 12)[0], ob.score_mt(a[:300], b[:300], c[:300], 12)[1].tolist())
