"""ctypes bindings for oracle/build/liboracle.so (O1 functional oracle + O2 RTL cycle model).

Test infrastructure only: the product (libtrialign) never loads this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "build", "liboracle.so")

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
i64p = C.POINTER(C.c_int64)


class RtlCfg(C.Structure):
    _fields_ = [
        ("variant", C.c_int),
        ("pe_len", C.c_int),
        ("score_bits", C.c_int),
        ("a_total_len", C.c_int),
        ("sram_addr_bits", C.c_int),
        ("strict", C.c_int),
        ("max_cycles", C.c_int64),
    ]


RTL_OK, RTL_EARG, RTL_EDOMAIN, RTL_ETIMEOUT, RTL_EX = 0, -1, -2, -3, -4

_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc/g++ only)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_score.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p, i32p]
        L.oracle_score.restype = C.c_int
        L.oracle_score_mt.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, i32p, i32p]
        L.oracle_score_mt.restype = C.c_int
        L.oracle_score_blocked.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, i32p,
                                           i32p]
        L.oracle_score_blocked.restype = C.c_int
        L.oracle_score_batch.argtypes = [u8p, u8p, u8p, C.c_int64, C.c_int, C.c_int, C.c_int, i32p]
        L.oracle_score_batch.restype = C.c_int
        L.oracle_cube.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, i32p]
        L.oracle_cube.restype = C.c_int
        L.oracle_synth.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, u8p]
        L.oracle_synth.restype = None
        L.oracle_align.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p, u8p, C.c_int,
                                   C.POINTER(C.c_int)]
        L.oracle_align.restype = C.c_int
        L.oracle_hw_threads.restype = C.c_int
        L.rtl_default_cfg.argtypes = [C.POINTER(RtlCfg), C.c_int]
        L.rtl_model_run.argtypes = [C.POINTER(RtlCfg), u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, i32p, i32p, i64p]
        L.rtl_model_run.restype = C.c_int
        L.rtl_cycle_formula.argtypes = [C.c_int] * 5
        L.rtl_cycle_formula.restype = C.c_int64
        L.rtl_model_run_batch.argtypes = [C.POINTER(RtlCfg), u8p, u8p, u8p, C.c_int64, C.c_int, C.c_int, i32p, i64p]
        L.rtl_model_run_batch.restype = C.c_int
        _lib = L
    return _lib


def _u8(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint8))


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def score(a, b, c, score_bits: int = 12, want_range: bool = False):
    """-> (score, final_states[7]) or (score, final_states, (lo, hi))."""
    a, b, c = _u8(a), _u8(b), _u8(c)
    s = C.c_int32()
    fin = np.zeros(7, np.int32)
    rng = np.zeros(2, np.int32)
    rc = lib().oracle_score(_ptr(a, u8p), len(a), _ptr(b, u8p), len(b), _ptr(c, u8p), len(c), score_bits,
                            C.byref(s), _ptr(fin, i32p), _ptr(rng, i32p))
    if rc:
        raise ValueError(f"oracle_score rc={rc}")
    if want_range:
        return int(s.value), fin, (int(rng[0]), int(rng[1]))
    return int(s.value), fin


def score_ex(a, b, c, score_bits: int, scoring):
    """Recurrence with explicit scoring [MATCH, MISMATCH, GO, GE, rule] -> (score, final_states[7])."""
    a, b, c = _u8(a), _u8(b), _u8(c)
    s = C.c_int32()
    fin = np.zeros(7, np.int32)
    sc = np.array(list(scoring), np.int32)
    L = lib()
    L.oracle_score_ex.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, C.POINTER(C.c_int32), i32p]
    rc = L.oracle_score_ex(_ptr(a, u8p), len(a), _ptr(b, u8p), len(b), _ptr(c, u8p), len(c), score_bits,
                           _ptr(sc, i32p), C.byref(s), _ptr(fin, i32p))
    if rc:
        raise ValueError(f"oracle_score_ex rc={rc}")
    return int(s.value), fin


def score_mt(a, b, c, score_bits: int = 12, threads: int = 0):
    a, b, c = _u8(a), _u8(b), _u8(c)
    s = C.c_int32()
    fin = np.zeros(7, np.int32)
    rc = lib().oracle_score_mt(_ptr(a, u8p), len(a), _ptr(b, u8p), len(b), _ptr(c, u8p), len(c), score_bits,
                               threads, C.byref(s), _ptr(fin, i32p))
    if rc:
        raise ValueError(f"oracle_score_mt rc={rc}")
    return int(s.value), fin


def score_blocked(a, b, c, score_bits: int = 12, threads: int = 0, blk: int = 128):
    """O1 on the RTL's slicing schedule (blk x blk blocks swept over all i): the fast form for long triples."""
    a, b, c = _u8(a), _u8(b), _u8(c)
    s = C.c_int32()
    fin = np.zeros(7, np.int32)
    rc = lib().oracle_score_blocked(_ptr(a, u8p), len(a), _ptr(b, u8p), len(b), _ptr(c, u8p), len(c), score_bits,
                                    threads, blk, C.byref(s), _ptr(fin, i32p))
    if rc:
        raise ValueError(f"oracle_score_blocked rc={rc}")
    return int(s.value), fin


def score_batch(a, b, c, score_bits: int = 12, threads: int = 0) -> np.ndarray:
    a, b, c = _u8(a), _u8(b), _u8(c)
    assert a.ndim == 2 and a.shape == b.shape == c.shape
    n, ln = a.shape
    out = np.zeros(n, np.int32)
    rc = lib().oracle_score_batch(_ptr(a, u8p), _ptr(b, u8p), _ptr(c, u8p), n, ln, score_bits, threads,
                                  _ptr(out, i32p))
    if rc:
        raise ValueError(f"oracle_score_batch rc={rc}")
    return out


def cube(a, b, c, score_bits: int = 12) -> np.ndarray:
    a, b, c = _u8(a), _u8(b), _u8(c)
    out = np.zeros((len(a) + 1, len(b) + 1, len(c) + 1, 7), np.int32)
    rc = lib().oracle_cube(_ptr(a, u8p), len(a), _ptr(b, u8p), len(b), _ptr(c, u8p), len(c), score_bits,
                           _ptr(out, i32p))
    if rc:
        raise ValueError(f"oracle_cube rc={rc}")
    return out


def align(a, b, c, score_bits: int = 12):
    """Traceback -> (score, start cell (i,j,k), moves uint8[n] in forward order; bit0/1/2 = A/B/C advances)."""
    a, b, c = _u8(a), _u8(b), _u8(c)
    s = C.c_int32()
    start = np.zeros(3, np.int32)
    cap = len(a) + len(b) + len(c) + 1
    moves = np.zeros(cap, np.uint8)
    n = C.c_int(0)
    rc = lib().oracle_align(_ptr(a, u8p), len(a), _ptr(b, u8p), len(b), _ptr(c, u8p), len(c), score_bits, C.byref(s),
                            _ptr(start, i32p), _ptr(moves, u8p), cap, C.byref(n))
    if rc:
        raise ValueError(f"oracle_align rc={rc}")
    return int(s.value), tuple(int(x) for x in start), moves[: n.value].copy()


# cost subtracted by destination state (row) and source state (column), order M,Ix,Iy,Iz,Ixy,Iyz,Ixz
COST = [[0, 0, 0, 0, 0, 0, 0], [4, 2, 3, 3, 3, 4, 3], [4, 3, 2, 3, 3, 3, 4], [4, 3, 3, 2, 4, 3, 3],
        [2, 1, 1, 2, 1, 2, 2], [2, 2, 1, 1, 2, 1, 2], [2, 1, 2, 1, 2, 2, 1]]
STATE_OF_MOVE = {7: 0, 1: 1, 2: 2, 4: 3, 3: 4, 6: 5, 5: 6}


def rescore_path(a, b, c, start, moves) -> int:
    """Independent (pure Python, no wrap) score of an alignment path: the sum of the column scores of SURVEY.md 8(a)
    minus the transition costs between consecutive columns.  The first column leaves a zero border cell, where all
    seven states are 0, so it pays the cheapest cost into its state."""
    i, j, k = start
    total, prev = 0, None
    for mv in (int(m) for m in moves):
        st = STATE_OF_MOVE[mv]
        i += mv & 1
        j += (mv >> 1) & 1
        k += (mv >> 2) & 1
        A, B, Cc = int(a[i - 1]) & 3, int(b[j - 1]) & 3, int(c[k - 1]) & 3
        if st == 0:
            add = 3 if (A == B == Cc) else (0 if A == B else -3)   # src/PE_1cyc.v:162 with Verilog precedence
        elif st == 4:
            add = 1 if A == B else -1
        elif st == 5:
            add = 1 if B == Cc else -1
        elif st == 6:
            add = 1 if A == Cc else -1
        else:
            add = 0
        total += add - (min(COST[st]) if prev is None else COST[st][prev])
        prev = st
    assert (i, j, k) == (len(a), len(b), len(c)), "path does not end at (la,lb,lc)"
    return total


def synth(seed: int, triple: int, which: int, length: int) -> np.ndarray:
    out = np.zeros(length, np.uint8)
    lib().oracle_synth(C.c_uint64(seed), triple, which, length, _ptr(out, u8p))
    return out


def synth_batch(seed: int, first: int, n: int, length: int):
    a = np.stack([synth(seed, first + t, 0, length) for t in range(n)])
    b = np.stack([synth(seed, first + t, 1, length) for t in range(n)])
    c = np.stack([synth(seed, first + t, 2, length) for t in range(n)])
    return a, b, c


def rtl_run(a, b, c, variant: int = 1, score_bits: int = 12, pe_len: int = 8, strict: bool = True,
            a_total_len: int = 0, sram_addr_bits: int = 0, max_cycles: int = 0):
    """-> (rc, score, final_states[7], cycles)."""
    a, b, c = _u8(a), _u8(b), _u8(c)
    cfg = RtlCfg()
    lib().rtl_default_cfg(C.byref(cfg), variant)
    cfg.score_bits, cfg.pe_len, cfg.strict = score_bits, pe_len, int(strict)
    cfg.a_total_len, cfg.sram_addr_bits, cfg.max_cycles = a_total_len, sram_addr_bits, max_cycles
    s = C.c_int32(0)
    fin = np.zeros(7, np.int32)
    cyc = C.c_int64(0)
    rc = lib().rtl_model_run(C.byref(cfg), _ptr(a, u8p), len(a), _ptr(b, u8p), len(b), _ptr(c, u8p), len(c),
                             C.byref(s), _ptr(fin, i32p), C.byref(cyc))
    return rc, int(s.value), fin, int(cyc.value)


def rtl_cycle_formula(variant: int, pe_len: int, la: int, lb: int, lc: int) -> int:
    return int(lib().rtl_cycle_formula(variant, pe_len, la, lb, lc))


def read_dat(path: str) -> np.ndarray:
    """dat/*_seq.dat: one decimal digit 0-3 per line, CRLF or LF."""
    with open(path, "rb") as f:
        txt = f.read().decode("ascii")
    return np.array([int(t) for t in txt.split()], dtype=np.uint8)


def mutated_triple(rng: np.random.Generator, length: int, rate: float = 0.15):
    """A random; B, C = A with `rate` substitutions (exercises large values)."""
    a = rng.integers(0, 4, length, dtype=np.uint8)
    out = [a]
    for _ in range(2):
        m = a.copy()
        mask = rng.random(length) < rate
        m[mask] = rng.integers(0, 4, int(mask.sum()), dtype=np.uint8)
        out.append(m)
    return out
