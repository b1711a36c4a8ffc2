"""CPU check of the fast kernels' choreography: the __host__ __device__ per-thread step functions of
csrc/fast_*.cuh are run by a sequential host emulation (tests/emu) and compared with the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)


@pytest.fixture(scope="module")
def emu():
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-C", d], check=True, capture_output=True)
    L = C.CDLL(os.path.join(d, "build", "libemu.so"))
    L.emu_fast_batch.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    return L


def _run(L, a, b, c, pairs_per_cta):
    """All CTA configurations of K1 must agree: 256 threads x 4x4 cells / 512 threads x 4x2 cells, barrier lag 1 / 2
    (lag 2 also with warps drifting in a pseudo-random order)."""
    n, ln = a.shape
    outs = []
    for th, lag, seed in [(256, 1, 0), (512, 1, 0), (256, 2, 0), (256, 2, 7), (256, 2, 11)]:
        out = np.full(n, -99999, np.int32)
        rc = L.emu_fast_batch(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln, 2 * 64 + 32,
                              pairs_per_cta, th, lag, seed, out.ctypes.data_as(i32p))
        assert rc == 0
        outs.append(out)
    # K1 "P2": two planes per step (fast_batch2_kernel)
    out = np.full(n, -99999, np.int32)
    L.emu_fast_batch2.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, i32p]
    rc = L.emu_fast_batch2(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln, 2 * 64 + 32,
                           pairs_per_cta, out.ctypes.data_as(i32p))
    assert rc == 0
    outs.append(out)
    # the same with neighbour-only synchronisation (fast_batch2r_kernel), warps in pseudo-random order
    L.emu_fast_batch2_ring.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    for seed in (1, 8, 23):
        out = np.full(n, -99999, np.int32)
        rc = L.emu_fast_batch2_ring(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln, 2 * 64 + 32,
                                    pairs_per_cta, seed, out.ctypes.data_as(i32p))
        assert rc == 0
        outs.append(out)
    assert all(o.tolist() == outs[0].tolist() for o in outs)
    return outs[0]


@pytest.mark.parametrize("ln,n,ppc", [(64, 7, 3), (64, 4, 4), (40, 5, 1), (1, 4, 2), (63, 3, 5), (17, 9, 4), (9, 2, 1)])
def test_batch_kernel_choreography(emu, ob, ln, n, ppc):
    rng = np.random.default_rng(ln * 1000 + n)
    a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    for t in range(0, n, 2):  # mutated copies exercise large positive values
        a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
    assert _run(emu, a, b, c, ppc).tolist() == ob.score_batch(a, b, c, 16).tolist()


def test_batch_kernel_extremes(emu, ob):
    z = np.zeros((2, 64), np.uint8)
    o = np.ones((2, 64), np.uint8)
    t = np.full((2, 64), 2, np.uint8)
    assert _run(emu, z, z, z, 1).tolist() == [192, 192]           # tb as shipped
    assert _run(emu, z, o, t, 1).tolist() == ob.score_batch(z, o, t, 16).tolist()  # all-mismatch: most negative


# ---- K1 with run-time scoring parameters (fast_batch_kernel_scored) ----
SCORINGS = [(1, -1, 2, 1, 0), (2, -1, 2, 1, 0), (2, -1, 2, 1, 1), (1, -1, 2, 1, 1), (3, -2, 3, 1, 0), (2, 0, 0, 0, 0),
            (4, -4, 8, 2, 1), (8, -8, 8, 4, 0), (1, -2, 4, 2, 1), (3, 1, 1, 1, 0), (0, 0, 1, 0, 1)]


@pytest.mark.parametrize("sc", SCORINGS)
def test_batch_kernel_with_runtime_scoring(emu, ob, sc):
    """The factored update with MATCH / MISMATCH / GO / GE (and the triple rule) as run-time values, on the range-checked
    lagged-barrier kernel with warps drifting: every triple the check clears must carry the oracle's score; at 16 bits
    nothing may be flagged, at a narrow width the flagged ones are exactly allowed to differ."""
    emu.emu_fast_batch_ex.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, i32p, C.POINTER(C.c_int)]
    M, X, GO, GE, rule = sc
    rng = np.random.default_rng(sum(abs(v) for v in sc) * 7 + rule)
    for ln, n, ppc in [(64, 6, 2), (37, 5, 3), (8, 3, 1)]:
        a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        for t in range(0, n, 2):
            a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
        for w, seed in [(16, 0), (16, 9), (9, 3)]:
            out = np.full(n, -99999, np.int32)
            flags = np.zeros((n + 1) // 2, np.int32)
            rc = emu.emu_fast_batch_ex(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln, ppc, seed,
                                       M, X, GO, GE, rule, w, out.ctypes.data_as(i32p),
                                       flags.ctypes.data_as(C.POINTER(C.c_int)))
            assert rc == 0
            want = [ob.score_ex(a[t], b[t], c[t], w, sc)[0] for t in range(n)]
            flagged = [bool(flags[t // 2] >> (t & 1) & 1) for t in range(n)]
            if w == 16:
                assert not any(flagged), (sc, ln, flags)
            for t in range(n):
                if not flagged[t]:
                    assert out[t] == want[t], (sc, ln, w, t, out.tolist(), want)


@pytest.mark.parametrize("sc", [(2, -1, 2, 1, 0), (1, -1, 2, 1, 1), (3, -2, 3, 1, 0), (2, 0, 0, 0, 0), (4, -4, 8, 2, 1)])
def test_stream_kernel_with_runtime_scoring(emu, ob, sc):
    """K2S with run-time scoring parameters: faces re-based with pi = 2 GE per block, zero borders, range check."""
    emu.emu_fast_stream_ex.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int, i32p, C.POINTER(C.c_int)]
    M, X, GO, GE, rule = sc
    rng = np.random.default_rng(sum(abs(v) for v in sc) * 5 + rule)
    for ln, n, ppc in [(130, 4, 2), (70, 3, 1)]:
        a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        a[0], b[0], c[0] = ob.mutated_triple(rng, ln)
        for w, seed in [(16, 4), (10, 0)]:
            out = np.full(n, -99999, np.int32)
            flags = np.zeros((n + 1) // 2, np.int32)
            rc = emu.emu_fast_stream_ex(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln, ppc, seed,
                                        M, X, GO, GE, rule, w, out.ctypes.data_as(i32p),
                                        flags.ctypes.data_as(C.POINTER(C.c_int)))
            assert rc == 0
            want = [ob.score_ex(a[t], b[t], c[t], w, sc)[0] for t in range(n)]
            flagged = [bool(flags[t // 2] >> (t & 1) & 1) for t in range(n)]
            if w == 16:
                assert not any(flagged), (sc, ln, flags)
            for t in range(n):
                if not flagged[t]:
                    assert out[t] == want[t], (sc, ln, w, t, out.tolist(), want)


# ---- K2S: streaming block sweep for batches of triples longer than 64 (fast_stream.cuh) ----
@pytest.mark.parametrize("ln,n,ppc", [(61, 9, 5), (64, 7, 4), (65, 3, 2), (100, 5, 2), (128, 4, 1), (130, 7, 4), (200, 5, 3), (256, 2, 1)])
def test_stream_kernel_choreography(emu, ob, ln, n, ppc):
    """All blocks of a CTA's pairs streamed back to back (task hand-over inside the array, faces through the rings,
    staging of the next task while two are in flight), warps drifting in different pseudo-random orders."""
    emu.emu_fast_stream.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    rng = np.random.default_rng(ln * 31 + n)
    a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    for t in range(0, n, 2):  # mutated copies exercise large positive values
        a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
    want = ob.score_batch(a, b, c, 16).tolist()
    emu.emu_fast_stream_dist.argtypes = emu.emu_fast_stream.argtypes
    # both forms of the kernel: block-edge duties on the edge threads / dealt to all threads by thread id
    for fn in (emu.emu_fast_stream, emu.emu_fast_stream_dist):
        for seed in (0, 5, 12):
            out = np.full(n, -99999, np.int32)
            rc = fn(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln,
                    2 * ln + 8 + 32, ppc, seed, out.ctypes.data_as(i32p))
            assert rc == 0
            assert out.tolist() == want, (seed, fn is emu.emu_fast_stream_dist)


@pytest.mark.parametrize("ln,n,ppc", [(65, 3, 2), (66, 5, 3), (100, 5, 2), (127, 4, 4), (128, 4, 1), (129, 3, 2), (131, 7, 4),
                                      (200, 5, 3), (255, 2, 1), (256, 2, 1)])
def test_stream2_kernel_choreography(emu, ob, ln, n, ppc):
    """K2S with two planes per step (fast_stream2.cuh): odd lengths run a padding plane, the block-edge duties are dealt
    to all threads by thread id (halos in one step ahead, faces out one step later)."""
    emu.emu_fast_stream2.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    rng = np.random.default_rng(ln * 17 + n)
    a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    for t in range(0, n, 2):  # mutated copies exercise large positive values
        a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
    want = ob.score_batch(a, b, c, 16).tolist()
    for seed in (0, 5, 12):
        out = np.full(n, -99999, np.int32)
        rc = emu.emu_fast_stream2(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln,
                                  2 * (ln + 1) + 8 + 32, ppc, seed, out.ctypes.data_as(i32p))
        assert rc == 0
        assert out.tolist() == want, seed


@pytest.mark.parametrize("sc", [None, (2, -1, 2, 1, 0), (1, -1, 2, 1, 1), (4, -4, 8, 2, 1), (2, 0, 0, 0, 0)])
def test_stream2_kernel_checked_and_with_runtime_scoring(emu, ob, sc):
    """The range-checked form (sc None: the RTL's constants, every 8th plane sampled) and the run-time-scoring form of the
    two-planes-per-step kernel: triples the check clears must carry the oracle's score, nothing is flagged at 16 bits."""
    emu.emu_fast_stream2_ex.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, i32p, C.POINTER(C.c_int)]
    M, X, GO, GE, rule = sc if sc else (1, -1, 2, 1, -1)
    rng = np.random.default_rng(99 + (sum(abs(v) for v in sc) if sc else 0))
    for ln, n, ppc in [(131, 4, 2), (70, 3, 1)]:
        a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        a[0], b[0], c[0] = ob.mutated_triple(rng, ln)
        for w, seed in [(16, 4), (9, 0)]:
            out = np.full(n, -99999, np.int32)
            flags = np.zeros((n + 1) // 2, np.int32)
            rc = emu.emu_fast_stream2_ex(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln, ppc, seed,
                                         M, X, GO, GE, rule, w, out.ctypes.data_as(i32p),
                                         flags.ctypes.data_as(C.POINTER(C.c_int)))
            assert rc == 0
            want = ([ob.score_ex(a[t], b[t], c[t], w, sc)[0] for t in range(n)] if sc
                    else ob.score_batch(a, b, c, w).tolist())
            flagged = [bool(flags[t // 2] >> (t & 1) & 1) for t in range(n)]
            if w == 16:
                assert not any(flagged), (sc, ln, flags)
            elif ln > 100 and M > 0:
                assert flagged[0], (sc, ln, flags)  # the mutated copy of length 131 scores far beyond 9 bits
            for t in range(n):
                if not flagged[t]:
                    assert out[t] == want[t], (sc, ln, w, t, out.tolist(), want)


# ---- K2L: one long triple on the K1/K2S step engine (fast_long.cuh) ----
@pytest.fixture(scope="module")
def emu_long2(emu):
    emu.emu_fast_long2.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   i32p, C.c_int, C.POINTER(C.c_int)]
    emu.emu_fast_long2_slabs.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p]
    return emu


def _long2(L, a, b, c, n_ctas, lag, gate=4, seed=1, check_bits=0):
    a, b, c = (np.ascontiguousarray(x, dtype=np.uint8) for x in (a, b, c))
    out = np.zeros(8, np.int32)
    viol = C.c_int(0)
    base = 2 * min(len(a), len(b), len(c)) + 64
    rc = L.emu_fast_long2(a.ctypes.data_as(u8p), len(a), b.ctypes.data_as(u8p), len(b), c.ctypes.data_as(u8p), len(c),
                          base, n_ctas, lag, gate, seed, out.ctypes.data_as(i32p), check_bits, C.byref(viol))
    assert rc == 0, "emulated CTAs deadlocked" if rc == -2 else rc
    return int(out[0]), out[1:].tolist(), int(viol.value)


@pytest.mark.parametrize("shape,n_ctas", [((5, 5, 5), 1), ((1, 5, 1), 1), ((20, 64, 64), 1), ((30, 64, 128), 1),
                                          ((40, 70, 130), 2), ((33, 130, 200), 3), ((100, 200, 300), 4),
                                          ((50, 129, 257), 5), ((10, 1, 1), 1), ((17, 260, 70), 2),
                                          ((260, 70, 130), 2)])  # long enough for the compact middle phase at both lags
def test_long2_choreography(emu_long2, ob, shape, n_ctas):
    """Both barrier lags, two publication rhythms, warps and CTAs advanced in pseudo-random order."""
    la, lb, lc = shape
    rng = np.random.default_rng(la * 7 + lb * 3 + lc)
    for mutated in (True, False):
        if mutated:
            a, b, c = ob.mutated_triple(rng, max(shape))
            a, b, c = a[:la], b[:lb], c[:lc]
        else:
            a, b, c = (rng.integers(0, 4, n, dtype=np.uint8) for n in shape)
        es, efin = ob.score(a, b, c, 16)
        for lag, gate, seed in ((1, 4, 1), (2, 4, 2), (2, 16, 3), (1, 8, 4), (1, 4, 6)):  # seeds 4, 6: row-major claim order
            s, fin, _ = _long2(emu_long2, a, b, c, n_ctas, lag, gate, seed)
            assert s == es and fin == efin.tolist(), (lag, gate, seed)


@pytest.mark.parametrize("shape", [(1, 154, 350), (1, 130, 365), (2, 70, 300), (3, 64, 129), (1, 64, 256)])
def test_long2_dead_low_half_stays_bounded(emu_long2, ob, shape):
    """Very short sweeps (la < 16) on as many CTAs as tasks, shared memory pre-filled with 0xFF / 0x00: after plane la
    the low half of a word is dead while the high half (16 planes behind) is still live; a dead lane fed from an
    unwritten halo slot would carry into the live one (found by tools/stress_gpu.py on the GPU)."""
    la, lb, lc = shape
    rng = np.random.default_rng(sum(shape))
    tasks = ((lb + 63) // 64) * ((lc + 127) // 128)
    for rate in (0.01, 0.2, 0.75):
        a, b, c = ob.mutated_triple(rng, max(shape), rate=rate)
        a, b, c = a[:la], b[:lb], c[:lc]
        es, efin = ob.score(a, b, c, 16)
        for lag, seed in ((1, 2), (1, 5), (2, 3), (2, 4)):
            s, fin, _ = _long2(emu_long2, a, b, c, tasks, lag, 4, seed)
            assert s == es and fin == efin.tolist(), (rate, lag, seed)


def test_long2_range_check_is_sound_and_not_trigger_happy(emu_long2, ob):
    rng = np.random.default_rng(77)
    a, b, c = ob.mutated_triple(rng, 150, rate=0.02)        # nearly identical: all grows to ~3*150 = 450
    _, _, (lo, hi) = ob.score(a, b, c, 32, want_range=True)
    for lag in (1, 2):
        for w in (9, 10, 11, 12):
            lim = 1 << (w - 1)
            _, _, viol = _long2(emu_long2, a, b, c, 2, lag, check_bits=w)
            if hi + 3 > lim - 1 or lo - 4 < -lim:
                assert viol, (w, lo, hi)                       # soundness: never miss a possible wrap
            if hi + 3 + 45 + 12 <= lim - 1 and lo - 8 - 60 - 12 >= -lim:
                assert not viol, (w, lo, hi)                   # and no false alarm when far from the limits
        x, y, z = (rng.integers(0, 4, 150, dtype=np.uint8) for _ in range(3))
        assert _long2(emu_long2, x, y, z, 2, lag, check_bits=10)[2] == 0


@pytest.mark.parametrize("shape,n_slabs,n_ctas,lag", [((40, 70, 130), 2, 2, 1), ((64, 64, 64), 4, 1, 2),
                                                      ((100, 129, 257), 3, 3, 1), ((37, 200, 140), 8, 2, 2)])
def test_long2_i_slab_pipeline(emu_long2, ob, shape, n_slabs, n_ctas, lag):
    la, lb, lc = shape
    rng = np.random.default_rng(sum(shape) + n_slabs)
    a, b, c = ob.mutated_triple(rng, max(shape))
    a, b, c = (np.ascontiguousarray(x) for x in (a[:la], b[:lb], c[:lc]))
    out = np.zeros(8, np.int32)
    rc = emu_long2.emu_fast_long2_slabs(a.ctypes.data_as(u8p), la, b.ctypes.data_as(u8p), lb, c.ctypes.data_as(u8p), lc,
                                        2 * min(shape) + 64, n_ctas, n_slabs, lag, out.ctypes.data_as(i32p))
    assert rc == 0
    es, efin = ob.score(a, b, c, 16)
    assert int(out[0]) == es and out[1:].tolist() == efin.tolist()


@pytest.mark.parametrize("shape,n_ranges,n_ctas", [((40, 70, 300), 2, 2), ((33, 130, 520), 4, 2), ((100, 64, 1000), 8, 1),
                                                   ((1, 100, 400), 3, 3), ((60, 200, 700), 3, 4)])
def test_long2_column_pipeline_concurrent_ranges(emu_long2, ob, shape, n_ranges, n_ctas):
    """The multi-GPU column pipeline: several emulated GPUs run at the same time, each sweeping a block-column range
    with its own counter / claim order / rings, boundary faces and epoch-tagged progress words in the neighbour's
    memory; two epochs over the same buffers (the faces poisoned in between).  A consumer that is not gated by its
    remote producer reads poison here."""
    emu_long2.emu_fast_long2_cols.argtypes = [u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, C.c_int, i32p]
    la, lb, lc = shape
    rng = np.random.default_rng(sum(shape) + n_ranges)
    a, b, c = ob.mutated_triple(rng, max(shape))
    a, b, c = (np.ascontiguousarray(x) for x in (a[:la], b[:lb], c[:lc]))
    es, efin = ob.score(a, b, c, 16)
    for seed in (1, 2, 7):
        out = np.zeros(8, np.int32)
        rc = emu_long2.emu_fast_long2_cols(a.ctypes.data_as(u8p), la, b.ctypes.data_as(u8p), lb, c.ctypes.data_as(u8p), lc,
                                           2 * min(shape) + 64, n_ctas, n_ranges, seed, 2, out.ctypes.data_as(i32p))
        assert rc == 0, "emulated GPUs deadlocked" if rc == -2 else rc
        assert int(out[0]) == es and out[1:].tolist() == efin.tolist(), seed
