"""Parity tests proper (-m gpu): every call goes through the C ABI of lib/libtrialign.so and is compared with the
golden vectors / the CPU oracle on the same inputs.  Bar: bit-exact (integer scores and states)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BACKENDS = {"auto": 0, "literal": 5}


def _abc(case):
    return (np.array(case[k], np.uint8) for k in ("a", "b", "c"))


@pytest.mark.parametrize("backend", list(BACKENDS))
def test_golden_cases_single_triple(gpu_ta, golden, backend):
    for case in golden["cases"]:
        a, b, c = _abc(case)
        eng = gpu_ta.TriAlign(score_bits=case["score_bits"], backend=BACKENDS[backend])
        s, fin = eng.align(a, b, c, return_states=True)
        assert s == case["score"], (case["name"], backend, gpu_ta.last_path())
        assert fin.tolist() == case["final_states"], (case["name"], backend, gpu_ta.last_path())


def test_dat_triple_and_tb_as_shipped(gpu_ta, golden):
    d = golden["dat"]
    eng = gpu_ta.TriAlign(score_bits=12)
    assert eng.align(d["A"], d["B"], d["C"]) == 1
    z = np.zeros(64, np.uint8)
    assert eng.align(z, z, z) == 192


def test_zero_and_ragged_lengths(gpu_ta, ob):
    eng = gpu_ta.TriAlign()
    a = np.array([0, 1, 2, 3, 1], np.uint8)
    e = np.zeros(0, np.uint8)
    for x, y, z in ((e, a, a), (a, e, a), (a, a, e), (e, e, e)):
        s, fin = eng.align(x, y, z, return_states=True)
        assert s == 0 and fin.tolist() == [0] * 7
    rng = np.random.default_rng(7)
    for (la, lb, lc) in [(1, 1, 1), (1, 40, 3), (70, 2, 9), (33, 65, 17), (5, 5, 130)]:
        x, y, z = (rng.integers(0, 4, n, dtype=np.uint8) for n in (la, lb, lc))
        s, fin = eng.align(x, y, z, return_states=True)
        es, efin = ob.score(x, y, z, 12)
        assert s == es and fin.tolist() == efin.tolist(), (la, lb, lc)


@pytest.mark.parametrize("backend", list(BACKENDS))
def test_synthetic_batch_config2_prefix(gpu_ta, golden, backend):
    ent = next(e for e in golden["synth"] if e["config"] == 2)
    n = len(ent["scores"])
    a, b, c = gpu_ta.synth_batch(ent["seed"], 0, n, ent["len"])
    got = gpu_ta.TriAlign(ent["score_bits"], backend=BACKENDS[backend]).align_batch(a, b, c)
    assert got.tolist() == ent["scores"]


def test_batch_against_oracle_random_and_mutated(gpu_ta, ob):
    rng = np.random.default_rng(99)
    for L, n in ((8, 65), (24, 33), (64, 130), (100, 9)):
        a = rng.integers(0, 4, (n, L), dtype=np.uint8)
        b = rng.integers(0, 4, (n, L), dtype=np.uint8)
        c = rng.integers(0, 4, (n, L), dtype=np.uint8)
        for t in range(0, n, 2):  # every other triple is a mutated copy so large values are exercised
            ma, mb, mc = ob.mutated_triple(rng, L)
            a[t], b[t], c[t] = ma, mb, mc
        for w in (12, 16, 8):
            want = ob.score_batch(a, b, c, w)
            got = gpu_ta.TriAlign(w).align_batch(a, b, c)
            assert got.tolist() == want.tolist(), (L, n, w, gpu_ta.last_path())


def test_wrap_semantics_small_score_bits(gpu_ta, ob):
    rng = np.random.default_rng(5)
    for w in (2, 3, 5, 7, 9):
        for L in (12, 40):
            a, b, c = ob.mutated_triple(rng, L, rate=0.05)
            s, fin = gpu_ta.TriAlign(w).align(a, b, c, return_states=True)
            es, efin = ob.score(a, b, c, w)
            assert s == es and fin.tolist() == efin.tolist(), (w, L)


def test_wide_score_bits(gpu_ta, ob):
    rng = np.random.default_rng(6)
    a, b, c = ob.mutated_triple(rng, 90)
    for w in (17, 24, 32):
        s, fin = gpu_ta.TriAlign(w).align(a, b, c, return_states=True)
        es, efin = ob.score(a, b, c, w)
        assert s == es and fin.tolist() == efin.tolist(), w


def test_long_single_triple_golden(gpu_ta, golden):
    for ent in golden["synth"]:
        if not ent.get("single"):
            continue
        L = ent["len"]
        a, b, c = (gpu_ta.synth(ent["seed"], 0, s, L) for s in range(3))
        s, fin = gpu_ta.TriAlign(ent["score_bits"]).align(a, b, c, return_states=True)
        assert s == ent["score"] and fin.tolist() == ent["final_states"], (L, ent["score_bits"], gpu_ta.last_path())


def test_device_resident_entry_points_match_host_ones(gpu_ta, ob):
    import torch

    n, L = 257, 48
    a, b, c = gpu_ta.synth_batch(0x5EED0002, 1000, n, L)
    da, db, dc = gpu_ta.trialign.synth_batch_cuda(0x5EED0002, 1000, n, L)
    torch.cuda.synchronize()
    assert da.cpu().numpy().tolist() == a.tolist() and dc.cpu().numpy().tolist() == c.tolist()
    eng = gpu_ta.TriAlign(12)
    got = eng.align_batch_cuda(da, db, dc)
    torch.cuda.synchronize()
    want = ob.score_batch(a, b, c, 12)
    assert got.cpu().numpy().tolist() == want.tolist()
    one = eng.align_cuda(da[3], db[3], dc[3])
    torch.cuda.synchronize()
    es, efin = ob.score(a[3], b[3], c[3], 12)
    assert one.cpu().numpy().tolist() == [es] + efin.tolist()


def test_properties_at_scale(gpu_ta):
    """Size-independent properties on inputs too big for the oracle: identical sequences score 3L (W=16),
    and a batch made of repeats of one triple returns one repeated score."""
    eng16 = gpu_ta.TriAlign(16)
    for L in (256, 700):
        x = gpu_ta.synth(0x5EED0003, 5, 0, L)
        assert eng16.align(x, x, x) == 3 * L
    a, b, c = gpu_ta.synth_batch(0x5EED0004, 0, 1, 256)
    n = 64
    got = gpu_ta.TriAlign(12).align_batch(np.repeat(a, n, 0), np.repeat(b, n, 0), np.repeat(c, n, 0))
    assert len(set(got.tolist())) == 1
    assert got[0] == gpu_ta.TriAlign(12).align(a[0], b[0], c[0])


def test_launches_are_counted(gpu_ta):
    before = gpu_ta.launch_count()
    gpu_ta.TriAlign().align([0, 1, 2, 3], [0, 1, 2, 3], [0, 1, 2, 3])
    assert gpu_ta.launch_count() > before
    sec, cells = gpu_ta.last_timing()
    assert sec > 0 and cells == 64


@pytest.mark.parametrize("shape", [(70, 65, 129), (100, 200, 300), (300, 64, 64), (64, 300, 64), (64, 64, 300),
                                   (257, 513, 129), (1, 200, 200), (200, 1, 130), (512, 512, 512)])
def test_fast_long_kernel_shapes(gpu_ta, ob, shape):
    la, lb, lc = shape
    rng = np.random.default_rng(sum(shape))
    eng = gpu_ta.TriAlign(16)
    for mutated in (True, False):
        if mutated:
            a, b, c = ob.mutated_triple(rng, max(shape))
            a, b, c = a[:la], b[:lb], c[:lc]
        else:
            a, b, c = (rng.integers(0, 4, n, dtype=np.uint8) for n in shape)
        s, fin = eng.align(a, b, c, return_states=True)
        assert gpu_ta.last_path() == "fast_long"
        es, efin = ob.score_mt(a, b, c, 16)
        assert s == es and fin.tolist() == efin.tolist(), (shape, mutated)


@pytest.mark.parametrize("ln,n", [(65, 7), (100, 12), (130, 5), (256, 6), (200, 601), (321, 3)])
def test_fast_long_batch_kernels(gpu_ta, ob, ln, n):
    """Batches of triples longer than 64: the streaming block-sweep kernel K2S against the oracle.  n = 601 gives
    every CTA several pairs, so tasks of different pairs follow each other inside the array."""
    rng = np.random.default_rng(ln * 31 + n)
    a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    for t in range(0, n, 2):
        a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
    want = {w: ob.score_batch(a, b, c, w).tolist() for w in (12, 16)}
    for w in (12, 16):
        got = gpu_ta.TriAlign(w).align_batch(a, b, c)
        assert gpu_ta.last_path() == "fast_stream"
        assert got.tolist() == want[w], (ln, n, w)


def test_batch_of_very_long_triples_runs_one_k2l_launch_each(gpu_ta, ob):
    """Beyond K2S's 2,048 every triple of a batch is a job of its own for K2L (one launch per triple)."""
    rng = np.random.default_rng(77)
    n, ln = 2, 2100
    a, b, c = (rng.integers(0, 4, (n, ln), dtype=np.uint8) for _ in range(3))
    b[0], c[0] = a[0], a[0]  # identical sequences score 3L; the other triple against the single-triple entry point
    before = gpu_ta.launch_count()
    got = gpu_ta.TriAlign(16).align_batch(a, b, c)
    assert gpu_ta.last_path() == "fast_long" and gpu_ta.launch_count() - before == n
    assert got.tolist() == [3 * ln, gpu_ta.TriAlign(16).align(a[1], b[1], c[1])]
    # the oracle on a prefix of the same data through the same path (uniform batches of length 2,060)
    m = 2060
    got = gpu_ta.TriAlign(16).align_batch(a[1:, :m], a[1:, :m][:, ::-1].copy(), c[1:, :m])
    x, y, z = a[1, :m], a[1, :m][::-1].copy(), c[1, :m]
    assert got.tolist() == [gpu_ta.TriAlign(16).align(x, y, z)]


def test_stream_kernel_extremes_and_repeats(gpu_ta, ob):
    """All-equal sequences (largest values, 3L), all-mismatch (most negative), and the same launch repeated (face
    rings and task buffers reused across launches)."""
    L = 192
    z = np.zeros((5, L), np.uint8)
    o = np.ones((5, L), np.uint8)
    t = np.full((5, L), 2, np.uint8)
    eng = gpu_ta.TriAlign(16)
    assert eng.align_batch(z, z, z).tolist() == [3 * L] * 5
    assert gpu_ta.last_path() == "fast_stream"
    want = ob.score_batch(z, o, t, 16).tolist()
    for _ in range(3):
        assert eng.align_batch(z, o, t).tolist() == want


def test_range_checked_fast_path_and_literal_fallback(gpu_ta, ob, golden):
    # W=12 at L=1024 is outside the a-priori wrap-free envelope (3*1024 > 2047): the fast kernel runs with its
    # run-time range check; random ACGT stays far from the limits, so no fallback is needed ...
    ent = next(e for e in golden["synth"] if e.get("single") and e["len"] == 1024 and e["score_bits"] == 12)
    a, b, c = (gpu_ta.synth(ent["seed"], 0, s, 1024) for s in range(3))
    s, fin = gpu_ta.TriAlign(12).align(a, b, c, return_states=True)
    assert gpu_ta.last_path() == "fast_long"
    assert s == ent["score"] and fin.tolist() == ent["final_states"]
    # ... while near-identical sequences really wrap at W=12 (3*700 > 2047): the check fires and the literal kernel
    # reproduces the RTL's wrapped result; forcing the fast backend is refused instead of returning a wrong score
    rng = np.random.default_rng(12)
    x, y, z = ob.mutated_triple(rng, 700, rate=0.01)
    es, efin = ob.score_mt(x, y, z, 12)
    s, fin = gpu_ta.TriAlign(12).align(x, y, z, return_states=True)
    assert gpu_ta.last_path() == "fast_then_literal"
    assert s == es and fin.tolist() == efin.tolist()
    with pytest.raises(gpu_ta.TriAlignError) as ei:
        gpu_ta.TriAlign(12, backend=gpu_ta.BACKEND_GPU_FAST).align(x, y, z)
    assert ei.value.rc == -9
    # at W=16 the same triple is provably wrap-free and scores ~3*700
    s16 = gpu_ta.TriAlign(16).align(x, y, z)
    assert gpu_ta.last_path() == "fast_long" and s16 == ob.score_mt(x, y, z, 16)[0] and s16 > 1800


@pytest.mark.parametrize("shape,n_slabs", [((300, 200, 260), 2), ((512, 512, 512), 4), ((1000, 130, 129), 8)])
def test_i_slab_pipeline_back_to_back_on_one_gpu(gpu_ta, ob, shape, n_slabs):
    """The multi-GPU i-slab path with the slabs run one after the other on a single GPU: each
    slab continues from the hand-over state the previous one wrote; result must equal the single-pass score."""
    import ctypes as C

    import torch

    la, lb, lc = shape
    rng = np.random.default_rng(sum(shape))
    a, b, c = ob.mutated_triple(rng, max(shape))
    a, b, c = a[:la], b[:lb], c[:lc]
    L = gpu_ta.library()
    sb, fb = C.c_int64(0), C.c_int64(0)
    assert L.trialign_slab_buffer_bytes(lb, lc, C.byref(sb), C.byref(fb)) == 0
    bufs = []
    for _ in range(2):
        ps, pf = C.c_void_p(), C.c_void_p()
        assert L.trialign_device_alloc(0, sb.value, C.byref(ps)) == 0
        assert L.trialign_device_alloc(0, fb.value, C.byref(pf)) == 0
        bufs.append((ps, pf))
    da, db, dc = (torch.from_numpy(x.copy()).cuda() for x in (a, b, c))
    out = torch.zeros(8, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    try:
        for epoch in (1, 2):  # the buffers are reused without clearing: flags compare against the epoch
            for g in range(n_slabs):
                i0, i1 = gpu_ta.slab_range(la, g, n_slabs)
                src = bufs[(g - 1) & 1] if g > 0 else (None, None)
                dst = bufs[g & 1] if g + 1 < n_slabs else (None, None)
                rc = L.trialign_slab_score_dev(0, da.data_ptr(), la, db.data_ptr(), lb, dc.data_ptr(), lc, i0, i1,
                                               src[0], src[1], dst[0], dst[1], epoch, 16, out.data_ptr(), st)
                assert rc == 0, rc
            torch.cuda.synchronize()
            es, efin = ob.score_mt(a, b, c, 16)
            assert out.cpu().numpy().tolist() == [es] + efin.tolist()
    finally:
        for ps, pf in bufs:
            L.trialign_device_free(0, ps)
            L.trialign_device_free(0, pf)


@pytest.mark.parametrize("shape,n_ranges", [((300, 200, 260), 2), ((200, 130, 700), 3), ((512, 512, 1024), 8),
                                            ((64, 300, 1000), 4)])
def test_column_pipeline_back_to_back_on_one_gpu(gpu_ta, ob, shape, n_ranges):
    """The column pipeline (trialign_cols_score_dev) with the ranges run one after the other on a single GPU: each
    range reads the boundary faces and progress words the previous one wrote; result must equal the single-pass score.
    Run twice with different epochs over the same buffers."""
    import ctypes as C

    import torch

    la, lb, lc = shape
    rng = np.random.default_rng(sum(shape))
    a, b, c = ob.mutated_triple(rng, max(shape))
    a, b, c = a[:la], b[:lb], c[:lc]
    L = gpu_ta.library()
    fb, pb = C.c_int64(0), C.c_int64(0)
    assert L.trialign_cols_buffer_bytes(la, lb, C.byref(fb), C.byref(pb)) == 0
    bufs = []
    for _ in range(n_ranges - 1):
        pf, pp = C.c_void_p(), C.c_void_p()
        assert L.trialign_device_alloc(0, fb.value, C.byref(pf)) == 0
        assert L.trialign_device_alloc(0, max(pb.value, 256), C.byref(pp)) == 0
        bufs.append((pf, pp))
    da, db, dc = (torch.from_numpy(x.copy()).cuda() for x in (a, b, c))
    out = torch.zeros(8, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    es, efin = ob.score_mt(a, b, c, 16)
    try:
        for epoch in (1, 2):
            out.zero_()
            for g in range(n_ranges):
                c0, c1 = gpu_ta.col_range(lc, g, n_ranges)
                src = bufs[g - 1] if g > 0 else (None, None)
                dst = bufs[g] if g + 1 < n_ranges else (None, None)
                rc = L.trialign_cols_score_dev(0, da.data_ptr(), la, db.data_ptr(), lb, dc.data_ptr(), lc, c0, c1,
                                               src[0], src[1], dst[0], dst[1], epoch, 16, out.data_ptr(), st)
                assert rc == 0, rc
            torch.cuda.synchronize()
            assert out.cpu().numpy().tolist() == [es] + efin.tolist()
    finally:
        for pf, pp in bufs:
            L.trialign_device_free(0, pf)
            L.trialign_device_free(0, pp)


def test_fast_kernels_agree_with_literal_kernel_beyond_oracle_sizes(gpu_ta):
    """At sizes the CPU oracle cannot reach in seconds, the factored block-sweep kernel and the literal plane-sweep
    kernel (two independent implementations) must agree; identical sequences must score 3L."""
    a, b, c = (gpu_ta.synth(0x5EED0005, 0, s, n) for s, n in ((0, 1500), (1, 1100), (2, 900)))
    fast = gpu_ta.TriAlign(16).align(a, b, c, return_states=True)
    assert gpu_ta.last_path() == "fast_long"
    lit = gpu_ta.TriAlign(16, backend=gpu_ta.BACKEND_GPU_LITERAL).align(a, b, c, return_states=True)
    assert gpu_ta.last_path() == "literal"
    assert fast[0] == lit[0] and fast[1].tolist() == lit[1].tolist()
    x = gpu_ta.synth(0x5EED0005, 0, 0, 4096)
    assert gpu_ta.TriAlign(16).align(x, x, x) == 3 * 4096


@pytest.mark.parametrize("variant", ["256r2", "256l", "256m", "512m"])
def test_batch_kernel_variants_match_the_oracle(gpu_ta, ob, variant):
    """The A/B forms of K1 (option k1): neighbour-only synchronisation, lagged barrier, LAG-1 forms."""
    rng = np.random.default_rng(len(variant) * 7 + ord(variant[-2]))
    try:
        gpu_ta.set_option("k1", variant)
        for n, ln in ((1301, 64), (77, 41), (5, 1)):
            a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
            b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
            c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
            for t in range(0, n, 3):
                a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
            got = gpu_ta.TriAlign(12).align_batch(a, b, c)
            assert gpu_ta.last_path() == "fast_batch"
            assert got.tolist() == ob.score_batch(a, b, c, 12).tolist(), (variant, n, ln)
    finally:
        gpu_ta.set_option("k1", None)


def test_stream_kernel_with_duties_dealt_to_all_threads(gpu_ta, ob):
    """Option k2s=dist (A/B form of K2S: face fetches, border halos and face stores by thread id) against the oracle."""
    rng = np.random.default_rng(2718)
    try:
        gpu_ta.set_option("k2s", "dist")
        for n, ln in ((9, 130), (4, 256), (301, 70)):
            a = rng.integers(0, 4, (n, ln), dtype=np.uint8)
            b = rng.integers(0, 4, (n, ln), dtype=np.uint8)
            c = rng.integers(0, 4, (n, ln), dtype=np.uint8)
            for t in range(0, n, 2):
                a[t], b[t], c[t] = ob.mutated_triple(rng, ln)
            got = gpu_ta.TriAlign(16).align_batch(a, b, c)
            assert gpu_ta.last_path() == "fast_stream"
            assert got.tolist() == ob.score_batch(a, b, c, 16).tolist(), (n, ln)
    finally:
        gpu_ta.set_option("k2s", None)


# ---- traceback (trialign_align) ----
@pytest.mark.parametrize("shape,w", [((1, 1, 1), 12), ((5, 7, 3), 12), ((24, 24, 24), 7), ((40, 17, 33), 12),
                                     ((64, 64, 64), 12), ((30, 30, 30), 6), ((96, 80, 100), 16), ((128, 128, 128), 12)])
def test_traceback_matches_the_oracle_move_for_move(gpu_ta, ob, shape, w):
    """Same MAX7 tie rule on both sides, so the whole path -- not only its score -- must be identical, also where
    candidates wrap at score_bits."""
    rng = np.random.default_rng(sum(shape) * 31 + w)
    eng = gpu_ta.TriAlign(score_bits=w)
    for mutated in (True, False):
        if mutated:
            a, b, c = ob.mutated_triple(rng, max(shape), rate=0.1)
            a, b, c = a[: shape[0]], b[: shape[1]], c[: shape[2]]
        else:
            a, b, c = (rng.integers(0, 4, n, dtype=np.uint8) for n in shape)
        s, start, moves = eng.traceback(a, b, c)
        es, estart, emoves = ob.align(a, b, c, w)
        assert gpu_ta.last_path() == "traceback"
        assert (s, start) == (es, estart), (shape, w, mutated)
        assert moves.tolist() == emoves.tolist(), (shape, w, mutated)


def test_traceback_golden_dat_triple_and_edges(gpu_ta, ob, golden):
    d = golden["dat"]
    eng = gpu_ta.TriAlign(score_bits=12)
    s, start, moves = eng.traceback(d["A"], d["B"], d["C"])
    assert s == 1 and ob.rescore_path(d["A"], d["B"], d["C"], start, moves) == 1
    assert (s, start, moves.tolist()) == (lambda r: (r[0], r[1], r[2].tolist()))(ob.align(d["A"], d["B"], d["C"], 12))
    z = np.zeros(64, np.uint8)
    s, start, moves = eng.traceback(z, z, z)                      # the testbench as shipped
    assert s == 192 and start == (0, 0, 0) and moves.tolist() == [7] * 64
    rows = gpu_ta.format_alignment(z, z, z, start, moves)
    assert rows == ("A" * 64,) * 3
    e = np.zeros(0, np.uint8)
    assert eng.traceback(e, z, z)[0] == 0 and eng.traceback(e, z, z)[2].size == 0
    # capacity error instead of a truncated path
    import ctypes as C
    L = gpu_ta.library()
    u8p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
    sc, n = C.c_int32(0), C.c_int(0)
    st = np.zeros(3, np.int32)
    mv = np.zeros(8, np.uint8)
    rc = L.trialign_align(z.ctypes.data_as(u8p), 64, z.ctypes.data_as(u8p), 64, z.ctypes.data_as(u8p), 64, 12, 0,
                          C.byref(sc), st.ctypes.data_as(i32p), mv.ctypes.data_as(u8p), 8, C.byref(n))
    assert rc == -8 and n.value == 64 and sc.value == 192


@pytest.mark.parametrize("length", [256, 384])
def test_traceback_long_triple_path_is_worth_the_fast_kernels_score(gpu_ta, ob, length):
    """Beyond the sizes the CPU oracle's cube handles quickly: the path must be a valid alignment whose independent
    re-scoring equals the score the (separately verified) fast kernels report."""
    rng = np.random.default_rng(length)
    a, b, c = ob.mutated_triple(rng, length, rate=0.2)
    eng = gpu_ta.TriAlign(score_bits=16)
    s, start, moves = eng.traceback(a, b, c)
    assert s == eng.align(a, b, c) and min(start) == 0
    assert ob.rescore_path(a, b, c, start, moves) == s


@pytest.mark.parametrize("seg", [1, 5, 37, 100])
def test_traceback_checkpointed_recomputation_walks_the_same_path(gpu_ta, ob, seg):
    """Checkpointed form (segments of `kt_seg` anti-diagonal planes, every segment but the last re-swept from its
    checkpoint when the walk reaches it): the path must equal the oracle's move for move, also under wrap."""
    try:
        for shape, w in (((40, 17, 33), 12), ((64, 64, 64), 12), ((30, 30, 30), 6), ((96, 80, 100), 16), ((3, 1, 2), 12)):
            rng = np.random.default_rng(sum(shape) + seg)
            a, b, c = ob.mutated_triple(rng, max(shape), rate=0.15)
            a, b, c = a[: shape[0]], b[: shape[1]], c[: shape[2]]
            gpu_ta.set_option("kt_seg", seg)
            s, start, moves = gpu_ta.TriAlign(score_bits=w).traceback(a, b, c)
            nseg = gpu_ta.last_segments()
            assert nseg == max(1, -(-(sum(shape) - 2) // seg)), (shape, seg, nseg)
            es, estart, emoves = ob.align(a, b, c, w)
            assert (s, start) == (es, estart) and moves.tolist() == emoves.tolist(), (shape, w, seg)
    finally:
        gpu_ta.set_option("kt_seg", None)


def test_traceback_switches_to_checkpoints_when_the_cube_exceeds_the_budget(gpu_ta, ob):
    """256^3 at 16 bits: the direction cube takes 67 MB; with a 60 MB budget the path must come out of the segmented
    form (8 segments of 104 planes, 49 MB) and equal the whole-cube one; 20 MB cannot hold even that: ENOMEM."""
    rng = np.random.default_rng(256)
    a, b, c = ob.mutated_triple(rng, 256, rate=0.2)
    eng = gpu_ta.TriAlign(score_bits=16)
    whole = eng.traceback(a, b, c)
    assert gpu_ta.last_segments() == 1
    try:
        gpu_ta.set_option("kt_budget_mb", 60)
        seg = eng.traceback(a, b, c)
        assert gpu_ta.last_segments() == 8
        assert seg[0] == whole[0] and seg[1] == whole[1] and seg[2].tolist() == whole[2].tolist()
        assert ob.rescore_path(a, b, c, seg[1], seg[2]) == seg[0] == eng.align(a, b, c)
        gpu_ta.set_option("kt_budget_mb", 20)
        with pytest.raises(Exception) as e:
            eng.traceback(a, b, c)
        assert "memory" in str(e.value).lower()
    finally:
        gpu_ta.set_option("kt_budget_mb", None)


def test_traceback_golden_paths(gpu_ta, golden):
    """Every golden case (dat/ triple, prefixes, role permutations, wrap cases, ragged lengths): the GPU traceback
    must reproduce the frozen path, start cell and score."""
    for case in golden["cases"]:
        a, b, c = _abc(case)
        s, start, moves = gpu_ta.TriAlign(score_bits=case["score_bits"]).traceback(a, b, c)
        assert s == case["score"], case["name"]
        assert list(start) == case["traceback"]["start"], case["name"]
        assert "".join(str(int(m)) for m in moves) == case["traceback"]["moves"], case["name"]


# ---- all triples of a sequence set (trialign_score_all_triples) ----
@pytest.mark.parametrize("nseq,ln,w", [(3, 64, 12), (7, 40, 12), (12, 64, 12), (9, 100, 12), (6, 33, 7), (2, 10, 12)])
def test_all_triples_of_a_set_match_the_oracle(gpu_ta, ob, nseq, ln, w):
    rng = np.random.default_rng(nseq * 100 + ln)
    base = rng.integers(0, 4, ln, dtype=np.uint8)
    seqs = np.stack([base.copy() for _ in range(nseq)])
    for r in range(nseq):                                   # a family of related sequences, like an MSA input
        m = rng.random(ln) < 0.05 * (r + 1)
        seqs[r, m] = rng.integers(0, 4, int(m.sum()), dtype=np.uint8)
    idx = gpu_ta.triple_indices(nseq)
    assert len(idx) == gpu_ta.library().trialign_triple_count(nseq) == nseq * (nseq - 1) * (nseq - 2) // 6
    got = gpu_ta.TriAlign(score_bits=w).align_all_triples(seqs)
    if len(idx) == 0:
        assert got.size == 0
        return
    want = ob.score_batch(seqs[idx[:, 0]], seqs[idx[:, 1]], seqs[idx[:, 2]], w)
    assert got.tolist() == want.tolist()


def test_all_triples_of_186_sequences_is_the_million_triple_batch(gpu_ta):
    """186 sequences -> 1,055,240 triples (two gather chunks): identical to scoring the materialised batch."""
    rng = np.random.default_rng(186)
    seqs = rng.integers(0, 4, (186, 64), dtype=np.uint8)
    eng = gpu_ta.TriAlign(score_bits=12)
    got = eng.align_all_triples(seqs)
    assert got.size == 1055240 and gpu_ta.last_path() == "fast_batch"
    idx = gpu_ta.triple_indices(186)
    pick = np.concatenate([np.arange(0, 4096), np.arange(1 << 20, (1 << 20) + 2048), np.arange(got.size - 4096, got.size)])
    want = eng.align_batch(seqs[idx[pick, 0]], seqs[idx[pick, 1]], seqs[idx[pick, 2]])
    assert got[pick].tolist() == want.tolist()


# ---- explicit scoring parameters (trialign_score_ex / trialign_score_batch_ex) ----
@pytest.mark.parametrize("sc", [(1, -1, 2, 1, 0), (2, -1, 2, 1, 0), (2, 1, 2, 1, 0), (1, -1, 2, 1, 1), (5, -4, 10, 1, 1),
                                (3, -2, 0, 0, 0), (1, -3, 4, 2, 1)])
def test_explicit_scoring_matches_the_oracle(gpu_ta, ob, sc):
    """(2, 1, 2, 1) are the testbench's own unused MATCH/MISMATCH/GAP_OPEN/GAP_EXTEND localparams
    (src/TriAlign_tb.sv:37-40); rule 1 is the paper's sum of pairs.  Also at widths where candidates wrap."""
    rng = np.random.default_rng(sum(abs(x) for x in sc) + 100 * sc[4])
    scoring = gpu_ta.Scoring(*sc)
    for shape, w in [((30, 30, 30), 12), ((64, 64, 64), 12), ((20, 33, 9), 16), ((40, 40, 40), 7), ((25, 70, 66), 10)]:
        a, b, c = ob.mutated_triple(rng, max(shape), rate=0.12)
        a, b, c = a[:shape[0]], b[:shape[1]], c[:shape[2]]
        s, fin = gpu_ta.TriAlign(w, scoring=scoring).align(a, b, c, return_states=True)
        es, efin = ob.score_ex(a, b, c, w, sc)
        assert s == es and fin.tolist() == efin.tolist(), (sc, shape, w, gpu_ta.last_path())
    assert gpu_ta.last_path() == ("literal" if sc != (1, -1, 2, 1, 0) else gpu_ta.last_path())
    n, ln = 9, 48
    A = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    B = A.copy()
    B[:, ::5] = rng.integers(0, 4, B[:, ::5].shape, dtype=np.uint8)
    Cc = rng.integers(0, 4, (n, ln), dtype=np.uint8)
    got = gpu_ta.TriAlign(12, scoring=scoring).align_batch(A, B, Cc)
    assert got.tolist() == [ob.score_ex(A[t], B[t], Cc[t], 12, sc)[0] for t in range(n)]


@pytest.mark.parametrize("sc", [(2, -1, 2, 1, 0), (2, -1, 2, 1, 1), (1, -1, 2, 1, 1), (3, -2, 3, 1, 0), (2, 0, 0, 0, 0),
                                (4, -4, 8, 2, 1), (8, -8, 8, 4, 0), (3, 1, 1, 1, 0)])
def test_explicit_scoring_batches_run_the_factored_kernel(gpu_ta, ob, sc):
    """Batches of short triples with scoring parameters inside the fast domain: the range-checked K1 with the constants
    as a kernel argument, flagged triples re-scored by the literal kernel with the same parameters; every score against
    the independent table-driven oracle, at a width where nothing wraps and at widths where some triples do."""
    rng = np.random.default_rng(sum(abs(v) for v in sc) + 1000 * sc[4])
    scoring = gpu_ta.Scoring(*sc)
    for n, ln in ((61, 64), (16, 33), (5, 1)):
        A = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        B = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        Cc = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        for t in range(0, n, 3):
            A[t], B[t], Cc[t] = ob.mutated_triple(rng, ln, rate=0.1)
        for w in (16, 20, 9, 7):
            got = gpu_ta.TriAlign(w, scoring=scoring).align_batch(A, B, Cc)
            path, flagged = gpu_ta.last_path(), gpu_ta.last_flagged()
            want = [ob.score_ex(A[t], B[t], Cc[t], w, sc)[0] for t in range(n)]
            assert got.tolist() == want, (sc, n, ln, w, path)
            if w >= 16:
                assert path == "fast_batch" and flagged == 0, (sc, w, path, flagged)
            else:
                assert path in ("fast_batch", "fast_then_literal", "literal"), path


@pytest.mark.parametrize("sc", [(2, -1, 2, 1, 0), (1, -1, 2, 1, 1), (3, -2, 3, 1, 0), (4, -4, 8, 2, 1)])
def test_explicit_scoring_batches_longer_than_64_run_the_streaming_kernel(gpu_ta, ob, sc):
    """K2S with run-time scoring parameters (len 65..2048 inside the 16-bit lane range), against the oracle."""
    rng = np.random.default_rng(sum(abs(v) for v in sc) + 77)
    scoring = gpu_ta.Scoring(*sc)
    for n, ln, w in ((7, 130, 16), (5, 200, 12), (4, 96, 9)):
        A = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        B = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        Cc = rng.integers(0, 4, (n, ln), dtype=np.uint8)
        for t in range(0, n, 2):
            A[t], B[t], Cc[t] = ob.mutated_triple(rng, ln, rate=0.1)
        got = gpu_ta.TriAlign(w, scoring=scoring).align_batch(A, B, Cc)
        path = gpu_ta.last_path()
        assert got.tolist() == [ob.score_ex(A[t], B[t], Cc[t], w, sc)[0] for t in range(n)], (sc, n, ln, w, path)
        if w == 16:
            assert path == "fast_stream", path


def test_rtl_scoring_through_the_ex_entry_points_uses_the_fast_kernels(gpu_ta, ob):
    a, b, c = gpu_ta.synth_batch(0x5EED0002, 0, 64, 64)
    got = gpu_ta.TriAlign(12, scoring=gpu_ta.Scoring()).align_batch(a, b, c)
    assert gpu_ta.last_path() == "fast_batch"
    assert got.tolist() == ob.score_batch(a, b, c, 12).tolist()


def test_fast_kernels_agree_with_the_literal_kernel_on_random_shapes(gpu_ta):
    """A short run of tools/stress_gpu.py: K1 / K2S / K2L against the literal kernel (two independent GPU implementations)
    over random shapes, including very short sweeps of long b, c (la < 16: the case where a dead low half once carried
    into the live high half)."""
    import subprocess
    import sys
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_gpu.py"), "12"], capture_output=True, text=True)
    assert proc.returncode == 0 and "stress ok" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-2000:]
    eng = gpu_ta.TriAlign(12)
    lit = gpu_ta.TriAlign(12, backend=gpu_ta.BACKEND_GPU_LITERAL)
    rng = np.random.default_rng(4)
    for la, lb, lc in [(1, 154, 350), (1, 130, 365), (2, 70, 300), (3, 64, 129), (1, 64, 256), (5, 400, 400)]:
        a, b, c = (rng.integers(0, 4, n, dtype=np.uint8) for n in (la, lb, lc))
        b[:la] = a  # some real matches
        for _ in range(3):  # shared memory of the SMs holds whatever the previous launches left
            f = eng.align(a, b, c, return_states=True)
            assert gpu_ta.last_path() == "fast_long"
            g = lit.align(a, b, c, return_states=True)
            assert f[0] == g[0] and f[1].tolist() == g[1].tolist(), (la, lb, lc)


def test_explicit_scoring_fixture(gpu_ta):
    """tests/golden/golden_scoring.json (generated by make_golden_scoring.py from oracle_score_ex) through the C ABI."""
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", "golden_scoring.json")) as f:
        g = json.load(f)
    for case in g["cases"]:
        eng = gpu_ta.TriAlign(case["score_bits"], scoring=gpu_ta.Scoring(*case["scoring"]))
        s, fin = eng.align(np.array(case["a"], np.uint8), np.array(case["b"], np.uint8), np.array(case["c"], np.uint8),
                           return_states=True)
        assert s == case["score"] and fin.tolist() == case["final_states"], (case["name"], case["scoring"])


def test_packed_ram128_input_formats(gpu_ta, ob, golden):
    """The testbench's 128-bit sequence RAM words as a direct compute input (src/TriAlign_tb.sv:94-96,150-169): one
    triple and a batch, packed on the host, expanded on the GPU; code N = 4 aliases to A like in the RTL."""
    d = golden["dat"]
    eng = gpu_ta.TriAlign(score_bits=12)
    rams = [gpu_ta.pack_seq_ram128(d[k]) for k in "ABC"]
    assert eng.align_ram128(rams[0], 64, rams[1], 64, rams[2], 64) == 1
    x = np.array([4, 0, 1, 2, 3, 4, 4, 1, 2], np.uint8)
    s, fin = eng.align_ram128(gpu_ta.pack_seq_ram128(x), 9, gpu_ta.pack_seq_ram128(x[::-1].copy()), 9,
                              gpu_ta.pack_seq_ram128(x), 9, return_states=True)
    es, efin = ob.score(x, x[::-1].copy(), x, 12)
    assert s == es and fin.tolist() == efin.tolist()
    rng = np.random.default_rng(128)
    for n, ln in ((33, 64), (7, 45), (5, 130)):
        A, B, Cc = (rng.integers(0, 5, (n, ln), dtype=np.uint8) for _ in range(3))
        got = eng.align_batch_ram128(gpu_ta.pack_batch_ram128(A), gpu_ta.pack_batch_ram128(B), gpu_ta.pack_batch_ram128(Cc), n, ln)
        assert got.tolist() == ob.score_batch(A, B, Cc, 12).tolist(), (n, ln)
        assert gpu_ta.pack_batch_ram128(A)[0].tolist() == gpu_ta.pack_seq_ram128(A[0]).tolist()
