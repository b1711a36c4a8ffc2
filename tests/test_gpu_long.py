"""-m gpu: long single triples against the oracle goldens of tests/golden/golden_long.json (BASELINE configs 3 and 5
at full size, ragged shapes), the exactness envelope at its boundary, stream safety of the shared scratch buffers, and
the multi-GPU paths (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import long_cases  # noqa: E402


def _case(ta, name):
    e = long_cases.load()[name]
    a, b, c = long_cases.build_case(ta.synth, e["seed"], e["kind"], e["lens"])
    assert a[:16].tolist() == e["first_symbols"]["A"] and b[:16].tolist() == e["first_symbols"]["B"]
    return e, a, b, c


LONG = sorted(long_cases.load()) if os.path.exists(long_cases.PATH) else []


@pytest.mark.parametrize("name", LONG)
def test_long_goldens_on_one_gpu(gpu_ta, name):
    """Every frozen long case, incl. config 5's 8192^3 triple at W=16, on one GPU: score and the 7 final states."""
    e, a, b, c = _case(gpu_ta, name)
    s, fin = gpu_ta.TriAlign(score_bits=e["score_bits"]).align(a, b, c, return_states=True)
    path = gpu_ta.last_path()
    assert (s, fin.tolist()) == (e["score"], e["final_states"]), (name, path)
    if name == "long_1024_mut_w12":
        # 3*1024 > 2047: values reach the 12-bit limit, the range check fires and the literal kernel re-runs
        assert path == "fast_then_literal"
    else:
        assert path == "fast_long", (name, path)


def test_config5_mutated_properties(gpu_ta):
    """Size-independent properties at config 5's size when the oracle golden is not there yet: x,x,x scores 3L."""
    L = 8192
    x = gpu_ta.synth(0x5EED0005, 0, 0, L)
    s, fin = gpu_ta.TriAlign(score_bits=16).align(x, x, x, return_states=True)
    assert s == 3 * L and fin[0] == 3 * L


@pytest.mark.parametrize("w", [7, 8, 9, 10, 11, 12])
def test_exactness_envelope_boundary(gpu_ta, ob, w):
    """provably_wrap_free() (common.cuh): 3m+5 <= 2^(W-1)-1 and 2m+12 <= 2^(W-1).  At the largest such m the unchecked
    fast kernels serve the call; one past it they must not (literal kernel, or the range-checked kernel with literal
    re-run).  All-equal (largest values) and all-different (most negative values) inputs, single triples and batches,
    every result against the oracle."""
    lim = 1 << (w - 1)
    m = min((lim - 1 - 5) // 3, (lim - 12) // 2)
    assert 3 * m + 5 <= lim - 1 and 2 * m + 12 <= lim and (3 * (m + 1) + 5 > lim - 1 or 2 * (m + 1) + 12 > lim)
    eng = gpu_ta.TriAlign(score_bits=w)
    for L in (m, m + 1, m + 7):
        if L < 1:
            continue
        z = np.zeros(L, np.uint8)
        o, t = np.ones(L, np.uint8), np.full(L, 2, np.uint8)
        for a, b, c in ((z, z, z), (z, o, t), (z, z, o), (o, z, z)):
            want = ob.score_blocked(a, b, c, w, 0, 64) if L > 200 else ob.score(a, b, c, w)
            s, fin = eng.align(a, b, c, return_states=True)
            path = gpu_ta.last_path()
            assert (s, fin.tolist()) == (want[0], want[1].tolist()), (w, L, path)
            if L <= m:
                assert path in ("fast_long", "fast_batch"), (w, L, path)
            else:
                assert path in ("literal", "fast_then_literal", "fast_long"), (w, L, path)
                if path == "fast_long":  # only through the run-time range check (10 <= W <= 16), never unchecked
                    assert 10 <= w <= 16
        if L <= 256:  # batches: ragged tail pair, same envelope rule
            n = 5
            A = np.stack([z, z, z, o, z])
            B = np.stack([z, o, z, z, z])
            Cc = np.stack([z, t, o, z, z])
            got = eng.align_batch(A, B, Cc)
            path = gpu_ta.last_path()
            assert got.tolist() == ob.score_batch(A, B, Cc, w).tolist(), (w, L, path)
            if L > m:  # never the unchecked kernels: the range-checked ones (w >= 6), flagged triples re-run literally
                assert path in ("literal", "fast_then_literal", "fast_batch", "fast_stream"), (w, L, path)
                assert path != "literal", (w, L, path)  # every length here is within K1's or K2S's reach


def test_two_streams_share_a_device_safely(gpu_ta, ob):
    """Calls in flight on different streams of one device share that device's scratch (faces, progress words):
    the library orders them on the device.  Two long triples on two streams, no host synchronisation in between."""
    import torch

    rng = np.random.default_rng(5)
    shapes = [(700, 300, 260), (300, 520, 390)]
    trip = [tuple(rng.integers(0, 4, n, dtype=np.uint8) for n in sh) for sh in shapes]
    want = [ob.score_mt(a, b, c, 16) for a, b, c in trip]
    dev = torch.device("cuda", 0)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    eng = gpu_ta.TriAlign(score_bits=16)
    tens = [[torch.from_numpy(x).to(dev) for x in t] for t in trip]
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        for (st, t) in ((s1, tens[0]), (s2, tens[1])):
            outs.append(eng.align_cuda(*t, stream=st))
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        w = want[i % 2]
        assert o.cpu().tolist() == [w[0]] + w[1].tolist(), i


def test_wide_score_bits_use_the_fast_kernels(gpu_ta, ob):
    """W in 17..32 cannot wrap at any supported length: the packed 16-bit kernels serve them, not the literal one."""
    rng = np.random.default_rng(3)
    a, b, c = ob.mutated_triple(rng, 300)
    for w in (17, 24, 32):
        s, fin = gpu_ta.TriAlign(score_bits=w).align(a, b, c, return_states=True)
        assert gpu_ta.last_path() == "fast_long"
        es, efin = ob.score(a, b, c, w)
        assert s == es and fin.tolist() == efin.tolist()
    A = rng.integers(0, 4, (33, 64), dtype=np.uint8)
    got = gpu_ta.TriAlign(score_bits=20).align_batch(A, A[::-1].copy(), A)
    assert gpu_ta.last_path() == "fast_batch"
    assert got.tolist() == ob.score_batch(A, A[::-1].copy(), A, 20).tolist()


def test_range_checked_batch_path_reruns_only_flagged_triples(gpu_ta, ob):
    """Batches whose wrap-freedom cannot be proved from the length (W=8 at len 64: 3*64 > 127): the range-checked K1
    flags, per triple, values near the limits; only those go to the literal kernel.  Random triples stay far from the
    limits, identical ones wrap for real."""
    rng = np.random.default_rng(808)
    n, L = 301, 64
    A, B, Cc = (rng.integers(0, 4, (n, L), dtype=np.uint8) for _ in range(3))
    hot = [0, 7, 150, 299, 300]
    for t in hot:
        B[t] = A[t]
        Cc[t] = A[t]
    eng = gpu_ta.TriAlign(score_bits=8)
    got = eng.align_batch(A, B, Cc)
    assert gpu_ta.last_path() == "fast_then_literal"
    assert len(hot) <= gpu_ta.last_flagged() <= 3 * len(hot) + 10
    want = ob.score_batch(A, B, Cc, 8)
    assert got.tolist() == want.tolist()
    assert want[0] != 192  # the identical triple really wrapped at 8 bits
    # forced fast backend refuses instead of guessing
    with pytest.raises(gpu_ta.TriAlignError) as ei:
        gpu_ta.TriAlign(score_bits=8, backend=gpu_ta.BACKEND_GPU_FAST).align_batch(A, B, Cc)
    assert ei.value.rc == -9
    # no false alarms from the padding beyond len, nor from random data: nothing flagged, no literal launch
    for Ls in (41, 50, 64):
        X, Y, Z = (rng.integers(0, 4, (77, Ls), dtype=np.uint8) for _ in range(3))
        got = eng.align_batch(X, Y, Z)
        assert gpu_ta.last_path() == "fast_batch" and gpu_ta.last_flagged() == 0, Ls
        assert got.tolist() == ob.score_batch(X, Y, Z, 8).tolist()
    # W=7: everything above len 19 is checked; mutated triples reach the limits often
    X, Y, Z = (np.stack(t) for t in zip(*[ob.mutated_triple(rng, 48) for _ in range(40)]))
    got = gpu_ta.TriAlign(score_bits=7).align_batch(X, Y, Z)
    assert got.tolist() == ob.score_batch(X, Y, Z, 7).tolist()
    assert gpu_ta.last_flagged() > 0


def test_range_checked_stream_path(gpu_ta, ob):
    """The same for batches longer than 64 (K2S): W=10 at len 200 (3*200 > 511)."""
    rng = np.random.default_rng(909)
    n, L = 9, 200
    A, B, Cc = (rng.integers(0, 4, (n, L), dtype=np.uint8) for _ in range(3))
    B[3] = A[3]
    Cc[3] = A[3]          # 600 > 511: wraps
    B[8], Cc[8] = A[8], A[8]
    got = gpu_ta.TriAlign(score_bits=10).align_batch(A, B, Cc)
    assert gpu_ta.last_path() == "fast_then_literal" and 2 <= gpu_ta.last_flagged() <= 4
    assert got.tolist() == ob.score_batch(A, B, Cc, 10).tolist()
    got = gpu_ta.TriAlign(score_bits=10).align_batch(A[:3], B[:3], Cc[:3])
    assert gpu_ta.last_path() == "fast_stream" and gpu_ta.last_flagged() == 0
    assert got.tolist() == ob.score_batch(A[:3], B[:3], Cc[:3], 10).tolist()


# ---- several GPUs ----
def _need_gpus(ta, n):
    if ta.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("multi", ["slab", "cols"])
@pytest.mark.parametrize("name", ["pipe_1500_w16", "pipe_1500_mut_w16", "pipe_ragged_1400_1100_900_w16"])
def test_single_process_multi_gpu_triple(gpu_ta, name, multi):
    """trialign_score with several bits in the device mask: one process drives all GPUs (peer access), the triple runs
    as an i-slab / column pipeline; score and final states against the oracle golden; the calling thread's current
    device is left alone; repeated calls reuse the hand-over buffers (epochs)."""
    _need_gpus(gpu_ta, 2)
    import torch

    e, a, b, c = _case(gpu_ta, name)
    G = min(gpu_ta.device_count(), 8)
    torch.cuda.set_device(0)
    gpu_ta.set_option("multi", multi)
    try:
        for g in sorted({2, G}):
            eng = gpu_ta.TriAlign(score_bits=e["score_bits"], devices=list(range(g)))
            for rep in range(3):
                s, fin = eng.align(a, b, c, return_states=True)
                assert (s, fin.tolist()) == (e["score"], e["final_states"]), (name, multi, g, rep)
                # the column pipeline needs a 128-column block pair per GPU; shorter inputs are served by one GPU
                want_g = g if (multi == "slab" or e["lens"][2] >= 128 * g) else 1
                assert gpu_ta.last_devices() == want_g
            assert torch.cuda.current_device() == 0
    finally:
        gpu_ta.set_option("multi", None)


def test_mask_with_a_missing_device_is_an_error(gpu_ta):
    z = np.zeros(8, np.uint8)
    n = gpu_ta.device_count()
    with pytest.raises(gpu_ta.TriAlignError) as ei:
        gpu_ta.TriAlign(devices=[0, n]).align(z, z, z)
    assert ei.value.rc == -1
    # a short triple with several GPUs in the mask is served by the first GPU alone, and says so
    if n >= 2:
        gpu_ta.TriAlign(devices=[0, 1]).align(z, z, z)
        assert gpu_ta.last_devices() == 1


def test_batch_and_all_triples_split_over_gpus(gpu_ta, ob):
    _need_gpus(gpu_ta, 2)
    rng = np.random.default_rng(11)
    n, L = 1001, 64
    A, B, Cc = (rng.integers(0, 4, (n, L), dtype=np.uint8) for _ in range(3))
    devs = list(range(min(gpu_ta.device_count(), 8)))
    got = gpu_ta.TriAlign(12, devices=devs).align_batch(A, B, Cc)
    assert gpu_ta.last_devices() == len(devs)
    assert got.tolist() == ob.score_batch(A, B, Cc, 12).tolist()
    seqs = rng.integers(0, 4, (14, 48), dtype=np.uint8)
    one = gpu_ta.TriAlign(12, devices=[0]).align_all_triples(seqs)
    many = gpu_ta.TriAlign(12, devices=devs).align_all_triples(seqs)
    assert gpu_ta.last_devices() == len(devs)
    assert one.tolist() == many.tolist()


_RANK_SCRIPT = r'''
import json, os, sys
sys.path.insert(0, os.environ["TA_ROOT"]); sys.path.insert(0, os.path.join(os.environ["TA_ROOT"], "tests", "golden"))
import torch, torch.distributed as dist
import trialign_b200 as ta, long_cases
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
cases = long_cases.load()
bad = []
for name in ("pipe_1500_w16", "pipe_1500_mut_w16", "pipe_ragged_1400_1100_900_w16"):
    e = cases[name]
    a, b, c = (torch.from_numpy(x).to(dev) for x in long_cases.build_case(ta.synth, e["seed"], e["kind"], e["lens"]))
    la, lb, lc = e["lens"]
    for kind in ("slab", "cols"):
        pipe = (ta.SlabPipeline(lb, lc, score_bits=16, device=rank) if kind == "slab"
                else ta.ColumnPipeline(la, lb, lc, score_bits=16, device=rank))
        for rep in range(3):   # back-to-back runs: epochs, slot reuse
            out = pipe.run(a, b, c)
        torch.cuda.synchronize()
        res = out.clone().to(torch.int64)
        if rank != world - 1: res.zero_()
        dist.all_reduce(res)
        got = res.cpu().tolist()
        if got != [e["score"]] + e["final_states"]:
            bad.append((name, kind, got))
        pipe.close()
dist.barrier()
if rank == 0:
    print("RESULT " + json.dumps(bad))
dist.destroy_process_group()
'''


def test_pipelines_one_process_per_gpu(gpu_ta, tmp_path):
    """SlabPipeline and ColumnPipeline as bench.py uses them: one process per GPU under torchrun (2 ranks), CUDA-IPC
    peer buffers, three runs back to back, against the oracle goldens."""
    _need_gpus(gpu_ta, 2)
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT)
    env = dict(os.environ, TA_ROOT=ROOT)
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                          env=env, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    line = [ln for ln in proc.stdout.splitlines() if ln.startswith("RESULT ")]
    assert line and line[-1] == "RESULT []", proc.stdout[-2000:]
