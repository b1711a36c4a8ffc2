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
    L.emu_fast_batch.argtypes = [u8p, u8p, u8p, C.c_longlong, C.c_int, C.c_int, C.c_int, i32p]
    return L


def _run(L, a, b, c, pairs_per_cta):
    n, ln = a.shape
    out = np.full(n, -99999, np.int32)
    rc = L.emu_fast_batch(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), c.ctypes.data_as(u8p), n, ln, 2 * 64 + 32,
                          pairs_per_cta, out.ctypes.data_as(i32p))
    assert rc == 0
    return out


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
