"""Host-side mirror of the reference's only interface to this path: the ``TRIALIGN`` module port list
(/root/reference/src/TriAlign_1cyc.v:1-22) as the testbench drives it (src/TriAlign_tb.sv:204-262).

    reference                                   here
    ------------------------------------------  -----------------------------------------------
    parameter SCORE_BITS                        TriAlign(score_bits=...)
    A_idx / B_idx / C_idx + symbol RAMs         TriAlign.align(A, B, C)  (lengths = len of each)
    start_align pulse ... finish                the call itself
    Score (signed SCORE_BITS)                   return value
    no error signalling at all                  TriAlignError with the C ABI's code

All work happens in ``lib/libtrialign.so``; if it is missing or no CUDA device is present the calls raise --
there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("TRIALIGN_LIB") or os.path.join(_HERE, "lib", "libtrialign.so")  # override: development A/B builds

BACKEND_AUTO, BACKEND_GPU, BACKEND_GPU_LITERAL, BACKEND_GPU_FAST = 0, 4, 5, 6
PATH_NAMES = {0: "none", 1: "literal", 2: "fast_batch", 3: "fast_long", 4: "fast_then_literal", 5: "traceback", 6: "fast_stream"}

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)


TRIPLE_RTL, TRIPLE_SUM_OF_PAIRS = 0, 1


class Scoring(C.Structure):
    """``trialign_scoring``: the values the testbench's commented-out .i_match / .i_mismatch / .i_gap_open /
    .i_gap_extend ports would carry (src/TriAlign_tb.sv:220-224); defaults = the PE localparams (src/PE_1cyc.v:53-56)."""

    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_extend", C.c_int32),
                ("triple_rule", C.c_int32)]

    def __init__(self, match=1, mismatch=-1, gap_open=2, gap_extend=1, triple_rule=TRIPLE_RTL):
        super().__init__(match, mismatch, gap_open, gap_extend, triple_rule)

    def as_list(self):
        return [self.match, self.mismatch, self.gap_open, self.gap_extend, self.triple_rule]


class TriAlignError(RuntimeError):
    def __init__(self, rc: int, what: str):
        self.rc = rc
        detail = ""
        try:
            msg = library().trialign_strerror(rc).decode()
            cuda = library().trialign_last_cuda_error().decode()
            detail = f": {msg}" + (f" [{cuda}]" if cuda and rc in (-2, -3, -7) else "")
        except Exception:  # pragma: no cover - library itself unavailable
            pass
        super().__init__(f"{what} failed with rc={rc}{detail}")


_lib: Optional[C.CDLL] = None


def library_path() -> str:
    return _LIB


def build_library(force: bool = False) -> str:
    """Compile lib/libtrialign.so for sm_100a with the package Makefile (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
    proc = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building libtrialign.so failed:\n" + proc.stdout[-4000:] + proc.stderr[-4000:])
    return _LIB


def library() -> C.CDLL:
    """Load the C-ABI library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise FileNotFoundError(f"{_LIB} is missing: run __graft_entry__.build() (make -C {_HERE}) first")
    L = C.CDLL(_LIB)
    L.trialign_version.restype = C.c_int
    L.trialign_strerror.argtypes = [C.c_int]
    L.trialign_strerror.restype = C.c_char_p
    L.trialign_last_cuda_error.restype = C.c_char_p
    L.trialign_device_count.restype = C.c_int
    L.trialign_score.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_uint, _i32p, _i32p]
    L.trialign_score.restype = C.c_int
    L.trialign_triple_count.argtypes = [C.c_int]
    L.trialign_triple_count.restype = C.c_int64
    L.trialign_score_all_triples.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, _i32p]
    L.trialign_score_all_triples.restype = C.c_int
    L.trialign_align.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_uint, _i32p, _i32p, _u8p,
                                 C.c_int, C.POINTER(C.c_int)]
    L.trialign_align.restype = C.c_int
    L.trialign_score_batch.argtypes = [_u8p, _u8p, _u8p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint, _i32p]
    L.trialign_score_batch.restype = C.c_int
    L.trialign_scoring_rtl.argtypes = [C.POINTER(Scoring)]
    L.trialign_scoring_rtl.restype = None
    L.trialign_score_ex.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.POINTER(Scoring), C.c_uint,
                                    _i32p, _i32p]
    L.trialign_score_ex.restype = C.c_int
    L.trialign_score_batch_ex.argtypes = [_u8p, _u8p, _u8p, C.c_int64, C.c_int, C.c_int, C.POINTER(Scoring), C.c_uint,
                                          _i32p]
    L.trialign_score_batch_ex.restype = C.c_int
    L.trialign_score_batch_dev.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p]
    L.trialign_score_batch_dev.restype = C.c_int
    L.trialign_score_dev.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p]
    L.trialign_score_dev.restype = C.c_int
    L.trialign_read_seq_dat.argtypes = [C.c_char_p, _u8p, C.c_int, C.POINTER(C.c_int)]
    L.trialign_read_seq_dat.restype = C.c_int
    L.trialign_write_seq_dat.argtypes = [C.c_char_p, _u8p, C.c_int, C.c_int]
    L.trialign_write_seq_dat.restype = C.c_int
    L.trialign_write_scores.argtypes = [C.c_char_p, _i32p, C.c_int64]
    L.trialign_write_scores.restype = C.c_int
    L.trialign_read_scores.argtypes = [C.c_char_p, _i32p, C.c_int64, C.POINTER(C.c_int64)]
    L.trialign_read_scores.restype = C.c_int
    L.trialign_pack_seq_ram128.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
    L.trialign_pack_seq_ram128.restype = C.c_int
    L.trialign_unpack_seq_ram128.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
    L.trialign_unpack_seq_ram128.restype = C.c_int
    L.trialign_write_tb_init.argtypes = [C.c_char_p, _u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, C.c_int]
    L.trialign_write_tb_init.restype = C.c_int
    L.trialign_synth.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, _u8p]
    L.trialign_synth.restype = None
    L.trialign_synth_batch_dev.argtypes = [C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
    L.trialign_synth_batch_dev.restype = C.c_int
    L.trialign_slab_buffer_bytes.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.trialign_slab_buffer_bytes.restype = C.c_int
    L.trialign_slab_score_dev.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]
    L.trialign_slab_score_dev.restype = C.c_int
    L.trialign_cols_buffer_bytes.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.trialign_cols_buffer_bytes.restype = C.c_int
    L.trialign_cols_score_dev.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]
    L.trialign_cols_score_dev.restype = C.c_int
    L.trialign_device_alloc.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
    L.trialign_device_alloc.restype = C.c_int
    L.trialign_device_free.argtypes = [C.c_int, C.c_void_p]
    L.trialign_device_free.restype = C.c_int
    L.trialign_ipc_export.argtypes = [C.c_void_p, C.POINTER(C.c_ubyte)]
    L.trialign_ipc_export.restype = C.c_int
    L.trialign_ipc_open.argtypes = [C.c_int, C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]
    L.trialign_ipc_open.restype = C.c_int
    L.trialign_ipc_close.argtypes = [C.c_void_p]
    L.trialign_ipc_close.restype = C.c_int
    L.trialign_last_timing.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.trialign_last_timing.restype = C.c_int
    L.trialign_last_path.restype = C.c_int
    L.trialign_last_devices.restype = C.c_int
    L.trialign_last_flagged.restype = C.c_int64
    L.trialign_last_segments.restype = C.c_int
    L.trialign_score_ram128.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_uint, _i32p,
                                        _i32p]
    L.trialign_score_ram128.restype = C.c_int
    L.trialign_score_batch_ram128.argtypes = [_u8p, _u8p, _u8p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint, _i32p]
    L.trialign_score_batch_ram128.restype = C.c_int
    L.trialign_multi_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.trialign_multi_plan.restype = C.c_int
    L.trialign_set_option.argtypes = [C.c_char_p, C.c_char_p]
    L.trialign_set_option.restype = C.c_int
    L.trialign_get_option.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    L.trialign_get_option.restype = C.c_int
    L.trialign_launch_count.restype = C.c_int64
    _lib = L
    return L


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise TriAlignError(rc, what)


def _u8(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint8))


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def version() -> int:
    return int(library().trialign_version())


def device_count() -> int:
    return int(library().trialign_device_count())


def launch_count() -> int:
    return int(library().trialign_launch_count())


def last_path() -> str:
    return PATH_NAMES.get(int(library().trialign_last_path()), "?")


def last_devices() -> int:
    """GPUs that served the last compute call of this thread (a device mask with several bits may still be served by
    one GPU when the input does not qualify for the multi-GPU pipeline)."""
    return int(library().trialign_last_devices())


def multi_plan(la: int, lb: int, lc: int, score_bits: int, n_devices: int) -> str:
    """How one long triple is spread over n_devices GPUs: "single", "slab" or "cols" (trialign_multi_plan)."""
    rc = int(library().trialign_multi_plan(la, lb, lc, score_bits, n_devices))
    _check(min(rc, 0), "trialign_multi_plan")
    return ("single", "slab", "cols")[rc]


def last_flagged() -> int:
    """Triples of the last batch call that the range check handed to the literal kernel."""
    return int(library().trialign_last_flagged())


def last_segments() -> int:
    """Segments of the last align() call (1 = whole direction cube kept, > 1 = checkpointed recomputation)."""
    return int(library().trialign_last_segments())


def set_option(name: str, value=None) -> None:
    """Run-time option of the library (kernel A/B variants, gate intervals, multi-GPU decomposition); None restores
    the default.  The library never reads the environment."""
    v = None if value is None else str(value).encode()
    _check(library().trialign_set_option(name.encode(), v), f"trialign_set_option({name})")


def get_option(name: str) -> str:
    buf = C.create_string_buffer(64)
    _check(library().trialign_get_option(name.encode(), buf, 64), f"trialign_get_option({name})")
    return buf.value.decode()


def apply_option_args(pairs) -> None:
    """Helper for the tools' `--opt name=value` flags."""
    for kv in pairs or []:
        k, _, v = kv.partition("=")
        set_option(k, v or None)


def last_timing() -> Tuple[float, float]:
    """(kernel seconds, cells) of the last host-pointer compute call of this thread."""
    s, c = C.c_double(), C.c_double()
    library().trialign_last_timing(C.byref(s), C.byref(c))
    return float(s.value), float(c.value)


def _devmask(devices: Optional[Sequence[int]]) -> int:
    if not devices:
        return 0
    m = 0
    for d in devices:
        m |= 1 << int(d)
    return m


class TriAlign:
    """One configured aligner, the counterpart of a ``TRIALIGN #(.SCORE_BITS(W))`` instance."""

    def __init__(self, score_bits: int = 12, backend: int = BACKEND_AUTO, devices: Optional[Sequence[int]] = None,
                 scoring: Optional[Scoring] = None):
        self.score_bits = int(score_bits)
        self.backend = int(backend)
        self.devices = list(devices) if devices else None
        self.scoring = scoring  # None: the RTL's constants through the plain entry points

    # -- host-memory entry points (these are what the reference-side binding would call) --
    def align(self, A, B, C_, return_states: bool = False):
        a, b, c = _u8(A), _u8(B), _u8(C_)
        s = C.c_int32(0)
        fin = np.zeros(7, np.int32)
        if self.scoring is not None:
            rc = library().trialign_score_ex(_p(a, _u8p), a.size, _p(b, _u8p), b.size, _p(c, _u8p), c.size,
                                             self.score_bits, C.byref(self.scoring), _devmask(self.devices),
                                             C.byref(s), _p(fin, _i32p))
            _check(rc, "trialign_score_ex")
            return (int(s.value), fin) if return_states else int(s.value)
        rc = library().trialign_score(_p(a, _u8p), a.size, _p(b, _u8p), b.size, _p(c, _u8p), c.size,
                                      self.score_bits, self.backend, _devmask(self.devices), C.byref(s),
                                      _p(fin, _i32p))
        _check(rc, "trialign_score")
        return (int(s.value), fin) if return_states else int(s.value)

    def align_ram128(self, ram_a, la: int, ram_b, lb: int, ram_c, lc: int, return_states: bool = False):
        """One triple given as testbench RAM images (pack_seq_ram128): 128-bit words of 4-bit symbols."""
        ra, rb, rc_ = (np.ascontiguousarray(np.asarray(x, np.uint8)).reshape(-1) for x in (ram_a, ram_b, ram_c))
        s = C.c_int32(0)
        fin = np.zeros(7, np.int32)
        rc = library().trialign_score_ram128(_p(ra, _u8p), la, _p(rb, _u8p), lb, _p(rc_, _u8p), lc, self.score_bits,
                                             self.backend, _devmask(self.devices), C.byref(s), _p(fin, _i32p))
        _check(rc, "trialign_score_ram128")
        return (int(s.value), fin) if return_states else int(s.value)

    def align_batch_ram128(self, ram_a, ram_b, ram_c, n: int, length: int) -> np.ndarray:
        """A uniform batch as RAM images: three uint8 arrays [n][ceil(length/32)][16] (pack_batch_ram128)."""
        ra, rb, rc_ = (np.ascontiguousarray(np.asarray(x, np.uint8)).reshape(-1) for x in (ram_a, ram_b, ram_c))
        need = n * ((length + 31) // 32) * 16
        if ra.size < need or rb.size < need or rc_.size < need:
            raise ValueError("RAM images are smaller than n x ceil(length/32) words")
        out = np.zeros(n, np.int32)
        rc = library().trialign_score_batch_ram128(_p(ra, _u8p), _p(rb, _u8p), _p(rc_, _u8p), n, length, self.score_bits,
                                                   self.backend, _devmask(self.devices), _p(out, _i32p))
        _check(rc, "trialign_score_batch_ram128")
        return out

    def traceback(self, A, B, C_):
        """Optimal alignment of one triple -> (score, start cell (i0,j0,k0), moves uint8[columns]).  moves are in
        forward order, bit0/1/2 = A/B/C advances (7 = all three aligned); see format_alignment()."""
        a, b, c = _u8(A), _u8(B), _u8(C_)
        s = C.c_int32(0)
        start = np.zeros(3, np.int32)
        cap = a.size + b.size + c.size + 1
        moves = np.zeros(cap, np.uint8)
        n = C.c_int(0)
        rc = library().trialign_align(_p(a, _u8p), a.size, _p(b, _u8p), b.size, _p(c, _u8p), c.size, self.score_bits,
                                      _devmask(self.devices), C.byref(s), _p(start, _i32p), _p(moves, _u8p), cap,
                                      C.byref(n))
        _check(rc, "trialign_align")
        return int(s.value), tuple(int(x) for x in start), moves[: n.value].copy()

    def align_all_triples(self, seqs) -> np.ndarray:
        """Scores of every triple (x < y < z, roles A=x, B=y, C=z) of a sequence set [nseq][len], in the order of
        triple_indices(nseq).  The set is uploaded once; the triples are laid out on the device."""
        s = _u8(seqs)
        if s.ndim != 2:
            raise ValueError("align_all_triples wants one [nseq][len] array")
        nseq, ln = s.shape
        out = np.zeros(int(library().trialign_triple_count(nseq)), np.int32)
        rc = library().trialign_score_all_triples(_p(s, _u8p), nseq, ln, self.score_bits, self.backend,
                                                  _devmask(self.devices), _p(out, _i32p))
        _check(rc, "trialign_score_all_triples")
        return out

    def align_batch(self, A, B, C_) -> np.ndarray:
        a, b, c = _u8(A), _u8(B), _u8(C_)
        if a.ndim != 2 or a.shape != b.shape or a.shape != c.shape:
            raise ValueError("align_batch wants three [n][len] arrays of identical shape")
        n, ln = a.shape
        out = np.zeros(n, np.int32)
        if self.scoring is not None:
            rc = library().trialign_score_batch_ex(_p(a, _u8p), _p(b, _u8p), _p(c, _u8p), n, ln, self.score_bits,
                                                   C.byref(self.scoring), _devmask(self.devices), _p(out, _i32p))
            _check(rc, "trialign_score_batch_ex")
            return out
        rc = library().trialign_score_batch(_p(a, _u8p), _p(b, _u8p), _p(c, _u8p), n, ln, self.score_bits,
                                            self.backend, _devmask(self.devices), _p(out, _i32p))
        _check(rc, "trialign_score_batch")
        return out

    # -- device-memory entry points: torch CUDA tensors (uint8 [n][len]) in, int32 tensor out --
    def align_batch_cuda(self, A, B, C_, out=None, stream=None):
        import torch

        if not (A.is_cuda and B.is_cuda and C_.is_cuda):
            raise ValueError("align_batch_cuda wants CUDA tensors")
        if A.dtype != torch.uint8 or A.dim() != 2 or A.shape != B.shape or A.shape != C_.shape:
            raise ValueError("align_batch_cuda wants three uint8 [n][len] tensors of identical shape")
        A, B, C_ = A.contiguous(), B.contiguous(), C_.contiguous()
        n, ln = A.shape
        if out is None:
            out = torch.empty(n, dtype=torch.int32, device=A.device)
        st = stream if stream is not None else torch.cuda.current_stream(A.device)
        rc = library().trialign_score_batch_dev(A.device.index or 0, A.data_ptr(), B.data_ptr(), C_.data_ptr(), n, ln,
                                                self.score_bits, self.backend, out.data_ptr(), st.cuda_stream)
        _check(rc, "trialign_score_batch_dev")
        return out

    def align_cuda(self, A, B, C_, out=None, stream=None):
        """One triple resident on the GPU -> int32[8] tensor (score, then the 7 final states)."""
        import torch

        A, B, C_ = A.contiguous(), B.contiguous(), C_.contiguous()
        if out is None:
            out = torch.empty(8, dtype=torch.int32, device=A.device)
        st = stream if stream is not None else torch.cuda.current_stream(A.device)
        rc = library().trialign_score_dev(A.device.index or 0, A.data_ptr(), A.numel(), B.data_ptr(), B.numel(),
                                          C_.data_ptr(), C_.numel(), self.score_bits, self.backend, out.data_ptr(),
                                          st.cuda_stream)
        _check(rc, "trialign_score_dev")
        return out


def triple_indices(nseq: int) -> np.ndarray:
    """[C(nseq,3)][3] indices (x < y < z) in the order trialign_score_all_triples scores them (lexicographic)."""
    from itertools import combinations

    return np.array(list(combinations(range(nseq), 3)), dtype=np.int32).reshape(-1, 3)


def format_alignment(A, B, C_, start, moves, alphabet: str = "ATCGN") -> Tuple[str, str, str]:
    """Three gapped rows for the path TriAlign.traceback() returns (testbench symbol names: A=0 T=1 C=2 G=3 N=4)."""
    seqs = (_u8(A), _u8(B), _u8(C_))
    pos = [int(x) for x in start]
    rows = ["", "", ""]
    for mv in (int(m) for m in moves):
        for w in range(3):
            if (mv >> w) & 1:
                rows[w] += alphabet[int(seqs[w][pos[w]]) & 7] if int(seqs[w][pos[w]]) < len(alphabet) else "?"
                pos[w] += 1
            else:
                rows[w] += "-"
    return rows[0], rows[1], rows[2]


# ---- file formats of the reference's dat/ directory ----
def read_seq_dat(path: str, cap: int = 1 << 20) -> np.ndarray:
    buf = np.zeros(cap, np.uint8)
    n = C.c_int(0)
    _check(library().trialign_read_seq_dat(os.fsencode(path), _p(buf, _u8p), cap, C.byref(n)), "trialign_read_seq_dat")
    return buf[: n.value].copy()


def write_seq_dat(path: str, seq, crlf: bool = True) -> None:
    s = _u8(seq)
    _check(library().trialign_write_seq_dat(os.fsencode(path), _p(s, _u8p), s.size, int(crlf)), "trialign_write_seq_dat")


def write_scores(path: str, scores) -> None:
    s = np.ascontiguousarray(np.asarray(scores, dtype=np.int32))
    _check(library().trialign_write_scores(os.fsencode(path), _p(s, _i32p), s.size), "trialign_write_scores")


def read_scores(path: str, cap: int = 1 << 22) -> np.ndarray:
    buf = np.zeros(cap, np.int32)
    n = C.c_int64(0)
    _check(library().trialign_read_scores(os.fsencode(path), _p(buf, _i32p), cap, C.byref(n)), "trialign_read_scores")
    return buf[: n.value].copy()


# ---- the testbench's packed sequence RAMs and stimulus block (src/TriAlign_tb.sv:94-96, :423-1959) ----
def pack_seq_ram128(seq) -> np.ndarray:
    """Symbols -> the testbench RAM image: uint8 [ceil(n/32)][16], one row per 128-bit word, LSB first."""
    s = _u8(seq)
    words = (s.size + 31) // 32
    ram = np.zeros((words, 16), np.uint8)
    _check(library().trialign_pack_seq_ram128(_p(s, _u8p), s.size, _p(ram, _u8p), words), "trialign_pack_seq_ram128")
    return ram


def pack_batch_ram128(seqs) -> np.ndarray:
    """[n][len] symbols -> [n][ceil(len/32)][16] RAM images (vectorised pack_seq_ram128, one image per row)."""
    s = np.ascontiguousarray(np.asarray(seqs, dtype=np.uint8))
    n, ln = s.shape
    words = (ln + 31) // 32
    pad = np.zeros((n, words * 32), np.uint8)
    pad[:, :ln] = s & 15
    return (pad[:, 0::2] | (pad[:, 1::2] << 4)).reshape(n, words, 16)


def unpack_seq_ram128(ram, n: int) -> np.ndarray:
    """The first n 4-bit symbol codes of a RAM image produced by pack_seq_ram128 (or dumped from the testbench)."""
    r = np.ascontiguousarray(np.asarray(ram, dtype=np.uint8)).reshape(-1)
    if n > 2 * r.size:
        raise ValueError("RAM image holds fewer than n symbols")
    out = np.zeros(max(n, 0), np.uint8)
    _check(library().trialign_unpack_seq_ram128(_p(r, _u8p), n, _p(out, _u8p), out.size), "trialign_unpack_seq_ram128")
    return out


def write_tb_init(path: str, a, b, c, total_len: int = 512) -> None:
    """Writes the `seq?_ram[w][hi:lo] <= SYM;` lines that replace the testbench's initial block for this triple."""
    a, b, c = _u8(a), _u8(b), _u8(c)
    _check(library().trialign_write_tb_init(os.fsencode(path), _p(a, _u8p), a.size, _p(b, _u8p), b.size, _p(c, _u8p),
                                            c.size, total_len), "trialign_write_tb_init")


# ---- synthetic inputs ----
def synth(seed: int, triple: int, which: int, length: int) -> np.ndarray:
    out = np.zeros(max(length, 0), np.uint8)
    library().trialign_synth(C.c_uint64(seed), triple, which, length, _p(out, _u8p))
    return out


def synth_batch(seed: int, first: int, n: int, length: int):
    """Host arrays [n][len] for A, B, C (slow path; benchmarks generate on the device instead)."""
    a = np.zeros((n, length), np.uint8)
    b = np.zeros((n, length), np.uint8)
    c = np.zeros((n, length), np.uint8)
    L = library()
    for t in range(n):
        L.trialign_synth(C.c_uint64(seed), first + t, 0, length, _p(a[t], _u8p))
        L.trialign_synth(C.c_uint64(seed), first + t, 1, length, _p(b[t], _u8p))
        L.trialign_synth(C.c_uint64(seed), first + t, 2, length, _p(c[t], _u8p))
    return a, b, c


def synth_batch_cuda(seed: int, first: int, n: int, length: int, device=0, stream=None):
    import torch

    dev = torch.device("cuda", device) if isinstance(device, int) else device
    a = torch.empty((n, length), dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    c = torch.empty_like(a)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    rc = library().trialign_synth_batch_dev(dev.index or 0, C.c_uint64(seed), first, n, length, a.data_ptr(),
                                            b.data_ptr(), c.data_ptr(), st.cuda_stream)
    _check(rc, "trialign_synth_batch_dev")
    return a, b, c
