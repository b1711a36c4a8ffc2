"""Multi-GPU plumbing, one process per GPU (torch.distributed): the two shardings of SURVEY.md 8(e).

1. Independent triples: `shard_range` gives every rank a contiguous slice of the batch; no data-path collective, the
   scores are gathered (or simply checksummed) at the end.
2. One long triple: `SlabPipeline` splits the planes i = 1..la into consecutive slabs, one per rank.  Rank g's kernel
   writes the plane state of every (j,k) block task straight into rank g+1's memory over NVLink (CUDA IPC mapping)
   and raises a per-task flag there; rank g+1's kernel starts each block as soon as its flag is up.  The traffic is a
   point-to-point chain g -> g+1; torch.distributed only carries the 64-byte IPC handles at set-up time.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

from . import trialign as _ta


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n independent triples: rank r owns [n*r/world, n*(r+1)/world)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard request")
    return (n * rank) // world, (n * (rank + 1)) // world


def slab_range(la: int, rank: int, world: int) -> Tuple[int, int]:
    """Planes (i_begin, i_end] of rank `rank`: i_begin = la*rank/world, i_end = la*(rank+1)/world."""
    if world < 1 or not (0 <= rank < world) or la < world:
        raise ValueError("need at least one plane per rank")
    return (la * rank) // world, (la * (rank + 1)) // world


def col_range(lc: int, rank: int, world: int) -> Tuple[int, int]:
    """Columns [c_begin, c_end) of C that rank `rank` sweeps in the column pipeline: whole block-column pairs (128
    columns), split as evenly as the pair count allows."""
    pairs = (lc + 127) // 128
    if world < 1 or not (0 <= rank < world) or pairs < world:
        raise ValueError("need at least one block-column pair (128 columns) per rank")
    p0, p1 = (pairs * rank) // world, (pairs * (rank + 1)) // world
    return 128 * p0, min(lc, 128 * p1)


def exchange_objects(obj, group=None) -> List:
    """all_gather of a small picklable object (IPC handles, shapes) over the given process group."""
    import torch.distributed as dist

    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, obj, group=group)
    return out


def make_pipeline(la: int, lb: int, lc: int, score_bits: int = 16, device: Optional[int] = None, group=None):
    """The pipeline trialign_score() itself would use for this triple on the group's GPUs (trialign_multi_plan):
    a SlabPipeline or a ColumnPipeline; raises if the triple does not qualify for either."""
    import torch.distributed as dist

    plan = _ta.multi_plan(la, lb, lc, score_bits, dist.get_world_size(group))
    if plan == "cols":
        return ColumnPipeline(la, lb, lc, score_bits=score_bits, device=device, group=group)
    if plan == "slab":
        return SlabPipeline(lb, lc, score_bits=score_bits, device=device, group=group)
    raise ValueError("the triple does not qualify for a multi-GPU pipeline (too short, or not provably wrap-free)")


class SlabPipeline:
    """i-slab pipeline for one long triple; build once per (lb, lc), call run() as often as needed.

    There is ONE hand-over slot per block task: run r+1 may overwrite a slot on rank g+1 only after that rank's run r
    has consumed it.  run() therefore ends with a stream-ordered all_reduce over the group (nobody's next run starts
    before every rank's kernels of this run have finished); pass sync=False only when the caller orders the runs itself.
    """

    def __init__(self, lb: int, lc: int, score_bits: int = 16, device: Optional[int] = None, group=None):
        import torch
        import torch.distributed as dist

        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.cuda.current_device() if device is None else device
        self.lb, self.lc, self.score_bits = lb, lc, score_bits
        self.epoch = 0
        self._sync = torch.zeros(1, device=torch.device("cuda", self.device))
        L = _ta.library()
        sb, fb = C.c_int64(0), C.c_int64(0)
        _ta._check(L.trialign_slab_buffer_bytes(lb, lc, C.byref(sb), C.byref(fb)), "trialign_slab_buffer_bytes")
        # hand-over buffers of THIS rank (written by rank-1 through its peer mapping); rank 0 owns none.  They are whole
        # cudaMalloc allocations made by the library, because an IPC handle names an allocation, not a sub-range.
        self.state = self.flags = None
        handles = None
        if self.rank > 0:
            ps, pf = C.c_void_p(), C.c_void_p()
            _ta._check(L.trialign_device_alloc(self.device, sb.value, C.byref(ps)), "trialign_device_alloc")
            _ta._check(L.trialign_device_alloc(self.device, fb.value, C.byref(pf)), "trialign_device_alloc")
            self.state, self.flags = ps.value, pf.value
            hs, hf = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
            _ta._check(L.trialign_ipc_export(C.c_void_p(self.state), hs), "trialign_ipc_export")
            _ta._check(L.trialign_ipc_export(C.c_void_p(self.flags), hf), "trialign_ipc_export")
            handles = (bytes(hs), bytes(hf))
        everyone = exchange_objects(handles, group)
        self.next_state = self.next_flags = None
        if self.rank + 1 < self.world:
            hs, hf = everyone[self.rank + 1]
            ps, pf = C.c_void_p(), C.c_void_p()
            _ta._check(L.trialign_ipc_open(self.device, (C.c_ubyte * 64).from_buffer_copy(hs), C.byref(ps)), "trialign_ipc_open")
            _ta._check(L.trialign_ipc_open(self.device, (C.c_ubyte * 64).from_buffer_copy(hf), C.byref(pf)), "trialign_ipc_open")
            self.next_state, self.next_flags = ps.value, pf.value

    def run(self, A, B, C_, out=None, stream=None, sync=True):
        """A, B, C_: uint8 CUDA tensors holding the WHOLE sequences on every rank.  Returns the int32[8] result tensor
        (score + 7 final states), meaningful on the last rank only."""
        import torch
        import torch.distributed as dist

        la = A.numel()
        i0, i1 = slab_range(la, self.rank, self.world)
        if out is None:
            out = torch.zeros(8, dtype=torch.int32, device=A.device)
        st = stream if stream is not None else torch.cuda.current_stream(A.device)
        self.epoch += 1
        rc = _ta.library().trialign_slab_score_dev(
            self.device, C.c_void_p(A.data_ptr()), la, C.c_void_p(B.data_ptr()), B.numel(), C.c_void_p(C_.data_ptr()),
            C_.numel(), i0, i1,
            C.c_void_p(self.state) if self.rank > 0 else None,
            C.c_void_p(self.flags) if self.rank > 0 else None,
            C.c_void_p(self.next_state) if self.next_state else None,
            C.c_void_p(self.next_flags) if self.next_flags else None,
            self.epoch, self.score_bits, C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream))
        _ta._check(rc, "trialign_slab_score_dev")
        if sync:
            with torch.cuda.stream(st):
                dist.all_reduce(self._sync, group=self.group)
        return out

    def close(self):
        L = _ta.library()
        for p in (self.next_state, self.next_flags):
            if p:
                L.trialign_ipc_close(C.c_void_p(p))
        for p in (self.state, self.flags):
            if p:
                L.trialign_device_free(self.device, C.c_void_p(p))
        self.next_state = self.next_flags = self.state = self.flags = None


class ColumnPipeline:
    """Column pipeline for one long triple (trialign_cols_score_dev): rank g sweeps all planes of its block-column
    range; the K faces crossing a range edge stream into rank g+1's memory (CUDA IPC mapping) while the sweep runs.
    Build once per (la, lb, lc); a run must have finished on every rank before the next one starts: run() ends with a
    stream-ordered all_reduce over the group unless sync=False."""

    def __init__(self, la: int, lb: int, lc: int, score_bits: int = 16, device: Optional[int] = None, group=None):
        import torch
        import torch.distributed as dist

        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.cuda.current_device() if device is None else device
        self.la, self.lb, self.lc, self.score_bits = la, lb, lc, score_bits
        self.c_begin, self.c_end = col_range(lc, self.rank, self.world)
        self.epoch = 0
        self._sync = torch.zeros(1, device=torch.device("cuda", self.device))
        L = _ta.library()
        fb, pb = C.c_int64(0), C.c_int64(0)
        _ta._check(L.trialign_cols_buffer_bytes(la, lb, C.byref(fb), C.byref(pb)), "trialign_cols_buffer_bytes")
        self.face = self.prog = None
        handles = None
        if self.rank > 0:  # boundary buffers of THIS rank, written by rank-1 through its peer mapping
            pf, pp = C.c_void_p(), C.c_void_p()
            _ta._check(L.trialign_device_alloc(self.device, fb.value, C.byref(pf)), "trialign_device_alloc")
            _ta._check(L.trialign_device_alloc(self.device, max(pb.value, 256), C.byref(pp)), "trialign_device_alloc")
            self.face, self.prog = pf.value, pp.value
            hf, hp = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
            _ta._check(L.trialign_ipc_export(C.c_void_p(self.face), hf), "trialign_ipc_export")
            _ta._check(L.trialign_ipc_export(C.c_void_p(self.prog), hp), "trialign_ipc_export")
            handles = (bytes(hf), bytes(hp))
        everyone = exchange_objects(handles, group)
        self.next_face = self.next_prog = None
        if self.rank + 1 < self.world:
            hf, hp = everyone[self.rank + 1]
            pf, pp = C.c_void_p(), C.c_void_p()
            _ta._check(L.trialign_ipc_open(self.device, (C.c_ubyte * 64).from_buffer_copy(hf), C.byref(pf)), "trialign_ipc_open")
            _ta._check(L.trialign_ipc_open(self.device, (C.c_ubyte * 64).from_buffer_copy(hp), C.byref(pp)), "trialign_ipc_open")
            self.next_face, self.next_prog = pf.value, pp.value

    def run(self, A, B, C_, out=None, stream=None, sync=True):
        """A, B, C_: uint8 CUDA tensors holding the WHOLE sequences on every rank.  Returns the int32[8] result tensor
        (score + 7 final states), meaningful on the last rank only."""
        import torch
        import torch.distributed as dist

        if out is None:
            out = torch.zeros(8, dtype=torch.int32, device=A.device)
        st = stream if stream is not None else torch.cuda.current_stream(A.device)
        self.epoch = self.epoch % 2047 + 1
        rc = _ta.library().trialign_cols_score_dev(
            self.device, C.c_void_p(A.data_ptr()), A.numel(), C.c_void_p(B.data_ptr()), B.numel(),
            C.c_void_p(C_.data_ptr()), C_.numel(), self.c_begin, self.c_end,
            C.c_void_p(self.face) if self.rank > 0 else None,
            C.c_void_p(self.prog) if self.rank > 0 else None,
            C.c_void_p(self.next_face) if self.next_face else None,
            C.c_void_p(self.next_prog) if self.next_prog else None,
            self.epoch, self.score_bits, C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream))
        _ta._check(rc, "trialign_cols_score_dev")
        if sync:
            with torch.cuda.stream(st):
                dist.all_reduce(self._sync, group=self.group)
        return out

    def close(self):
        L = _ta.library()
        for p in (self.next_face, self.next_prog):
            if p:
                L.trialign_ipc_close(C.c_void_p(p))
        for p in (self.face, self.prog):
            if p:
                L.trialign_device_free(self.device, C.c_void_p(p))
        self.next_face = self.next_prog = self.face = self.prog = None
