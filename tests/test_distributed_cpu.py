"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard / slab partitioning and the handle exchange that
sets up the i-slab pipeline.  The CUDA side of the pipeline is covered by the -m gpu tests (slabs run back to back on
one GPU) and by bench.py --gpus N on the GPU box."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import trialign_b200 as ta
        from importlib import import_module

        par = import_module("hw-accelerator-three-sequence-alignment_b200.parallel")
        # 1. independent triples: contiguous shards cover the batch exactly once, no exchange needed
        n = 1001
        lo, hi = ta.shard_range(n, rank, world)
        covered = torch.zeros(n, dtype=torch.int32)
        covered[lo:hi] = 1
        dist.all_reduce(covered)
        assert int(covered.min()) == 1 and int(covered.max()) == 1
        # every rank scores its shard; a checksum of per-rank checksums replaces the gather
        local = torch.tensor([sum(range(lo, hi))], dtype=torch.int64)
        dist.all_reduce(local)
        assert int(local.item()) == sum(range(n))
        # 2. one long triple: slabs tile the planes, chain order g -> g+1
        la = 8192
        i0, i1 = ta.slab_range(la, rank, world)
        spans = par.exchange_objects((i0, i1))
        assert spans[0][0] == 0 and spans[-1][1] == la
        assert all(spans[g][1] == spans[g + 1][0] for g in range(world - 1))
        # 2b. column pipeline: whole block-column pairs (128 columns) per rank, tiling C
        for lc in (8192, 1000, 128 * world, 128 * world + 5):
            c0, c1 = ta.col_range(lc, rank, world)
            cols = par.exchange_objects((c0, c1))
            assert cols[0][0] == 0 and cols[-1][1] == lc
            assert all(cols[g][1] == cols[g + 1][0] and cols[g][1] % 128 == 0 for g in range(world - 1))
        # 3. the set-up exchange: every rank > 0 publishes two 64-byte handles, rank g maps those of rank g+1
        mine = (bytes([rank] * 64), bytes([rank + 100] * 64)) if rank > 0 else None
        everyone = par.exchange_objects(mine)
        if rank + 1 < world:
            nxt = everyone[rank + 1]
            assert nxt[0] == bytes([rank + 1] * 64) and nxt[1] == bytes([rank + 101] * 64)
        assert everyone[0] is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioning_and_handle_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_partition_edge_cases():
    sys.path.insert(0, ROOT)
    import trialign_b200 as ta

    assert ta.shard_range(0, 0, 4) == (0, 0)
    assert [ta.shard_range(5, r, 8) for r in range(8)] == [(0, 0), (0, 1), (1, 1), (1, 2), (2, 3), (3, 3), (3, 4), (4, 5)]
    assert [ta.slab_range(8192, r, 8) for r in range(8)][3] == (3072, 4096)
    with pytest.raises(ValueError):
        ta.slab_range(3, 0, 4)
    with pytest.raises(ValueError):
        ta.shard_range(10, 4, 4)
    assert [ta.col_range(8192, r, 8) for r in range(8)][5] == (5120, 6144)
    assert [ta.col_range(300, r, 3) for r in range(3)] == [(0, 128), (128, 256), (256, 300)]
    with pytest.raises(ValueError):
        ta.col_range(300, 0, 4)  # three column pairs cannot feed four ranks
