"""CPU, world_size 2 over gloo: the N > 1 path of the harness -- partitioning rules, the timing max-reduce and the optional
result gather -- exercised with the CPU oracle standing in for the per-rank GPU work (a frame-sharded motion search and a
block-sharded IDCT must reassemble to exactly the unsharded result)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from libav_b200 import shard, synth


def test_split_rules():
    for n in (0, 1, 7, 68, 1 << 20):
        for world in (1, 2, 3, 8):
            spans = [shard.split_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert shard.frames_for_rank(8, 3, 8) == [3] and shard.frames_for_rank(10, 1, 4) == [1, 5, 9]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import loader
    from oracle.loader import ptr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = loader.port()
    w, h = 96, 80
    cur, ref = synth.me_frames(w, h, seed=3)
    lo, hi = shard.mb_row_range(h // 16, rank, world)
    out = np.zeros(((h // 16) * (w // 16), 3), np.int32)
    o.full_search(ptr(cur), ptr(ref), w, w, h, 16, lo, hi, ptr(out), 1)
    mbw = w // 16
    mv = shard.gather_rows(out[lo * mbw:hi * mbw], lo * mbw, hi * mbw, out.shape[0])
    blocks = synth.dense_blocks(1000, seed=4)
    b0, b1 = shard.split_range(1000, rank, world)
    mine = blocks[b0:b1].copy()
    o.idct_batch(2, ptr(mine), None, None, 0, b1 - b0, 1)
    allb = shard.gather_rows(mine, b0, b1, 1000)
    t = shard.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((mv, allb, t))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(built):
    from oracle import loader
    from oracle.loader import ptr
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    mv, allb, t = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    o = loader.port()
    w, h = 96, 80
    cur, ref = synth.me_frames(w, h, seed=3)
    want = np.zeros(((h // 16) * (w // 16), 3), np.int32)
    o.full_search(ptr(cur), ptr(ref), w, w, h, 16, 0, h // 16, ptr(want), 1)
    assert np.array_equal(mv, want)
    blocks = synth.dense_blocks(1000, seed=4)
    o.idct_batch(2, ptr(blocks), None, None, 0, 1000, 1)
    assert np.array_equal(allb, blocks)
    assert t == 11.0
