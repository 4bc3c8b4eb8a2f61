"""CPU, world_size 2, gloo: the multi-GPU host logic of the sequence path (frame-range
sharding, barrier, max-over-ranks timing, per-frame summaries gathered to rank 0).  The data
path itself has no collective, so each rank stands in for its GPU with the cv2 oracle chain on
a tiny configuration and the test checks that the sharded job reproduces the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from multicamera_stitching_b200.sequence import ring_chunks, shard_range


def test_shard_ranges_tile_the_sequence():
    for n in (0, 1, 7, 10, 16, 10000):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(n, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _checksum(img):
    return int(np.asarray(img, dtype=np.uint64).sum() % (2 ** 62))


def _worker(rank, world, port, n_frames, tmpdir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from helpers import synthetic_chain
    from multicamera_stitching_b200 import synthetic
    from multicamera_stitching_b200.shard import ShardContext
    from oracle import stitcher_ref

    ctx = ShardContext.from_env(backend="gloo", device="cpu")
    assert ctx.world_size == world and ctx.rank == rank
    st, states, labels, _ = synthetic_chain(3, 48, 64, 3, kind="smooth")
    lo, hi = ctx.frame_range(n_frames)
    ctx.barrier()
    sums = []
    for f in range(lo, hi):
        frames = synthetic.make_frames(3, 48, 64, 3, frame_index=f, kind="smooth")
        sums.append(_checksum(stitcher_ref.stitch_chain(states, labels, frames)))
    ctx.barrier()
    elapsed = ctx.max_over_ranks(10.0 + rank)          # stands in for a per-rank device time
    total = ctx.sum_over_ranks(hi - lo)
    gathered = ctx.gather_frame_summaries(lo, sums)
    if rank == 0:
        torch.save({"elapsed": elapsed, "total": total, "gathered": gathered}, os.path.join(tmpdir, "rank0.pt"))
    else:
        assert gathered is None
    ctx.close()


@pytest.mark.timeout(300)
def test_two_rank_sharded_sequence_equals_unsharded(tmp_path):
    n_frames, world = 7, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    res = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    assert res["elapsed"] == 11.0                      # max over ranks
    assert res["total"] == n_frames                    # every frame processed exactly once
    # unsharded reference, frame by frame
    from helpers import synthetic_chain
    from multicamera_stitching_b200 import synthetic
    from oracle import stitcher_ref
    st, states, labels, _ = synthetic_chain(3, 48, 64, 3, kind="smooth")
    want = [_checksum(stitcher_ref.stitch_chain(states, labels,
                                                synthetic.make_frames(3, 48, 64, 3, frame_index=f, kind="smooth")))
            for f in range(n_frames)]
    assert res["gathered"].tolist() == want


def test_single_rank_context_needs_no_process_group():
    from multicamera_stitching_b200.shard import ShardContext
    ctx = ShardContext(0, 1, None)
    assert ctx.frame_range(10) == (0, 10)
    ctx.barrier()
    assert ctx.max_over_ranks(3.5) == 3.5 and ctx.sum_over_ranks(2) == 2.0
    assert ctx.gather_frame_summaries(0, [1, 2, 3]).tolist() == [1, 2, 3]


def test_ring_chunks_cover_the_range_without_wrapping():
    """Config 5 cycles a long sequence through a ring of host slots: the chunks of any frame range
    cover it exactly once, in order, and never run past the end of the ring."""
    for lo, hi, ring, chunk in ((0, 100, 32, 16), (3, 14, 5, 2), (1250, 2500, 32, 16), (7, 8, 4, 16), (0, 0, 8, 4)):
        for ramp in (False, True):
            chunks = ring_chunks(lo, hi, ring, chunk, ramp=ramp)
            assert sum(n for _, n in chunks) == hi - lo
            f = lo
            for slot, n in chunks:
                assert slot == f % ring and 1 <= n <= chunk and slot + n <= ring
                f += n
    # ramped: the pipeline fills and drains on single frame-sets, full-size chunks in between
    sizes = [n for _, n in ring_chunks(0, 240, 240, 16, ramp=True)]
    assert sizes[:5] == [1, 2, 4, 8, 16] and sizes[-1] == 1 and max(sizes) == 16 and sizes.count(16) >= 10
    assert [n for _, n in ring_chunks(0, 3, 8, 16, ramp=True)] == [1, 1, 1]
    # the shards of a 10 000-frame sequence over 8 ranks tile it, and so do their chunks
    total = 0
    for rank in range(8):
        a, b = shard_range(10000, 8, rank)
        total += sum(n for _, n in ring_chunks(a, b, 32, 16))
    assert total == 10000
