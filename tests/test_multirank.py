"""N > 1 host logic on CPU: two gloo ranks shard the blocks of a stream, each compresses its share
(with the oracle port standing in for the GPU library -- tests may use the oracle), the sizes are
exchanged, and the resulting container layout must equal the single-process one."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from libbsc_b200 import blocks as blk


def test_split_and_assign():
    parts = blk.split_blocks(10 * 1000 + 7, 1000)
    assert len(parts) == 11 and parts[-1] == (10000, 7) and sum(s for _, s in parts) == 10007
    assert blk.split_blocks(0, 1000) == []
    owned = [blk.assign(11, 4, r) for r in range(4)]
    assert sorted(sum(owned, [])) == list(range(11))
    assert owned[1] == [1, 5, 9]
    assert blk.container_offsets([100, 50]) == [18, 128]
    with pytest.raises(ValueError):
        blk.assign(4, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gen, codec = pyoracle.Gen(), pyoracle.Port()
    data = gen.text(5, 5 * 40000 + 123)
    blocks = [bytes(data[o:o + s]) for o, s in blk.split_blocks(data.size, 40000)]

    def compress(b):
        r, out = codec.compress(np.frombuffer(b, dtype=np.uint8), 1, 1, 3)
        assert r > 0
        return bytes(out)

    ids, outs, sizes, offsets = blk.run_sharded(blocks, world, rank, compress, dist=dist)
    q.put((rank, ids, [len(o) for o in outs], sizes, offsets))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from oracle import pyoracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    gen, codec = pyoracle.Gen(), pyoracle.Port()
    data = gen.text(5, 5 * 40000 + 123)
    ref_sizes = [codec.compress(data[o:o + s], 1, 1, 3)[0] for o, s in blk.split_blocks(data.size, 40000)]
    assert len(ref_sizes) == 6
    for rank, ids, my_sizes, sizes, offsets in got:
        assert ids == blk.assign(6, 2, rank)
        assert my_sizes == [ref_sizes[b] for b in ids]
        assert sizes == ref_sizes                                    # every rank learned every size
        assert offsets == blk.container_offsets(ref_sizes)
