"""libbsc_b200/blocks.py -- host-side block partitioning for multi-GPU runs.

libbsc blocks are independent: the only inter-block structure is the CLI container's
table of (offset, size) pairs (bsc.cpp:50-57, 401-417).  So the multi-GPU path is "shard whole
blocks over ranks, no data-path collective"; the one exchange step is the tiny gather of the
compressed sizes that lets rank 0 lay the blocks out in file order.  This module holds exactly that
logic, backend-agnostic (NCCL on the GPU box, gloo in the CPU tests).
"""
from typing import Callable, List, Sequence, Tuple


def split_blocks(total_bytes: int, block_bytes: int) -> List[Tuple[int, int]]:
    """(offset, size) of every block of a stream cut CLI-style into fixed-size blocks (bsc.cpp:163-178)."""
    if total_bytes < 0 or block_bytes <= 0:
        raise ValueError("bad sizes")
    out, off = [], 0
    while off < total_bytes:
        size = min(block_bytes, total_bytes - off)
        out.append((off, size))
        off += size
    return out


def assign(nblocks: int, world: int, rank: int) -> List[int]:
    """Block b -> rank b mod world; returns this rank's block ids in file order."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, nblocks, world))


def gather_sizes(dist, my_ids: Sequence[int], my_sizes: Sequence[int], nblocks: int, device=None) -> List[int]:
    """All ranks learn every block's compressed size (one all_reduce of an int64 vector)."""
    import torch
    v = torch.zeros(nblocks, dtype=torch.int64, device=device)
    for b, s in zip(my_ids, my_sizes):
        v[b] = int(s)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v)                 # each block is owned by exactly one rank -> SUM == gather
    return [int(x) for x in v.tolist()]


def container_offsets(sizes: Sequence[int], header_bytes: int = 8, per_block_header: int = 10) -> List[int]:
    """File offset of every compressed block in a bsc container: 'bsc1' + int32 nBlocks, then per block a
    10-byte record {int64 offset, int8 recordSize, int8 sortingContexts} followed by the block (bsc.cpp:401-417)."""
    out, off = [], header_bytes
    for s in sizes:
        off += per_block_header
        out.append(off)
        off += int(s)
    return out


def run_sharded(blocks: Sequence[bytes], world: int, rank: int, compress: Callable[[bytes], bytes], dist=None, device=None):
    """Compress this rank's share of `blocks` with `compress` and return (my_ids, my_outputs, all_sizes, offsets)."""
    ids = assign(len(blocks), world, rank)
    outs = [compress(blocks[b]) for b in ids]
    sizes = gather_sizes(dist, ids, [len(o) for o in outs], len(blocks), device=device)
    return ids, outs, sizes, container_offsets(sizes)
