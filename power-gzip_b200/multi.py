"""Chunk-range partitioning of one deflate stream across the GPUs of a box.

Rank r compresses bytes [r*n, (r+1)*n) as raw deflate whose every chunk ends on the empty stored
block the reference uses as joiner (lib/nx_deflate.c:220-243), BFINAL only on the last rank.  The
only exchange is small: sizes + CRCs (one all_gather), an exclusive scan of the sizes, the compressed
ranges sent to rank 0 at their offsets, and the CRCs folded with crc32_combine (lib/nx_crc.c:374)
exactly as the host code folds the checksums of consecutive jobs (lib/nx_deflate.c:1562-1577).

Works on any torch.distributed backend: NCCL with device tensors on the GPU box (bench.py), gloo
with CPU tensors in the tests.  No compression happens here.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

GZIP_HEADER = bytes([0x1f, 0x8b, 0x08, 0, 0, 0, 0, 0, 0, 0x03])   # lib/nx_deflate.c:473-489 (blank header)


def partition(n_chunks: int, world: int) -> List[Tuple[int, int]]:
    """contiguous chunk ranges [lo, hi) per rank; the first n_chunks % world ranks get one more"""
    base, extra = divmod(n_chunks, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def exclusive_scan(sizes: Sequence[int], base: int = 0) -> List[int]:
    """offsets[r] = base + sum(sizes[:r]); one extra entry with the total"""
    offs = [base]
    for s in sizes:
        offs.append(offs[-1] + int(s))
    return offs


def fold_crc32(combine: Callable[[int, int, int], int], crcs: Sequence[int], lens: Sequence[int]) -> int:
    """crc of the concatenation from per-range crcs (seed 0 each) and uncompressed lengths"""
    crc = 0
    for c, n in zip(crcs, lens):
        crc = combine(crc, int(c) & 0xffffffff, int(n))
    return crc


def stitch_to_rank0(dist, torch, local: "torch.Tensor", local_size: int, local_crc: int, local_len: int,
                    combine: Callable[[int, int, int], int], out: Optional["torch.Tensor"] = None):
    """All ranks call this with their compressed range (a uint8 tensor, first local_size bytes valid).
    Returns (stream tensor or None, total bytes, crc32, isize) — the tensor only on rank 0: one gzip
    member = blank header + the ranges in rank order + CRC32 + ISIZE."""
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = torch.tensor([int(local_size), int(local_crc), int(local_len)], dtype=torch.int64, device=local.device)
    allm = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(allm, meta)
    rows = [[int(x) for x in m.tolist()] for m in allm]
    sizes = [r[0] for r in rows]
    offs = exclusive_scan(sizes, len(GZIP_HEADER))
    total = offs[-1] + 8
    crc = fold_crc32(combine, [r[1] for r in rows], [r[2] for r in rows])
    isize = sum(r[2] for r in rows) & 0xffffffff
    batched = hasattr(dist, "batch_isend_irecv") and hasattr(dist, "P2POp")
    if rank != 0:
        if local_size:
            if batched:
                # one grouped launch per rank: NCCL runs the seven transfers into rank 0 concurrently instead of one
                # after the other (unbatched point-to-point ops are serialised on the process group)
                for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, local[:local_size], 0)]):
                    q.wait()
            else:
                dist.send(local[:local_size], dst=0)
        return None, total, crc, isize
    if out is None or out.numel() < total:
        out = torch.empty(total, dtype=torch.uint8, device=local.device)
    out[:len(GZIP_HEADER)] = torch.tensor(list(GZIP_HEADER), dtype=torch.uint8, device=local.device)
    out[offs[0]:offs[1]].copy_(local[:sizes[0]])
    if batched:
        ops = [dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r) for r in range(1, world) if sizes[r]]
        reqs = dist.batch_isend_irecv(ops) if ops else []
    else:
        reqs = [dist.irecv(out[offs[r]:offs[r + 1]], src=r) for r in range(1, world) if sizes[r]]
    for q in reqs:
        q.wait()
    trailer = list(crc.to_bytes(4, "little") + isize.to_bytes(4, "little"))
    out[offs[-1]:total] = torch.tensor(trailer, dtype=torch.uint8, device=local.device)
    return out, total, crc, isize
