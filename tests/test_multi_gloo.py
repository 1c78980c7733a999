"""World-size-2 gloo test of the multi-GPU host logic (power-gzip_b200/multi.py):
partition by chunk range, all_gather of sizes/CRCs, exclusive scan, point-to-point stitch to rank 0,
crc32_combine fold, gzip header/trailer.  The per-rank compressor is a zlib stand-in here (raw deflate
ending on a full flush, the joiner the engine also writes); on the GPU box bench.py --gpus N runs the
same functions over NCCL with the CUDA engine."""
import gzip
import os
import socket
import sys
import zlib

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, data, q):
    import importlib.util
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(ROOT, "power-gzip_b200")])
    pg = importlib.util.module_from_spec(spec); sys.modules["power_gzip_b200"] = pg; spec.loader.exec_module(pg)
    from power_gzip_b200 import multi
    lib = pg.load_library()                       # host-side combine only; no CUDA call is made
    chunk = 65536
    n_chunks = -(-len(data) // chunk)
    lo, hi = multi.partition(n_chunks, world)[rank]
    mine = data[lo * chunk: hi * chunk]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    blob = co.compress(mine) + (co.flush() if rank == world - 1 else co.flush(zlib.Z_FULL_FLUSH))
    local = torch.frombuffer(bytearray(blob) + bytearray(64), dtype=torch.uint8)
    out, total, crc, isize = multi.stitch_to_rank0(dist, torch, local, len(blob), zlib.crc32(mine), len(mine),
                                                   lambda a, b, n: int(lib.nxgpu_crc32_combine(a, b, n)))
    if rank == 0:
        q.put((bytes(out[:total].numpy().tobytes()), crc, isize))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_stitch_over_gloo(world, alice):
    import torch.multiprocessing as mp
    data = (alice * 5)[:700001]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data, q)) for r in range(world)]
    for p in procs:
        p.start()
    stream, crc, isize = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gzip.decompress(stream) == data            # one valid gzip member
    assert crc == zlib.crc32(data) and isize == len(data)


def test_partition_and_scan(pg):
    from power_gzip_b200 import multi
    assert multi.partition(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert multi.partition(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sum(hi - lo for lo, hi in multi.partition(65536, 8)) == 65536
    assert multi.exclusive_scan([5, 0, 7], 10) == [10, 15, 15, 22]
