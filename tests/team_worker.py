#!/usr/bin/env python3
"""One rank of a multi-process nxgpu_team (tests/test_gpu_parity.py::test_team_deflate_processes_on_two_gpus).
usage: team_worker.py <name> <rank> <nranks> <log2 total> <host|device>   -> rank 0 prints one JSON line"""
import ctypes as C
import gzip
import importlib.util
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200", "__init__.py"))
pg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pg)
name, rank, nranks, log2, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
alice = gzip.decompress(open(os.path.join(ROOT, "tests", "golden", "alice29.txt.gz"), "rb").read())
total, chunk = 1 << log2, 262144
per = (total // chunk + nranks - 1) // nranks * chunk
lo, hi = min(total, rank * per), min(total, (rank + 1) * per)
lib = pg.load_library()
shard = C.create_string_buffer(hi - lo)
assert lib.nxgpu_makedata_range(5, log2, alice, len(alice), lo, hi, shard) == hi - lo
mem = pg.MEM_HOST if mode == "host" else pg.MEM_DEVICE
with pg.Engine(rank) as eng:                       # rank r on GPU r
    team = pg.Team(eng, name, rank, nranks, total // 2 + 4096, mem)
    d = eng.alloc(hi - lo); d.upload(shard.raw)
    res = team.deflate(d.ptr, hi - lo, level=6, wrap=pg.WRAP_GZIP, chunk=chunk, src_mem=pg.MEM_DEVICE)
    res2 = team.deflate(C.addressof(shard), hi - lo, level=6, wrap=pg.WRAP_GZIP, chunk=chunk, src_mem=pg.MEM_HOST)
    assert (res.out_len, res.crc32) == (res2.out_len, res2.crc32)
    if rank == 0:
        if mem == pg.MEM_HOST:
            blob = C.string_at(team.dst(), res.out_len)
        else:
            hb = C.create_string_buffer(res.out_len)
            eng._check(eng.lib.nxgpu_memcpy_d2h(eng.ctx, C.addressof(hb), team.dst(), res.out_len), "d2h")
            blob = hb.raw
        whole = pg.makedata(5, log2, alice)
        ok = zlib.decompress(blob, 31) == whole and res.crc32 == zlib.crc32(whole) and res.src_len == total
        print(json.dumps({"ok": bool(ok), "out_len": int(res.out_len), "ms": float(res.device_ms)}))
    team.close()
