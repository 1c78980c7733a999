#!/usr/bin/env python3
"""Developer probe: tiny members through the batched inflate, many times (hunting nondeterminism)."""
import ctypes as C, importlib.util, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
eng = pg.Engine(0)
def streams(data):
    yield "zlib6", zlib.compress(data, 6)
    yield "zlib1", zlib.compress(data, 1)
    yield "zlib0", zlib.compress(data, 0)
    c = zlib.compressobj(6, zlib.DEFLATED, 31); yield "gzip", c.compress(data) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, -15); yield "raw", c.compress(data) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED); yield "fixed", c.compress(data) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    parts = [c.compress(data[i:i + 7000]) + c.flush(zlib.Z_SYNC_FLUSH) for i in range(0, len(data), 7000)]
    yield "sync", b"".join(parts) + c.flush()
bad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    for pad in range(0, 4):
        keep, items, meta = [], [], []
        for d in (b"", b"a", b"ab" * 5, b"hello world" * 3, bytes(300)):
            for name, b in streams(d):
                sb = C.create_string_buffer(bytes(pad) + b, len(b) + pad)
                ob = (C.c_char * (len(d) + 32))()
                keep.append((sb, ob))
                items.append(pg.InflateItem(C.addressof(sb) + pad, len(b), C.addressof(ob), len(d) + 32, pg.WRAP_AUTO, 0))
                meta.append((name, d, b))
        res = eng.inflate_batch(items, mem=pg.MEM_HOST)
        for r, (name, d, b), (sb, ob) in zip(res, meta, keep):
            if r.rc != 0 or bytes(memoryview(ob)[: r.out_len]) != d:
                bad += 1
                print("BAD", rep, pad, name, len(d), b.hex(), "rc", r.rc, "out", r.out_len, "used", r.in_used, flush=True)
print("bad", bad)
