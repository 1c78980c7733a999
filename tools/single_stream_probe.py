#!/usr/bin/env python3
"""Developer probe: ONE stream on an otherwise idle GPU (the LD_PRELOAD single-z_stream case): inflate of one
zlib member on one warp, deflate of one 1 MiB job on one CTA."""
import gzip, importlib.util, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
data = pg.makedata(1, 24, alice)
eng = pg.Engine(0)
blob = zlib.compress(data, 6)
dc = eng.alloc(len(blob)); dc.upload(blob)
do = eng.alloc(len(data))
for it in range(2):
    eng.kernel_time_reset()
    r = eng.inflate_batch([pg.InflateItem(dc.ptr, len(blob), do.ptr, len(data), pg.WRAP_ZLIB, 0)], mem=pg.MEM_DEVICE)[0]
    kms, _ = eng.kernel_time("inflate")
print(f"one 16 MiB zlib member on one warp: {kms:.1f} ms = {len(data)/kms/1e3:.1f} MB/s, rc {r.rc}")
one = data[: 1 << 20]
ds = eng.alloc(len(one)); ds.upload(one)
dd = eng.alloc(2 << 20)
for level in (1, 6):
    for it in range(2):
        eng.kernel_time_reset()
        res = eng.deflate_stream_device(ds.ptr, len(one), dd.ptr, 2 << 20, level=level, wrap=pg.WRAP_RAW, chunk=1 << 20)
        kms, _ = eng.kernel_time("deflate")
    print(f"one 1 MiB deflate job on one CTA, level {level}: {kms:.2f} ms = {len(one)/kms/1e3:.1f} MB/s")
