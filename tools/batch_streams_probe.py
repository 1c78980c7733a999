#!/usr/bin/env python3
"""Developer probe: N foreign zlib-6 streams of S MiB each in ONE inflate batch, device-resident: the cost model's choice
(default) against every stream through the many-warp decode (NXGPU_INFLATE_PAR_MIN=65536) and against none (=0).
usage: batch_streams_probe.py [n=8] [log2 bytes per stream=20]"""
import gzip, importlib.util, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 20
text = pg.makedata(1, max(lg + 3, 24), alice)
eng = pg.Engine(0)
parts = [text[i * (1 << lg): (i + 1) * (1 << lg)] for i in range(n)]
blobs = [zlib.compress(p, 6) for p in parts]
packed = b"".join(blobs)
dc = eng.alloc(len(packed)); dc.upload(packed)
do = eng.alloc(n << lg)
items, o = [], 0
for i, b in enumerate(blobs):
    items.append(pg.InflateItem(dc.ptr + o, len(b), do.ptr + (i << lg), 1 << lg, pg.WRAP_ZLIB, 0)); o += len(b)
for mode in (None, "65536", "0"):
    if mode is None:
        os.environ.pop("NXGPU_INFLATE_PAR_MIN", None)
    else:
        os.environ["NXGPU_INFLATE_PAR_MIN"] = mode
    best = 1e9
    for it in range(3):
        t0 = time.perf_counter()
        res = eng.inflate_batch(items, mem=pg.MEM_DEVICE)
        best = min(best, time.perf_counter() - t0)
    ok = all(r.rc == 0 and r.crc32 == zlib.crc32(p) for r, p in zip(res, parts))
    print(f"{n} streams x {(1 << lg) >> 10} KiB, {'cost model' if mode is None else 'all many-warp' if mode != '0' else 'one warp pair each'}: {best * 1e3:.2f} ms = {(n << lg) / best / 1e9:.2f} GB/s ok={ok}", flush=True)
