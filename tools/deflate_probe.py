#!/usr/bin/env python3
"""Developer probe: device-resident deflate of makedata text (not the bench).
usage: deflate_probe.py [log2 bytes=28] [levels=6] [seed=1]"""
import gzip, importlib.util, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
levels = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "6").split(",")]
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
data = pg.makedata(seed, lg, alice)
n = len(data)
eng = pg.Engine(0)
dsrc = eng.alloc(n); dsrc.upload(data)
cap = eng.deflate_bound(n); ddst = eng.alloc(cap)
for level in levels:
    best = 1e9
    for it in range(3):
        eng.kernel_time_reset()
        res = eng.deflate_stream_device(dsrc.ptr, n, ddst.ptr, cap, level=level, wrap=pg.WRAP_GZIP)
        kms, kn = eng.kernel_time("deflate")
        best = min(best, kms)
    blob = ddst.download(res.out_len)
    ok = zlib.crc32(gzip.decompress(blob)) == zlib.crc32(data) if n <= (1 << 28) else None
    print(f"deflate L{level} seed {seed} 2^{lg}: kernel {best:.2f} ms = {n/best/1e6:.2f} GB/s, ratio {n/res.out_len:.3f}, tokens {res.n_tokens}, roundtrip={ok}", flush=True)
