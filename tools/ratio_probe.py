#!/usr/bin/env python3
"""Developer probe: level-6 output size of the ratio fixtures relative to zlib -6 (gate: <= 1.05) and kernel speed.
Run with NXGPU_LZ_PARAMS=depth,lazy,nice,d1 to try other parse parameters."""
import gzip, importlib.util, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
eng = pg.Engine(0)
out = []
for name, data in [("s1", pg.makedata(1, 20, alice)), ("s4", pg.makedata(4, 20, alice)), ("s5", pg.makedata(5, 20, alice)), ("alice", alice)]:
    got = len(eng.compress(data, level=6, wrap=pg.WRAP_ZLIB))
    out.append(f"{name} {got / len(zlib.compress(data, 6)):.4f}")
big = pg.makedata(1, 27, alice)
d = eng.alloc(len(big)); d.upload(big); cap = eng.deflate_bound(len(big)); o = eng.alloc(cap)
best = 1e9
for _ in range(3):
    eng.kernel_time_reset()
    res = eng.deflate_stream_device(d.ptr, len(big), o.ptr, cap, level=6, wrap=pg.WRAP_GZIP)
    best = min(best, eng.kernel_time("deflate")[0])
print(os.environ.get("NXGPU_LZ_PARAMS", "default"), "| size/zlib6:", ", ".join(out), f"| 2^27 seed1: {len(big)/best/1e6:.2f} GB/s ratio {len(big)/res.out_len:.3f}", flush=True)
