#!/usr/bin/env python3
"""Developer probe: crc32 + adler32 of one device-resident buffer (kernel time of the ranges pass + whole call)."""
import importlib.util, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << lg
data = (os.urandom(1 << 20) * max(1, n >> 20))[:n]
eng = pg.Engine(0)
d = eng.alloc(n); d.upload(data)
for it in range(4):
    eng.kernel_time_reset()
    eng.timer_start()
    r = eng.checksum_batch([(d.ptr, n, 0, 1)], mem=pg.MEM_DEVICE)
    ms = eng.timer_stop()
    kms, _ = eng.kernel_time("checksum")
print(f"crc32+adler32 of 2^{lg} B: ranges kernel {kms:.3f} ms = {n/kms/1e6:.1f} GB/s; whole call {ms:.3f} ms = {n/ms/1e6:.1f} GB/s; ok={r[0] == (zlib.crc32(data), zlib.adler32(data))}")
