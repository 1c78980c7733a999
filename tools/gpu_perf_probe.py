#!/usr/bin/env python3
"""Developer probe: device-resident throughput of each kernel family (not the bench)."""
import ctypes as C, gzip, importlib.util, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
t = time.time(); data = pg.makedata(1, lg, alice); print(f"makedata 2^{lg} in {time.time()-t:.2f}s", flush=True)
n = len(data)
eng = pg.Engine(0)
dsrc = eng.alloc(n); dsrc.upload(data)
cap = eng.deflate_bound(n); ddst = eng.alloc(cap)
for level in (1, 6, 9):
    for it in range(3):
        eng.kernel_time_reset()
        eng.timer_start()
        res = eng.deflate_stream_device(dsrc.ptr, n, ddst.ptr, cap, level=level, wrap=pg.WRAP_GZIP)
        ms = eng.timer_stop()
        kms, kn = eng.kernel_time("deflate")
        cms, cn = eng.kernel_time("checksum")
    print(f"deflate L{level}: {n/ms/1e6:.2f} GB/s total ({ms:.2f} ms), kernel {kms:.2f} ms = {n/kms/1e6:.2f} GB/s, checksum {cms:.3f} ms = {n/cms/1e6:.1f} GB/s, ratio {n/res.out_len:.3f}, tokens {res.n_tokens}", flush=True)
# inflate of 64 KiB members
M = 65536
nm = n // M
blobs = [zlib.compress(data[i*M:(i+1)*M], 6) for i in range(min(nm, 4096))]
while len(blobs) < nm: blobs += blobs[: nm - len(blobs)]
packed = b"".join(blobs); offs = [0]
for b in blobs: offs.append(offs[-1] + len(b))
dcomp = eng.alloc(len(packed)); dcomp.upload(packed)
dout = eng.alloc(nm * M)
items = [pg.InflateItem(dcomp.ptr + offs[i], len(blobs[i]), dout.ptr + i * M, M, pg.WRAP_ZLIB, 0) for i in range(nm)]
for it in range(3):
    eng.kernel_time_reset()
    eng.timer_start(); res = eng.inflate_batch(items, mem=pg.MEM_DEVICE); ms = eng.timer_stop()
    kms, _ = eng.kernel_time("inflate"); cms, _ = eng.kernel_time("checksum")
bad = sum(1 for r in res if r.rc != 0)
print(f"inflate {nm} x 64KiB: {nm*M/ms/1e6:.2f} GB/s total ({ms:.2f} ms), kernel {kms:.2f} ms = {nm*M/kms/1e6:.2f} GB/s, checksum {cms:.3f} ms; bad={bad}", flush=True)
# checksum alone, one big buffer
for it in range(3):
    eng.kernel_time_reset()
    r = eng.checksum_batch([(dsrc.ptr, n, 0, 1)], mem=pg.MEM_DEVICE)
    cms, _ = eng.kernel_time("checksum")
print(f"crc32+adler32 of {n} B: {n/cms/1e6:.1f} GB/s ({cms:.3f} ms) ok={r[0][0]==zlib.crc32(data)}")
