#!/usr/bin/env python3
"""Developer smoke script for a GPU box: runs each kernel family on small inputs and checks
against zlib (stdlib) so failures are easy to localise.  Not part of the test suite."""
import ctypes as C, gzip, importlib.util, os, sys, time, zlib, random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)

alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
size_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 22
eng = pg.Engine(0)
ok = True

def check(name, cond, extra=""):
    global ok
    print(("PASS " if cond else "FAIL ") + name, extra, flush=True)
    ok = ok and cond

# ---- checksums ----
for n in (0, 1, 5, 63, 64, 65, 4096, 32768, 32769, 100000, len(alice)):
    d = alice[:n]
    c, a = eng.crc32(d), eng.adler32(d)
    check(f"cksum n={n}", c == zlib.crc32(d) and a == zlib.adler32(d), f"{c:08x}/{zlib.crc32(d):08x} {a:08x}/{zlib.adler32(d):08x}")
d = alice[3:70001]
check("cksum seeded+unaligned", eng.crc32(d, 0x12345678) == zlib.crc32(d, 0x12345678) and eng.adler32(d, 0xabcd0123) == zlib.adler32(d, 0xabcd0123))

# ---- deflate ----
md = pg.makedata(1, size_log2, alice)
for name, data in (("empty", b""), ("tiny", b"hello hello hello hello"), ("alice", alice), ("zeros", bytes(300000)),
                   ("random", random.Random(1).randbytes(200000)), ("makedata", md)):
    for level in (1, 6):
        for wrap in (pg.WRAP_GZIP, pg.WRAP_ZLIB, pg.WRAP_RAW):
            t = time.time()
            try:
                blob = eng.compress(data, level=level, wrap=wrap)
            except Exception as e:
                check(f"deflate {name} L{level} w{wrap}", False, repr(e)); continue
            dt = time.time() - t
            try:
                if wrap == pg.WRAP_GZIP: back = gzip.decompress(blob)
                elif wrap == pg.WRAP_ZLIB: back = zlib.decompress(blob)
                else: back = zlib.decompress(blob, -15)
                good = back == data
            except Exception as e:
                good = False; back = repr(e)
            zl = len(zlib.compress(data, level))
            check(f"deflate {name} L{level} w{wrap}", good, f"in={len(data)} out={len(blob)} zlib={zl} ratio_vs_zlib={len(blob)/max(zl,1):.3f} {dt*1e3:.1f}ms" + ("" if good else f" back={str(back)[:80]}"))

# ---- inflate ----
blobs, caps, exp = [], [], []
for i, data in enumerate([b"", b"a", alice, bytes(100000), md[:65536], md[65536:200000], random.Random(2).randbytes(70000)]):
    for mk in (lambda d: gzip.compress(d, 6, mtime=0), lambda d: zlib.compress(d, 6), lambda d: zlib.compress(d, 1), lambda d: zlib.compress(d, 0),
               lambda d: (lambda c: c.compress(d) + c.flush())(zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED))):
        b = mk(data)
        blobs.append(b); caps.append(len(data) + 16); exp.append(data)
try:
    outs = eng.uncompress_many(blobs, caps)
    bad = [i for i, (o, e) in enumerate(zip(outs, exp)) if o != e]
    check("inflate batch", not bad, f"n={len(blobs)} bad={bad[:10]}")
except Exception as e:
    check("inflate batch", False, repr(e))
# round trip own deflate -> own inflate
try:
    blob = eng.compress(md, level=6, wrap=pg.WRAP_GZIP)
    out = eng.uncompress(blob, len(md))
    check("roundtrip own inflate", out == md, f"{len(md)} -> {len(blob)}")
except Exception as e:
    check("roundtrip own inflate", False, repr(e))
print("launches", eng.launch_count(), "deflate", eng.kernel_time("deflate"), "inflate", eng.kernel_time("inflate"), "checksum", eng.kernel_time("checksum"))
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
