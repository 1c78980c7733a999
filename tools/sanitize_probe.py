#!/usr/bin/env python3
"""Developer probe for compute-sanitizer runs: a few small streams through every inflate path (batch kernel, warp-pair
kernel, many-warp decode incl. the dry run), a small deflate and a checksum call.
usage: compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_probe.py"""
import ctypes as C, gzip, importlib.util, os, random, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
eng = pg.Engine(0)
rnd = random.Random(1)
data = (alice * 3)[:400000]
mix = alice[:60000] + rnd.randbytes(40000) + bytes(50000) + alice[60000:120000]
ok = True
for name, d, lvl in (("text6", data, 6), ("text1", data, 1), ("mix9", mix, 9), ("stored", rnd.randbytes(100000), 0)):
    z = zlib.compress(d, lvl)
    for mode in ("65536", "0"):
        os.environ["NXGPU_INFLATE_PAR_MIN"] = mode
        sb = C.create_string_buffer(z, len(z)); ob = (C.c_char * len(d))()
        r = eng.inflate_batch([pg.InflateItem(C.addressof(sb), len(z), C.addressof(ob), len(d), pg.WRAP_AUTO, 0)], mem=pg.MEM_HOST)[0]
        good = r.rc == 0 and bytes(ob) == d
        ok &= good
        print(name, "par" if mode != "0" else "pair", "ok" if good else f"FAILED rc {r.rc}", flush=True)
os.environ["NXGPU_INFLATE_SOLO_MAX"] = "0"
sb = C.create_string_buffer(zlib.compress(data, 6)); ob = (C.c_char * len(data))()
os.environ["NXGPU_INFLATE_PAR_MIN"] = "0"
r = eng.inflate_batch([pg.InflateItem(C.addressof(sb), len(sb.raw) - 1, C.addressof(ob), len(data), pg.WRAP_AUTO, 0)], mem=pg.MEM_HOST)[0]
print("batch kernel", r.rc == 0 and bytes(ob) == data, flush=True)
del os.environ["NXGPU_INFLATE_SOLO_MAX"]
os.environ["NXGPU_INFLATE_PAR_MIN"] = "65536"
blob = gzip.compress(data, 6, mtime=0) + gzip.compress(mix, 6, mtime=0)
got, m = eng.gunzip(blob, len(data) + len(mix))
print("gunzip two members", got == data + mix and m == 2, flush=True)
res = eng.compress(data, level=6, wrap=pg.WRAP_GZIP, chunk=65536)
print("deflate", gzip.decompress(res[0] if isinstance(res, tuple) else res) == data, flush=True)
print("crc", eng.crc32(data) == zlib.crc32(data), flush=True)
print("ALL OK" if ok else "FAILURES")
