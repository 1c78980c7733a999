#!/usr/bin/env python3
"""Developer probe: batched inflate of distinct 64 KiB gzip members of the seed-4 stream (BASELINE.json configs[2]),
device resident.  usage: inflate_members_probe.py [members=16384] [seed=4] [log2=33]"""
import ctypes as C, gzip, importlib.util, os, sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4
log2 = int(sys.argv[3]) if len(sys.argv) > 3 else 33
M = 65536
lib = pg.load_library()
raw = C.create_string_buffer(nm * M)
lib.nxgpu_makedata_range(seed, log2, alice, len(alice), 0, nm * M, raw)
mv = memoryview(raw).cast("B")
with ThreadPoolExecutor(os.cpu_count()) as ex:
    blobs = list(ex.map(lambda i: zlib.compress(mv[i * M:(i + 1) * M], 6, wbits=31), range(nm)))
packed = b"".join(blobs)
eng = pg.Engine(0)
dcomp = eng.alloc(len(packed)); dcomp.upload(packed)
dout = eng.alloc(nm * M)
items = (pg.InflateItem * nm)()
o = 0
for i, b in enumerate(blobs):
    items[i] = pg.InflateItem(dcomp.ptr + o, len(b), dout.ptr + i * M, M, pg.WRAP_GZIP, 0); o += len(b)
res = (pg.InflateResult * nm)()
for name in ("warp-per-member",):
    best = 1e9
    for it in range(4):
        eng.kernel_time_reset()
        eng._check(lib.nxgpu_inflate_batch(eng.ctx, items, nm, res, pg.MEM_DEVICE), "inflate")
        ms, k = eng.kernel_time("inflate")
        best = min(best, ms)
    ok = all(r.rc == 0 and r.out_len == M for r in res) and res[nm - 1].crc32 == zlib.crc32(mv[(nm - 1) * M: nm * M])
    print(f"seed {seed} -b {log2} {name}: {nm} members, kernel {best:.2f} ms = {nm * M / best / 1e6:.1f} GB/s out, compressed {len(packed) / nm:.0f} B/member, ok={ok}", flush=True)
