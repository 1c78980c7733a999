#!/usr/bin/env python3
"""Developer probe: batched inflate of 64 KiB members, device resident (not the bench).
usage: inflate_probe.py [log2 distinct bytes=26] [members=16384] [level=6]"""
import ctypes as C, gzip, importlib.util, os, sys, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("power_gzip_b200", os.path.join(ROOT, "power-gzip_b200/__init__.py"))
pg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pg)
alice = gzip.decompress(open(os.path.join(ROOT, "tests/golden/alice29.txt.gz"), "rb").read())
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
nm = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
level = int(sys.argv[3]) if len(sys.argv) > 3 else 6
data = pg.makedata(int(os.environ.get("SEED", "1")), lg, alice)
M = 65536
nd = len(data) // M
blobs = [zlib.compress(data[i * M:(i + 1) * M], level) for i in range(nd)]
# a few text members too (alice29 slices): different statistics than makedata
for i in range(0, len(alice) - M, M):
    blobs[i // M] = zlib.compress(alice[i:i + M], level)
    data = data[:(i // M) * M] + alice[i:i + M] + data[(i // M + 1) * M:]
packed = b"".join(blobs)
offs = [0]
for b in blobs:
    offs.append(offs[-1] + len(b))
eng = pg.Engine(0)
dcomp = eng.alloc(len(packed)); dcomp.upload(packed)
dout = eng.alloc(nm * M)
items = [pg.InflateItem(dcomp.ptr + offs[i % nd], len(blobs[i % nd]), dout.ptr + i * M, M, pg.WRAP_ZLIB, 0) for i in range(nm)]
arr = (pg.InflateItem * nm)(*items)
res = (pg.InflateResult * nm)()
best = 1e9
for it in range(4):
    eng.kernel_time_reset()
    eng._check(eng.lib.nxgpu_inflate_batch(eng.ctx, arr, nm, res, pg.MEM_DEVICE), "inflate")
    kms, _ = eng.kernel_time("inflate")
    best = min(best, kms)
bad = sum(1 for r in res if r.rc != 0 or r.out_len != M)
ok = all(dout.download(M, i * M) == data[(i % nd) * M:(i % nd + 1) * M] for i in list(range(0, nm, max(1, nm // 64))) + [nm - 1, 1, 2])
csum = sum(len(b) for b in blobs) / nd
print(f"inflate {nm} x 64KiB (zlib -{level}, avg member {csum:.0f} B): kernel {best:.3f} ms = {nm*M/best/1e6:.2f} GB/s; bad={bad} data_ok={ok}", flush=True)
